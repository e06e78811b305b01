#!/bin/bash
# One GPU session: parity of the new traversal kernel, A/B timing of the library variants, then the bench and profiles of the
# library that is in the tree.  Everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
V=prt_b200/csrc/variants

echo "== parity (new kernel, default tuning)" | tee $O/session.log
timeout 420 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/parity.log 2>&1
echo "parity exit $?" | tee -a $O/session.log
tail -5 $O/parity.log | tee -a $O/session.log

echo "== A/B" | tee -a $O/session.log
: > $O/ab.jsonl
if [ -f $V/base.so ]; then
  echo '{"lib": "base"}' >> $O/ab.jsonl
  PRT_B200_LIB=$PWD/$V/base.so timeout 300 python tools/sweep.py --reps 5 >> $O/ab.jsonl 2>> $O/ab.err
fi
if [ -f $V/occl0.so ]; then
  echo '{"lib": "occl0"}' >> $O/ab.jsonl
  PRT_B200_LIB=$PWD/$V/occl0.so timeout 300 python tools/sweep.py --reps 5 work_list=0,1 >> $O/ab.jsonl 2>> $O/ab.err
fi
echo '{"lib": "tree"}' >> $O/ab.jsonl
timeout 400 python tools/sweep.py --reps 5 work_list=0,1 horizon_near=36,38,40 >> $O/ab.jsonl 2>> $O/ab.err
cat $O/ab.jsonl | cut -c1-260 | tee -a $O/session.log

echo "== bench (tree library)" | tee -a $O/session.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench exit $?" | tee -a $O/session.log
cut -c1-400 $O/bench_n1.json | tee -a $O/session.log

echo "== launch list" | tee -a $O/session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
echo "ncu launches exit $?" | tee -a $O/session.log

echo "== ncu full (traversal + horizon kernels)" | tee -a $O/session.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave|horizon_kernel" -s 4 -c 2 -f -o $O/wave_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $O/session.log
ls -la $O | tee -a $O/session.log
