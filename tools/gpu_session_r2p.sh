#!/bin/bash
# Round-2 session P (1 GPU): the state to be judged -- horizon slabs + gain rule (defaults horizon_near 157, horizon_mid 24, horizon_gain 64),
# no slab filter.  Whole GPU test-suite, library variants, bench lines of both workloads and configs 3 / 4, launch list, full ncu captures.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2p.log
: > $L
echo "== pytest -m gpu (whole suite)" | tee -a $L
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x > $O/r2p_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2p_pytest.log)" | tee -a $L
: > $O/r2p_sweep.jsonl
sw() { echo "{\"sweep\": \"$*\"}" >> $O/r2p_sweep.jsonl; timeout 600 python tools/sweep.py --reps 4 --flush "$@" >> $O/r2p_sweep.jsonl 2>> $O/r2p_sweep.err; }
sw --mesh torus horizon_mid=24 horizon_gain=64,32
sw --mesh torus horizon_near=30 horizon_mid=0 horizon_slabs=0
for v in prmt_imm minb6 hzminb8; do
  echo "{\"lib\": \"$v\"}" >> $O/r2p_sweep.jsonl
  PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2p_sweep.jsonl 2>> $O/r2p_sweep.err
done
sw --mesh torus --world 8 --rank 3 horizon_mid=24,16
cut -c1-330 $O/r2p_sweep.jsonl | tee -a $L
echo "== bench lines" | tee -a $L
timeout 600 python bench.py > $O/r2p_bench_n1.json 2> $O/r2p_bench_n1.err; echo "bench exit $?" | tee -a $L
timeout 600 python bench.py --workload folds > $O/r2p_bench_c1_folds.json 2> $O/r2p_bench_folds.err; echo "folds exit $?" | tee -a $L
timeout 600 python bench.py --config 4 > $O/r2p_bench_c4.json 2> $O/r2p_bench_c4.err; echo "c4 exit $?" | tee -a $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2p_bench_n1_reference.json 2> $O/r2p_bench_ref.err; echo "ref exit $?" | tee -a $L
for f in n1 c1_folds c4 n1_reference; do cut -c1-260 $O/r2p_bench_$f.json | tee -a $L; done
echo "== launch list" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2p_launches_bench.log 2>&1
echo "exit $?" | tee -a $L
echo "== ncu full" | tee -a $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave|horizon_kernel" -s 8 -c 2 -f -o $O/r2p_full_torus python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2p_ncu_torus.log 2>&1
echo "exit $?" | tee -a $L
tail -3 $O/r2p_sweep.err | cut -c1-300 | tee -a $L
