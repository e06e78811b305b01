#!/bin/bash
# Final GPU session of a round: the whole GPU test suite, the bench line, the ncu launch list and one full ncu capture of the two
# headline kernels, and the timing lines of the other configs.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O

echo "== pytest -m gpu" | tee $O/session.log
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a $O/session.log
tail -3 $O/pytest_gpu.log | tee -a $O/session.log

echo "== smoke" | tee -a $O/session.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $O/session.log

echo "== bench" | tee -a $O/session.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench exit $?" | tee -a $O/session.log
cut -c1-300 $O/bench_n1.json | tee -a $O/session.log

echo "== launch list" | tee -a $O/session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1
echo "ncu launches exit $?" | tee -a $O/session.log

echo "== ncu full (traversal + horizon kernels)" | tee -a $O/session.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave|horizon_kernel" -s 4 -c 2 -f -o $O/wave_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_full.log 2>&1
echo "ncu full exit $?" | tee -a $O/session.log

echo "== other configs" | tee -a $O/session.log
timeout 900 python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err
echo "configs exit $?" | tee -a $O/session.log
timeout 600 python tools/bench_configs.py --only c5 --budget 128 --near 20 > $O/config5_tuned.jsonl 2>> $O/configs.err
cut -c1-200 $O/configs.jsonl $O/config5_tuned.jsonl | tee -a $O/session.log
ls -la $O | tee -a $O/session.log
