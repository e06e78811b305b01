#!/bin/bash
# Round-2 session U (1 GPU): final artefacts of the round -- whole GPU suite, smoke, default bench + reference arm, folds / config 4 / 5
# lines, launch list, full ncu captures of the two headline kernels.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2u.log
: > $L
t0=$(date +%s); timeout 2000 python -m pytest tests -q -m gpu > $O/r2u_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2u_pytest.log)" | tee -a $L
grep -E "^FAILED|^ERROR" $O/r2u_pytest.log | cut -c1-200 | tee -a $L
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $L
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2u_bench_n1_reference.json 2> $O/r2u_bench_ref.err; echo "ref exit $?" | tee -a $L
timeout 600 python bench.py > $O/r2u_bench_n1.json 2> $O/r2u_bench_n1.err; echo "bench exit $?" | tee -a $L
timeout 600 python bench.py --workload folds > $O/r2u_bench_c1_folds.json 2> $O/r2u_bench_folds.err; echo "folds exit $?" | tee -a $L
timeout 600 python bench.py --config 4 > $O/r2u_bench_c4.json 2> $O/r2u_bench_c4.err; echo "c4 exit $?" | tee -a $L
timeout 600 python bench.py --config 3 > $O/r2u_bench_c3.json 2> $O/r2u_bench_c3.err; echo "c3 exit $?" | tee -a $L
timeout 600 python bench.py --config f1 > $O/r2u_bench_cf1.json 2> $O/r2u_bench_cf1.err; echo "f1 exit $?" | tee -a $L
for f in n1_reference n1 c1_folds c4 c3 cf1; do cut -c1-230 $O/r2u_bench_$f.json | tee -a $L; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2u_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2u_launches_bench.log 2>&1
echo "launch list exit $?" | tee -a $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave|horizon_kernel" -s 8 -c 2 -f -o $O/r2u_full_torus python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2u_ncu_torus.log 2>&1
echo "ncu exit $?" | tee -a $L
if [ "${1:-}" = "c5" ]; then
  t0=$(date +%s); timeout 1500 python bench.py --config 5 > $O/r2u_bench_c5.json 2> $O/r2u_bench_c5.err; echo "config 5 exit $? after $(( $(date +%s) - t0 )) s" | tee -a $L
  cut -c1-230 $O/r2u_bench_c5.json | tee -a $L
fi
