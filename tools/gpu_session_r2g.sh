#!/bin/bash
# Round-2 session G (1 GPU): full ncu captures with source of the two headline kernels on both workloads, launch list, knob sweeps
# (horizon refinement on the folded mesh, L2 prefetch on an 8-way shard, node-test variants), config 5 at N = 1.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2g.log
: > $L
echo "== pytest" | tee -a $L
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_baseline_sizes.py > $O/r2g_pytest.log 2>&1; echo "pytest exit $?: $(tail -1 $O/r2g_pytest.log)" | tee -a $L
timeout 300 python tools/group_probe.py --gpus 1 --check > $O/r2g_probe_n1.json 2>> $O/r2g_sweep.err; cut -c1-400 $O/r2g_probe_n1.json | tee -a $L
echo "== launch list" | tee -a $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2g_launches_bench.log 2>&1
echo "exit $?" | tee -a $L
for wl in torus folds; do
  echo "== ncu full $wl" | tee -a $L
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave|horizon_kernel" -s 4 -c 2 -f -o $O/r2g_full_$wl python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2g_ncu_$wl.log 2>&1
  echo "exit $?" | tee -a $L
done
echo "== sweeps" | tee -a $L
: > $O/r2g_sweep.jsonl
sw() { echo "{\"sweep\": \"$*\"}" >> $O/r2g_sweep.jsonl; timeout 600 python tools/sweep.py --reps 3 --flush "$@" >> $O/r2g_sweep.jsonl 2>> $O/r2g_sweep.err; }
sw --mesh folds horizon_near=30,22,16 horizon_budget=64,160
sw --mesh torus horizon_near=30,24 horizon_budget=64,128
sw --mesh torus --world 8 --rank 0 l2_prefetch=0,1
sw --mesh torus --world 8 --rank 3 l2_prefetch=0,1 work_list=0,1
for v in prmt1 prmt2 prefetch prefetch_prmt2; do
  echo "{\"lib\": \"$v\"}" >> $O/r2g_sweep.jsonl
  PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 4 --flush horizon_near=30 >> $O/r2g_sweep.jsonl 2>> $O/r2g_sweep.err
  PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1 | tee -a $L
done
echo "{\"lib\": \"tree\"}" >> $O/r2g_sweep.jsonl
timeout 300 python tools/sweep.py --reps 4 --flush horizon_near=30 >> $O/r2g_sweep.jsonl 2>> $O/r2g_sweep.err
cut -c1-230 $O/r2g_sweep.jsonl | tee -a $L
echo "== config 5, N = 1" | tee -a $L
t0=$(date +%s); timeout 1500 python bench.py --config 5 > $O/r2g_c5_n1.json 2> $O/r2g_c5_n1.err; echo "exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2g_c5_n1.err | cut -c1-200)" | tee -a $L
cut -c1-300 $O/r2g_c5_n1.json | tee -a $L
