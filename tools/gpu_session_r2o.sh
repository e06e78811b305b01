#!/bin/bash
# (historical record of the session: the slab filter and its knob wave_filter were removed afterwards -- see DESIGN.md section 4)
# Round-2 session O (1 GPU): the slab filter with the stack capacities fixed (node stack 320, leaf stack 128, ready queue 64; ready items
# count in the admission rule).  A/B against the plain traversal on both workloads, library variants, parity tests, ncu of the filter kernel.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2o.log
: > $L
: > $O/r2o_sweep.jsonl
sw() { echo "{\"sweep\": \"$*\"}" >> $O/r2o_sweep.jsonl; timeout 600 python tools/sweep.py --reps 3 --flush "$@" >> $O/r2o_sweep.jsonl 2>> $O/r2o_sweep.err; }
sw --mesh torus horizon_mid=12,16,24 horizon_gain=64 wave_filter=1,0
sw --mesh torus horizon_near=30 horizon_mid=0 horizon_slabs=0 wave_filter=0,1
sw --mesh folds horizon_mid=16,24 horizon_gain=128 wave_filter=1,0
for v in minb6 filtinline room1; do
  echo "{\"lib\": \"$v\"}" >> $O/r2o_sweep.jsonl
  PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 3 --flush horizon_mid=16 wave_filter=1 >> $O/r2o_sweep.jsonl 2>> $O/r2o_sweep.err
done
# config 4 (interreflection, order 4, 64 x 64 samples) on a 131 k-vertex slice: filter on / off
echo "{\"sweep\": \"config 4 slice\"}" >> $O/r2o_sweep.jsonl
timeout 600 python tools/sweep.py --reps 2 --nu 1448 --nv 1448 --order 4 --su 64 --sv 64 --mode 2 --bounces 3 --world 16 --rank 3 wave_filter=1,0 >> $O/r2o_sweep.jsonl 2>> $O/r2o_sweep.err
cut -c1-330 $O/r2o_sweep.jsonl | tee -a $L
echo "== pytest (parity, group, pinned)" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_pinned.py -q -m gpu -x > $O/r2o_pytest.log 2>&1; echo "pytest exit $?: $(tail -1 $O/r2o_pytest.log)" | tee -a $L
echo "== ncu full, filter" | tee -a $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave" -s 4 -c 1 -f -o $O/r2o_full_f1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ncu > $O/r2o_ncu_f1.log 2>&1
echo "exit $?" | tee -a $L
tail -3 $O/r2o_sweep.err | cut -c1-300 | tee -a $L
