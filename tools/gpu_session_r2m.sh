#!/bin/bash
# (historical record of the session: the slab filter and its knob wave_filter were removed afterwards -- see DESIGN.md section 4)
# Round-2 session M (1 GPU): oriented slabs in the horizon pass + slab filter in the traversal pass.  Parity tests of the touched paths,
# A/B sweeps of the new knobs against the previous algorithm on both bench workloads, library variants, one default bench line.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2m.log
: > $L
echo "== pytest (parity, group, pinned, abi, shim)" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_pinned.py tests/test_abi.py tests/test_shim.py -q -m gpu -x > $O/r2m_pytest.log 2>&1; echo "pytest exit $?: $(tail -1 $O/r2m_pytest.log)" | tee -a $L
echo "== sweeps" | tee -a $L
: > $O/r2m_sweep.jsonl
sw() { echo "{\"sweep\": \"$*\"}" >> $O/r2m_sweep.jsonl; timeout 600 python tools/sweep.py --reps 3 --flush "$@" >> $O/r2m_sweep.jsonl 2>> $O/r2m_sweep.err; }
# the previous algorithm, then the new parts one at a time
sw --mesh torus horizon_near=30 horizon_mid=0 horizon_slabs=0 wave_filter=0
sw --mesh torus horizon_near=30 horizon_mid=0 horizon_slabs=0 wave_filter=1
sw --mesh torus horizon_near=30 horizon_mid=0 horizon_slabs=1 wave_filter=0
sw --mesh torus horizon_near=157 horizon_mid=12 horizon_slabs=1 horizon_gain=32,64,96,160 wave_filter=1
sw --mesh torus horizon_near=157 horizon_mid=8,16,24 horizon_slabs=1 horizon_gain=64 wave_filter=1
sw --mesh torus horizon_near=60,90 horizon_mid=12 horizon_slabs=1 horizon_gain=64 wave_filter=1
sw --mesh folds horizon_near=30 horizon_mid=0 horizon_slabs=0 wave_filter=0
sw --mesh folds horizon_near=157 horizon_mid=12 horizon_slabs=1 horizon_gain=32,64,128 wave_filter=1
sw --mesh folds horizon_near=157 horizon_mid=12 horizon_slabs=1 horizon_gain=64 wave_filter=0
for v in minb6 noframe; do
  echo "{\"lib\": \"$v\"}" >> $O/r2m_sweep.jsonl
  PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 3 --flush horizon_gain=64 >> $O/r2m_sweep.jsonl 2>> $O/r2m_sweep.err
done
echo "{\"lib\": \"tree\"}" >> $O/r2m_sweep.jsonl
timeout 300 python tools/sweep.py --reps 3 --flush horizon_gain=64 >> $O/r2m_sweep.jsonl 2>> $O/r2m_sweep.err
cut -c1-330 $O/r2m_sweep.jsonl | tee -a $L
echo "== bench (default)" | tee -a $L
timeout 600 python bench.py > $O/r2m_bench_n1.json 2> $O/r2m_bench_n1.err; echo "bench exit $?" | tee -a $L
cut -c1-400 $O/r2m_bench_n1.json | tee -a $L
tail -3 $O/r2m_sweep.err | cut -c1-300 | tee -a $L
