#!/bin/bash
# (historical record of the session: the slab filter and its knob wave_filter were removed afterwards -- see DESIGN.md section 4)
# Round-2 session N (1 GPU): why is the slab filter of the traversal pass slow on the GPU?  One full ncu capture (with source) of the
# filter variant and of the plain traversal on the same horizon settings, more A/B points without the filter, inline-filter variant.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2n.log
: > $L
: > $O/r2n_sweep.jsonl
sw() { echo "{\"sweep\": \"$*\"}" >> $O/r2n_sweep.jsonl; timeout 600 python tools/sweep.py --reps 3 --flush "$@" >> $O/r2n_sweep.jsonl 2>> $O/r2n_sweep.err; }
sw --mesh torus horizon_near=157 horizon_mid=12,16,24 horizon_gain=64,128 wave_filter=0
sw --mesh torus horizon_near=30 horizon_mid=16 horizon_gain=64,160 wave_filter=0
sw --mesh torus horizon_near=45 horizon_mid=0 wave_filter=0
sw --mesh folds horizon_near=30 horizon_mid=0,16 horizon_gain=128 wave_filter=0
sw --mesh folds horizon_near=157 horizon_mid=24 horizon_gain=128,320 wave_filter=0
echo "{\"lib\": \"filtinline\"}" >> $O/r2n_sweep.jsonl
PRT_B200_LIB=$PWD/prt_b200/csrc/variants/filtinline.so timeout 300 python tools/sweep.py --reps 3 --flush wave_filter=1 >> $O/r2n_sweep.jsonl 2>> $O/r2n_sweep.err
cut -c1-330 $O/r2n_sweep.jsonl | tee -a $L
for f in 1 0; do
  echo "== ncu full, wave_filter=$f" | tee -a $L
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bake_wave" -s 4 -c 1 -f -o $O/r2n_full_f$f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ncu --tune wave_filter=$f > $O/r2n_ncu_f$f.log 2>&1
  echo "exit $?" | tee -a $L
done
tail -3 $O/r2n_sweep.err | cut -c1-300 | tee -a $L
