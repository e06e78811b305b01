#!/bin/bash
# Round-2 session Z (1 GPU): fourth slab axis in the node test of the traversal pass (Dop32), A/B by the knob wave_dop
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r2z_sweep.jsonl
sw() { echo "{\"sweep\": \"$*\"}" >> $O/r2z_sweep.jsonl; timeout 600 python tools/sweep.py --reps 4 --flush "$@" >> $O/r2z_sweep.jsonl 2>> $O/r2z_sweep.err; }
sw --mesh torus wave_dop=1,0
sw --mesh folds wave_dop=1,0
sw --mesh torus --world 8 --rank 3 wave_dop=1,0
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pinned.py tests/test_gpu_group.py tests/test_gpu_baseline_sizes.py -x -q -m gpu 2>&1 | tail -1
cut -c1-260 $O/r2z_sweep.jsonl
