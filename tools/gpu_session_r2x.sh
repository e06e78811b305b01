#!/bin/bash
# Round-2 session X (1 GPU): packed x / y slab test in the candidate scan (entry-list layout: half extents x, y as an aligned pair)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r2x_sweep.jsonl
for v in tree scan0 tree scan0; do
  echo "{\"lib\": \"$v\"}" >> $O/r2x_sweep.jsonl
  if [ "$v" = "tree" ]; then timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2x_sweep.jsonl 2>> $O/r2x_sweep.err
  else PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2x_sweep.jsonl 2>> $O/r2x_sweep.err; fi
done
for v in tree scan0; do
  echo "{\"lib\": \"$v folds\"}" >> $O/r2x_sweep.jsonl
  if [ "$v" = "tree" ]; then timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds horizon_mid=24 >> $O/r2x_sweep.jsonl 2>> $O/r2x_sweep.err
  else PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds horizon_mid=24 >> $O/r2x_sweep.jsonl 2>> $O/r2x_sweep.err; fi
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pinned.py tests/test_gpu_probe.py tests/test_gpu_group.py tests/test_gpu_baseline_sizes.py -x -q -m gpu 2>&1 | tail -1
cut -c1-200 $O/r2x_sweep.jsonl
