#!/bin/bash
# Round-2 session V (1 GPU): packed FFMA2 / FMUL2 in to_world and the triangle test, FMNMX clamp in rcp_box, stack capacities 320 / 192
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r2v_sweep.jsonl
for v in tree packed packed_rcp packed_rcp_caps tree packed_rcp_caps; do
  echo "{\"lib\": \"$v\"}" >> $O/r2v_sweep.jsonl
  if [ "$v" = "tree" ]; then timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2v_sweep.jsonl 2>> $O/r2v_sweep.err
  else PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2v_sweep.jsonl 2>> $O/r2v_sweep.err; fi
done
for v in packed_rcp_caps tree; do
  echo "{\"lib\": \"$v folds\"}" >> $O/r2v_sweep.jsonl
  if [ "$v" = "tree" ]; then timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds horizon_mid=24 >> $O/r2v_sweep.jsonl 2>> $O/r2v_sweep.err
  else PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds horizon_mid=24 >> $O/r2v_sweep.jsonl 2>> $O/r2v_sweep.err; fi
done
PRT_B200_LIB=$PWD/prt_b200/csrc/variants/packed_rcp_caps.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pinned.py tests/test_gpu_probe.py tests/test_gpu_raytrace.py tests/test_gpu_gi.py -x -q -m gpu 2>&1 | tail -1
cut -c1-300 $O/r2v_sweep.jsonl
