#!/bin/bash
# Round-2 session T (1 GPU): packed FFMA2 in the node test (near plane by I2F + far plane by byte permute of an axis in one instruction)
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r2t_sweep.jsonl
for v in ffma2 tree ffma2 tree; do
  echo "{\"lib\": \"$v\"}" >> $O/r2t_sweep.jsonl
  if [ "$v" = "tree" ]; then timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2t_sweep.jsonl 2>> $O/r2t_sweep.err
  else PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2t_sweep.jsonl 2>> $O/r2t_sweep.err; fi
done
echo "{\"lib\": \"ffma2 folds\"}" >> $O/r2t_sweep.jsonl
PRT_B200_LIB=$PWD/prt_b200/csrc/variants/ffma2.so timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds horizon_mid=24 >> $O/r2t_sweep.jsonl 2>> $O/r2t_sweep.err
echo "{\"lib\": \"tree folds\"}" >> $O/r2t_sweep.jsonl
timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds horizon_mid=24 >> $O/r2t_sweep.jsonl 2>> $O/r2t_sweep.err
PRT_B200_LIB=$PWD/prt_b200/csrc/variants/ffma2.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
cut -c1-300 $O/r2t_sweep.jsonl
