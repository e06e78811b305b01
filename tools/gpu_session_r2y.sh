#!/bin/bash
# Round-2 session Y (1 GPU): admission threshold of the traversal pass (absolute: 32 / 48 / 64 = tree / 96), scan unroll 4, work list off
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r2y_sweep.jsonl
for v in tree room32 room48 room96 unroll4 tree; do
  echo "{\"lib\": \"$v\"}" >> $O/r2y_sweep.jsonl
  if [ "$v" = "tree" ]; then timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 work_list=-1,0 >> $O/r2y_sweep.jsonl 2>> $O/r2y_sweep.err
  else PRT_B200_LIB=$PWD/prt_b200/csrc/variants/$v.so timeout 300 python tools/sweep.py --reps 4 --flush horizon_mid=24 >> $O/r2y_sweep.jsonl 2>> $O/r2y_sweep.err; fi
done
cut -c1-200 $O/r2y_sweep.jsonl
