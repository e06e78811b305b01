#!/bin/bash
# Round-2 session I (1 GPU): where does the sorted work list pay?  Full mesh and 2- / 4- / 8-way shards, list on / off.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
: > $O/r2i_sweep.jsonl
for w in 1 2 4 8; do
  echo "{\"world\": $w}" >> $O/r2i_sweep.jsonl
  timeout 300 python tools/sweep.py --reps 5 --flush --world $w --rank 0 work_list=0,1 >> $O/r2i_sweep.jsonl 2>> $O/r2i_sweep.err
done
echo "{\"mesh\": \"folds\"}" >> $O/r2i_sweep.jsonl
timeout 300 python tools/sweep.py --reps 3 --flush --mesh folds work_list=0,1 >> $O/r2i_sweep.jsonl 2>> $O/r2i_sweep.err
cut -c1-150 $O/r2i_sweep.jsonl
