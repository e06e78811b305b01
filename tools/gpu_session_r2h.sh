#!/bin/bash
# Round-2 session H (1 GPU): full GPU suite on the tree with the guarded byte-permute node test and the sorted work list; A/B of the
# plane-conversion split and of the work-list granularity (on an 8-way shard); configs 3 / 4 with the shared node test; probe capture.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=$O/r2h.log
: > $L
echo "== pytest" | tee -a $L
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x > $O/r2h_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2h_pytest.log)" | tee -a $L
: > $O/r2h_sweep.jsonl
for v in tree p0 p2 p4 tree p0; do
  lib=$PWD/prt_b200/csrc/variants/$v.so; [ "$v" = tree ] && lib=$PWD/prt_b200/csrc/libprt_b200.so
  echo "{\"lib\": \"$v\"}" >> $O/r2h_sweep.jsonl
  PRT_B200_LIB=$lib timeout 300 python tools/sweep.py --reps 4 --flush horizon_near=30 >> $O/r2h_sweep.jsonl 2>> $O/r2h_sweep.err
done
for v in tree wb4 wb256 tree wb4 wb256; do
  lib=$PWD/prt_b200/csrc/variants/$v.so; [ "$v" = tree ] && lib=$PWD/prt_b200/csrc/libprt_b200.so
  echo "{\"lib\": \"$v (8-way shard, rank 3)\"}" >> $O/r2h_sweep.jsonl
  PRT_B200_LIB=$lib timeout 300 python tools/sweep.py --reps 6 --flush --world 8 --rank 3 work_list=0,1 >> $O/r2h_sweep.jsonl 2>> $O/r2h_sweep.err
done
cut -c1-200 $O/r2h_sweep.jsonl | tee -a $L
run() { name=$1; shift; echo "== $name: $*" | tee -a $L; t0=$(date +%s); timeout 1500 "$@" > $O/r2h_$name.json 2> $O/r2h_$name.err; echo "exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2h_$name.err | cut -c1-200)" | tee -a $L; cut -c1-200 $O/r2h_$name.json | tee -a $L; }
run c1 python bench.py
run c3 python bench.py --config 3 --no-cpu-baseline
run c4 python bench.py --config 4 --no-cpu-baseline
run probe python tools/group_probe.py --gpus 1 --check
