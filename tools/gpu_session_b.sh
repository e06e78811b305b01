#!/bin/bash
# GPU session: parity + timing of the horizon-map variants (prt_b200/csrc/variants/*.so).  Output in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
V=$PWD/prt_b200/csrc/variants
: > $O/b.jsonl
: > $O/b_session.log
for name in "$@"; do
  lib=$V/$name.so
  [ "$name" = tree ] && lib=$PWD/prt_b200/csrc/libprt_b200.so
  echo "== $name" | tee -a $O/b_session.log
  if [ -z "${SKIP_PARITY:-}" ]; then
    PRT_B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/b_parity_$name.log 2>&1
    echo "parity exit $?: $(tail -1 $O/b_parity_$name.log)" | tee -a $O/b_session.log
  fi
  echo "{\"lib\": \"$name\"}" >> $O/b.jsonl
  PRT_B200_LIB=$lib timeout 400 python tools/sweep.py --reps 4 ${SWEEP-horizon_near=30} >> $O/b.jsonl 2>> $O/b.err
done
cut -c1-250 $O/b.jsonl | tee -a $O/b_session.log
