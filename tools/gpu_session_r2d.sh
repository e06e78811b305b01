#!/bin/bash
# Round-2 session D: the new bench.py on one GPU -- every config, both arms of the default -- plus the GPU test suite.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
: > $O/r2d.log
run() { name=$1; shift; echo "== $name: $*" | tee -a $O/r2d.log; t0=$(date +%s); timeout 1500 "$@" > $O/r2d_$name.json 2> $O/r2d_$name.err; echo "exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2d_$name.err | cut -c1-200)" | tee -a $O/r2d.log; cut -c1-200 $O/r2d_$name.json | tee -a $O/r2d.log; }
run c1 python bench.py --steps 5 --warmup 3
run c1_folds python bench.py --steps 5 --warmup 3 --workload folds
run c2 python bench.py --config 2
run c3 python bench.py --config 3
run f1 python bench.py --config f1
run c4 python bench.py --config 4
run c1_ref python bench.py --impl reference --steps 2 --warmup 1
run c1_cube python bench.py --mesh tests/golden/sphere.obj --steps 3
echo "== pytest -m gpu" | tee -a $O/r2d.log
t0=$(date +%s); timeout 1500 python -m pytest tests -q -m gpu -x > $O/r2d_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2d_pytest.log)" | tee -a $O/r2d.log
