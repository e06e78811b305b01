#!/bin/bash
# Round-2 session A: sanity of the inherited tree (GPU tests, bench line), the PRT_WAVE_CULL variant against the tree,
# compute-sanitizer (memcheck + racecheck) over one small call of every shipped kernel family.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu" | tee $O/r2a.log
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2a_pytest.log 2>&1
echo "pytest exit $?: $(tail -1 $O/r2a_pytest.log)" | tee -a $O/r2a.log
echo "== bench" | tee -a $O/r2a.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err
echo "bench exit $?" | tee -a $O/r2a.log
cut -c1-260 $O/r2a_bench.json | tee -a $O/r2a.log
echo "== cull A/B" | tee -a $O/r2a.log
SKIP_PARITY= SWEEP="--flush horizon_near=30" bash tools/gpu_session_b.sh tree cull tree cull > /dev/null 2>&1
cat $O/b_session.log | tee -a $O/r2a.log
echo "== sanitizer" | tee -a $O/r2a.log
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $O/r2a_sanitizer_$tool.txt 2>&1
  echo "$tool exit $?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/r2a_sanitizer_$tool.txt | tail -1)" | tee -a $O/r2a.log
done
