#!/bin/bash
# Round-2 session R (1 GPU): compute-sanitizer (memcheck, racecheck) over one small call of every shipped kernel family, final kernels.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{
  echo "# compute-sanitizer over tools/sanitize_small.py (one small call of every shipped kernel family), round 2, final kernels (oriented slabs + gain rule in the horizon pass)"
  echo "## memcheck"
  timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|sanitize_small|Invalid|error" | head -20
  echo "## racecheck"
  timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | grep -E "COMPUTE-SANITIZER|RACECHECK SUMMARY|sanitize_small|hazard|Race" | head -20
} > $O/r2r_sanitizer.txt 2>&1
cat $O/r2r_sanitizer.txt
