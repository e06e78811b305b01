#!/bin/bash
# Round-2 session W (1 GPU): per-probe entry list in the capture kernel -- parity tests, config 3 with the list on / off
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_probe.py tests/test_gpu_group.py tests/test_gpu_gi.py tests/test_gpu_pinned.py -q -m gpu 2>&1 | tail -3
for e in 1 0 1 0; do
  timeout 600 python bench.py --config 3 --no-cpu-baseline --no-ncu --tune entry_list=$e > $O/r2w_c3_el$e.json 2> $O/r2w_c3_el$e.err
  echo "entry_list=$e: $(python -c "import json;d=json.load(open('$O/r2w_c3_el$e.json'));print(d['value'], d['ms_per_step'], d.get('results_ok'))")"
done
timeout 600 python bench.py --config 3 > $O/r2w_bench_c3.json 2> $O/r2w_bench_c3.err; cut -c1-300 $O/r2w_bench_c3.json
timeout 300 python tools/group_probe.py --gpus 1 --check > $O/r2w_probe_group_n1.json 2>> $O/r2w.err; cut -c1-400 $O/r2w_probe_group_n1.json
