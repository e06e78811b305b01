#!/bin/bash
# Round-2 session Q (N GPUs, default 1): whole GPU test-suite (no -x), default bench line, configs 2 / 3 / f1 / 5, and with N > 1 the
# torchrun bench of the driver's SCALE launch.
set -u
cd "$(dirname "$0")/.."
N=${1:-1}
O=gpurun_out
mkdir -p $O
L=$O/r2q_n$N.log
: > $L
if [ "$N" = "1" ]; then
  echo "== pytest -m gpu (whole suite)" | tee -a $L
  t0=$(date +%s); timeout 2000 python -m pytest tests -q -m gpu > $O/r2q_pytest.log 2>&1; echo "pytest exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2q_pytest.log)" | tee -a $L
  grep -E "^FAILED|^ERROR" $O/r2q_pytest.log | cut -c1-200 | tee -a $L
  python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a $L
  for c in 1 2 3 f1 5; do
    t0=$(date +%s); timeout 1500 python bench.py --config $c > $O/r2q_bench_c$c.json 2> $O/r2q_bench_c$c.err; echo "config $c exit $? after $(( $(date +%s) - t0 )) s" | tee -a $L
    cut -c1-230 $O/r2q_bench_c$c.json | tee -a $L
  done
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2q_n${N}_torchrun.json 2> $O/r2q_n${N}_torchrun.err
  echo "exit $?" | tee -a $L; cut -c1-400 $O/r2q_n${N}_torchrun.json | tee -a $L
  python bench.py --gpus $N --single-process --gather p2p --steps 10 --warmup 3 > $O/r2q_n${N}_group.json 2> $O/r2q_n${N}_group.err
  echo "exit $?" | tee -a $L; cut -c1-400 $O/r2q_n${N}_group.json | tee -a $L
  timeout 600 python -m pytest tests/test_gpu_group.py -q -m gpu 2>&1 | tail -2 | tee -a $L
fi
