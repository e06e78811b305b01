#!/bin/bash
# Round-2 session E (N GPUs, default 2): the multi-GPU driver behind the C ABI (NCCL and P2P gather on real peers), the torchrun arm,
# the group probe capture.  usage: gpu_session_r2e.sh N
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
O=gpurun_out
mkdir -p $O
L=$O/r2e_n$N.log
: > $L
run() { name=$1; shift; echo "== $name: $*" | tee -a $L; t0=$(date +%s); timeout 1500 "$@" > $O/r2e_n${N}_$name.json 2> $O/r2e_n${N}_$name.err; echo "exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2e_n${N}_$name.err | cut -c1-200)" | tee -a $L; cut -c1-200 $O/r2e_n${N}_$name.json | tee -a $L; }
nvidia-smi topo -m > $O/r2e_n${N}_topo.txt 2>&1
echo "== pytest group" | tee -a $L
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_shim.py -q -m gpu -x > $O/r2e_n${N}_pytest.log 2>&1; echo "pytest exit $?: $(tail -1 $O/r2e_n${N}_pytest.log)" | tee -a $L
run torchrun python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3
run group_p2p python bench.py --gpus $N --single-process --gather p2p --steps 10
run group_nccl python bench.py --gpus $N --single-process --gather nccl --steps 10
run probe python tools/group_probe.py --gpus $N --check
