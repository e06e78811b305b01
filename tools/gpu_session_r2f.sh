#!/bin/bash
# Round-2 session F (8 GPUs): the default workload under torchrun and through the C-ABI group driver, the group probe capture, and
# BASELINE config 5 (20 M vertices, order 5, 8192 samples) through the group driver with both gather modes.
set -u
cd "$(dirname "$0")/.."
N=${1:-8}
O=gpurun_out
mkdir -p $O
L=$O/r2f_n$N.log
: > $L
run() { name=$1; shift; echo "== $name: $*" | tee -a $L; t0=$(date +%s); timeout 1500 "$@" > $O/r2f_n${N}_$name.json 2> $O/r2f_n${N}_$name.err; echo "exit $? after $(( $(date +%s) - t0 )) s: $(tail -1 $O/r2f_n${N}_$name.err | cut -c1-200)" | tee -a $L; cut -c1-200 $O/r2f_n${N}_$name.json | tee -a $L; }
run torchrun python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5
run torchrun_nccl python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --gather nccl --no-ncu
run group_p2p python bench.py --gpus $N --single-process --gather p2p --steps 20 --warmup 5
run probe python tools/group_probe.py --gpus $N --check
if [ "${C5:-1}" = 1 ]; then run c5_group_p2p python bench.py --config 5 --gpus $N --single-process --gather p2p; fi
