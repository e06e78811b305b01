#!/bin/bash
# N-GPU default bench under torchrun (the driver's SCALE launch), both arms of the gather
set -u
cd "$(dirname "$0")/.."
N=${1:-4}
O=gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2k_n${N}_torchrun.json 2> $O/r2k_n${N}_torchrun.err
echo "exit $?"; cut -c1-250 $O/r2k_n${N}_torchrun.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > $O/r2k_n${N}_reference.json 2> $O/r2k_n${N}_reference.err
echo "exit $?"; cut -c1-250 $O/r2k_n${N}_reference.json
