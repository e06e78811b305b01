#!/bin/bash
# BASELINE config 5 through the C-ABI group driver on N GPUs
set -u
cd "$(dirname "$0")/.."
N=${1:-4}
python bench.py --config 5 --gpus $N --single-process --gather p2p > gpurun_out/r2l_c5_n$N.json 2> gpurun_out/r2l_c5_n$N.err
echo "exit $?"; cut -c1-250 gpurun_out/r2l_c5_n$N.json
