#!/bin/bash
# tools/build_variant.sh NAME "-DFLAG=.. ..."  -- builds prt_b200/csrc/variants/NAME.so with extra compile flags (tuning
# experiments; select it with PRT_B200_LIB=...).  Objects go to /tmp, the tree's own library is untouched.
set -e
name=$1; flags=$2
src=$(cd "$(dirname "$0")/../prt_b200/csrc" && pwd)
obj=/tmp/prt_variant_$name
mkdir -p $obj $src/variants
pids=()
for f in abi bake bake_wave bake_inter horizon env probe volume gi raytrace cache group; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-O2 $flags -c $src/$f.cu -o $obj/$f.o 2> $obj/$f.log &
  pids+=($!)
done
g++ -O2 -std=c++17 -fPIC -c $src/bvh_build.cpp -o $obj/bvh_build.o
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $src/variants/$name.so $obj/*.o -lpthread -ldl
ls -la $src/variants/$name.so
