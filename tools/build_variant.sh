#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]   -> build/variants/liboptmc_NAME.so (git-ignored, travels with gpurun)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
    -o build/variants/liboptmc_$name.so optpricing_b200/csrc/optmc.cu
echo built build/variants/liboptmc_$name.so
