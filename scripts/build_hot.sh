#!/bin/bash
# experiment build: only the hot-shape kernel instances (fast to compile) -> orion-sdr_b200/variants/liborion_b200_hot.so
set -e
cd "$(dirname "$0")/../orion-sdr_b200"
mkdir -p variants build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -cudart static -DORION_ONLY_HOT_SHAPES "$@" -Xptxas -v -c csrc/chain_kernels.cu -o build/ck_hot.o 2>&1 | grep -E "error|spill|registers" | grep -A1 -B1 "Li1ELi8ELi1ELi1ELi100\|Li1ELi8ELi1ELi1ELi0" | head -12 || true
[ build/orion_b200_api.o -nt csrc/orion_b200_api.cu ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -cudart static -c csrc/orion_b200_api.cu -o build/orion_b200_api.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o variants/liborion_b200_hot.so build/ck_hot.o build/orion_b200_api.o -lpthread -ldl -lrt
