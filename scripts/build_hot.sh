#!/bin/bash
# experiment build: recompiles the hot-shape instances, the launcher and the API with extra -D flags
# -> orion-sdr_b200/variants/liborion_b200_hot.so (the other instance objects come from the last `make -C orion-sdr_b200`)
set -e
cd "$(dirname "$0")/../orion-sdr_b200"
mkdir -p variants build
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -cudart static"
nvcc $F "$@" -Xptxas -v -c csrc/chain_inst_hot.cu -o build/ck_hot.o 2>&1 | grep -E "error|spill|registers" | head -12 || true
nvcc $F "$@" -c csrc/chain_launch.cu -o build/ck_launch.o
nvcc $F "$@" -c csrc/orion_b200_api.cu -o build/ck_api.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o variants/liborion_b200_hot.so build/ck_hot.o build/chain_inst_direct.o build/chain_inst_staged_u1.o build/chain_inst_staged_u2.o build/chain_inst_ws.o build/bank_kernels.o build/agc_kernels.o build/aux_kernels.o build/ck_launch.o build/ck_api.o -lpthread -ldl -lrt
