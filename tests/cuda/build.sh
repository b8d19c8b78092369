#!/bin/bash
# Build the standalone CUDA test binaries (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
CS=../../mlx_sharding_b200/ops/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
mkdir -p bin
nvcc $FLAGS -o bin/gemm_test gemm_test.cu $CS/gemm_tcgen05.cu $CS/gemm_persistent.cu
nvcc $FLAGS -o bin/gemm_sweep gemm_sweep.cu $CS/gemm_tcgen05.cu $CS/gemm_persistent.cu
echo built
