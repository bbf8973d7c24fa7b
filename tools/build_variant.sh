#!/bin/bash
# build a variant of the library with extra nvcc flags for mc_fused.cu only: tools/build_variant.sh <name> <flags...>
# -> build/variants/libbfb_<name>.so (use with BAYESFORMER_B200_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../bayesformer_b200/csrc"
mkdir -p ../../build/variants
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr "$@" -dc -o ../../build/variants/mc_fused_$name.o mc_fused.cu
$NVCC $ARCH -shared -o ../../build/variants/libbfb_$name.so runtime.o acquire.o topk.o transformer_layers.o transformer.o umma_selftest.o ../../build/variants/mc_fused_$name.o gemm_tc.o attention_tc.o train_fused.o -lcudart
echo built build/variants/libbfb_$name.so
