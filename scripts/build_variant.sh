#!/bin/bash
# usage: bash scripts/build_variant.sh <suffix> <extra nvcc flags...>
# Builds tagdust_b200/libtagdust_b200_<suffix>.so from the same sources with extra -D flags (A/B runs with TDG_LIB,
# scripts/gpu_ab.sh).  Objects go to a scratch directory; the product library is untouched.
set -e
cd "$(dirname "$0")/.."
SUF=$1; shift
TMP=$(mktemp -d)
SRC=tagdust_b200/csrc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O2,-Wall,-ffp-contract=off $*"
nvcc $FLAGS -Xptxas -v -c $SRC/tdg_kernels.cu -o $TMP/k.o 2> $TMP/ptxas.txt || { cat $TMP/ptxas.txt; exit 1; }
nvcc $FLAGS -c $SRC/tdg_host.cu -o $TMP/h.o
nvcc $FLAGS -x cu -c $SRC/tdg_arch.cpp -o $TMP/a.o
nvcc $FLAGS -x cu -c $SRC/tdg_stream.cpp -o $TMP/s.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tagdust_b200/libtagdust_b200_$SUF.so $TMP/k.o $TMP/h.o $TMP/a.o $TMP/s.o -cudart static -lpthread
grep -A2 "k_forward\|k_backward" $TMP/ptxas.txt | grep -i "registers\|spill" | head -8
rm -rf $TMP
echo built tagdust_b200/libtagdust_b200_$SUF.so
