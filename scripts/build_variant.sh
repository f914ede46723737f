#!/bin/bash
# Kernel-variant experiments: rebuilds ONE fiber translation unit with extra -D flags and links it with the objects of the last
# full build into dmft-ed_b200/build/variants/libedgpu_<name>.so (load it with EDGPU_LIB_PATH=...).
# usage: scripts/build_variant.sh <name> <nl> "<flags>"
set -e
cd "$(dirname "$0")/../dmft-ed_b200/csrc"
name=$1; nl=$2; flags=$3
mkdir -p ../build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC $flags -x cu -c fib_nl$nl.cu -o ../build/variants/fib_nl${nl}_$name.o
objs=$(ls ../build/*.o | grep -v "fib_nl$nl.o")
nvcc -shared -o ../build/variants/libedgpu_$name.so $objs ../build/variants/fib_nl${nl}_$name.o -lcudart -ldl
echo "built build/variants/libedgpu_$name.so"
