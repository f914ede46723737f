#!/bin/bash
# Kernel-variant experiments: rebuilds some fiber translation units with extra -D flags and links them with the objects of the
# last full build into dmft-ed_b200/build/variants/libedgpu_<name>.so (load it with EDGPU_LIB_PATH=...).
# usage: scripts/build_variant.sh <name> "<units, e.g. fib_nl8 fib_nl8h>" "<flags>"
set -e
cd "$(dirname "$0")/../dmft-ed_b200/csrc"
name=$1; units=$2; flags=$3
mkdir -p ../build/variants
objs=$(ls ../build/*.o)
pids=""
for u in $units; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC $flags -x cu -c $u.cu -o ../build/variants/${u}_$name.o &
  pids="$pids $!"
  objs=$(echo "$objs" | grep -v "/$u.o")
  objs="$objs ../build/variants/${u}_$name.o"
done
for p in $pids; do wait $p; done
nvcc -shared -o ../build/variants/libedgpu_$name.so $objs -lcudart -ldl
echo "built build/variants/libedgpu_$name.so"
