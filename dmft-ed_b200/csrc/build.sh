#!/bin/bash
# Builds libedgpu.so (sm_100a) in-tree: dmft-ed_b200/libedgpu.so
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --extended-lambda -Xcompiler -fPIC -Xcompiler -Wall"
SRCS="api.cu tables.cu hxv_generic.cu hxv_star.cu hxv_fiber.cu fib_nl8.cu fib_nl8h.cu fib_nl7.cu fib_nl7h.cu fib_nl6.cu fib_nl6h.cu fib_nl5.cu fib_nl5h.cu fib_nl4.cu fib_nl4h.cu fib_nl3.cu fib_nl3h.cu comm.cu eigs.cu lanczos.cu ops.cu csr.cu host/ed_main.cpp host/ed_capi.cpp"
mkdir -p ../build
OBJS=""
PIDS=""
for f in $SRCS; do
  [ -f "$f" ] || continue
  o=../build/$(basename ${f%.*}).o
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ edgpu_internal.h -nt "$o" ] || [ ../../include/edgpu.h -nt "$o" ] || [ star_info.h -nt "$o" ] || [ fiber_common.h -nt "$o" ] || { case "$f" in fib_nl*) [ fiber_kernels.cuh -nt "$o" ];; *) false;; esac; } || [ -f host/ed_host.h -a host/ed_host.h -nt "$o" ] || [ ../../include/ed_b200.h -nt "$o" ]; then
    echo "nvcc $f"
    rm -f "$o"
    $NVCC $FLAGS ${EXTRA_FLAGS} -x cu -c "$f" -o "$o" &
    PIDS="$PIDS $!"
  fi
  OBJS="$OBJS $o"
done
for p in $PIDS; do wait $p || { echo "build.sh: a compile job failed" >&2; exit 1; }; done
$NVCC -shared -o ../libedgpu.so $OBJS -lcudart -ldl
echo "built $(cd ..; pwd)/libedgpu.so"
