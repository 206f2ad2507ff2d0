#!/bin/bash
# Build an alternative libbee2_b200 with extra nvcc flags (kernel-variant experiments):
#   tools/build_variant.sh NAME "-DBIGN_THREADS=64 -DBIGN_MIN_BLOCKS=7"
# -> gpurun_scratch/NAME.so, selected at run time with BEE2_B200_LIB=gpurun_scratch/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
extra="$*"
src=bee2_b200/csrc
obj=gpurun_scratch/obj_$name
mkdir -p $obj
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=default -diag-suppress 20044"
for f in bign; do $NV $extra -c -o $obj/$f.o $src/$f.cu & done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xlinker -Bsymbolic \
  -o gpurun_scratch/$name.so $obj/bign.o $src/build/bash.o $src/build/belt.o $src/build/microbench.o \
  $src/build/belt_dwp.o $src/build/engine.o $src/build/host_bash.o $src/build/host_belt.o $src/build/host_bign.o -lpthread -ldl
echo built gpurun_scratch/$name.so
