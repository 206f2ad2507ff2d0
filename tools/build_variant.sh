#!/bin/bash
# Build an alternative libbee2_b200 with extra nvcc flags for ONE kernel file (variant experiments):
#   tools/build_variant.sh NAME FILE "-DBIGN_THREADS=64 -DBIGN_MIN_BLOCKS=7"
# -> gpurun_scratch/NAME.so, selected at run time with BEE2_B200_LIB=$PWD/gpurun_scratch/NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; file=$2; shift 2
extra="$*"
src=bee2_b200/csrc
obj=gpurun_scratch/obj_$name
mkdir -p $obj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
  -Xcompiler -fvisibility=default -diag-suppress 20044 $extra -c -o $obj/$file.o $src/$file.cu
objs=""
for f in bash belt bign bign_lowocc microbench belt_dwp engine host_bash host_belt host_bign; do
  if [ $f = $file ]; then objs="$objs $obj/$f.o"; else objs="$objs $src/build/$f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -Xlinker -Bsymbolic \
  -o gpurun_scratch/$name.so $objs -lpthread -ldl
echo built gpurun_scratch/$name.so
