#!/bin/bash
# usage: tools/build_variant.sh <name> <file.cu> "<extra nvcc flags>"  -> cnn-.../libmnv1_<name>.so with that one object rebuilt
set -e
here=$(cd "$(dirname "$0")/.." && pwd)
csrc=$(echo $here/cnn-*/csrc)
name=$1; src=$2; shift 2
mkdir -p $csrc/build/var
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $csrc/$src -o $csrc/build/var/${name}.o
objs=$(ls $csrc/build/*.o | grep -v "/$(basename $src .cu).o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $csrc/../libmnv1_${name}.so $objs $csrc/build/var/${name}.o
echo built libmnv1_${name}.so
