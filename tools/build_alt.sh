#!/bin/bash
# Build a variant of the library with extra nvcc flags for ONE source file (kernel experiments):
#   tools/build_alt.sh <name> <source.cu> "<flags>"   -> audian_b200/libaudian_b200_<name>.so
set -e
name=$1; src=$2; flags=$3
cd "$(dirname "$0")/../audian_b200"
mkdir -p build/alt_$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -c csrc/$src -o build/alt_$name/${src%.cu}.o
objs=""
for f in api misc minmax sosfilt zerophase sosfwd ingest spectrogram; do
  if [ "$f.cu" == "$src" ]; then objs="$objs build/alt_$name/$f.o"; else objs="$objs build/$f.o"; fi
done
nvcc -shared -o libaudian_b200_$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo audian_b200/libaudian_b200_$name.so
