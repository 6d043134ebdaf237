#!/bin/bash
# Build tuning variants of libslpr.so into vkscanlinepr_b200/variants/ (git-ignored; they travel to the GPU box):
#   tools/variants.sh name1 "-DSLPR_X=1 -DSLPR_Y=2" name2 "..."
# Run one with SLPR_LIB=vkscanlinepr_b200/variants/libslpr_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/../vkscanlinepr_b200/csrc"
mkdir -p ../variants
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ \
    -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $defs -shared -o ../variants/libslpr_$name.so slpr.cu host_scene.cpp 2> ../variants/build_$name.log &
done
wait
ls -la ../variants/*.so
