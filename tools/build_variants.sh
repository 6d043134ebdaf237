#!/bin/bash
# build onesweep tuning variants into vkscanlinepr_b200/variants/ (git-ignored .so files)
cd /root/repo/vkscanlinepr_b200/csrc
mkdir -p ../variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v -shared"
build() { # tag threads items blocks late
  $NV -DSLPR_RS_THREADS=$2 -DSLPR_RS_ITEMS=$3 -DSLPR_RS_BLOCKS=$4 -DSLPR_RS_LATE_VALUES=$5 $6 -o ../variants/libslpr_$1.so slpr.cu host_scene.cpp 2> /tmp/build_$1.log || { echo "build $1 failed"; tail -5 /tmp/build_$1.log; }
  grep -A2 "k_onesweep" /tmp/build_$1.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1"
}
build A 384 16 2 0 &
build B 384 16 2 1 &
build C 256 12 4 1 &
build D 256 16 3 1 &
wait
build E 512 8 2 1 &
build F 512 12 2 1 &
build G 256 8 5 1 &
build H 384 12 3 1 &
wait
