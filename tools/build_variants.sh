#!/bin/bash
# build onesweep tuning variants into vkscanlinepr_b200/variants/ (git-ignored .so files)
cd /root/repo/vkscanlinepr_b200/csrc
mkdir -p ../variants; rm -f ../variants/*.so
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v -shared"
build() { # tag "defines"
  $NV $2 -o ../variants/libslpr_$1.so slpr.cu host_scene.cpp 2> /tmp/build_$1.log || { echo "build $1 failed"; tail -5 /tmp/build_$1.log; }
  grep -A2 "${GREP_KERNEL:-k_onesweep}" /tmp/build_$1.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1 $2"
}
i=0
while read -r tag defs; do
  [ -z "$tag" ] && continue
  build $tag "$defs" &
  i=$((i+1)); [ $((i%4)) -eq 0 ] && wait
done < "${1:-/dev/stdin}"
wait
