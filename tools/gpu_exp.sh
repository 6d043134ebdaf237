#!/bin/bash
# one GPU call: micro-benchmarks, the full GPU test suite on the main build, then parity + timings of every
# library variant under vkscanlinepr_b200/variants/ (tools/build_variants.sh)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for m in tools/micro/*; do [ -x "$m" ] && [ ! -d "$m" ] && { echo "=== $m"; timeout 120 "$m"; }; done
echo "=== pytest gpu (main build)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in $(ls vkscanlinepr_b200/variants | sed "s/libslpr_//;s/.so//"); do
  export SLPR_LIB=$PWD/vkscanlinepr_b200/variants/libslpr_$v.so
  echo "=== variant $v"
  [ "$v" != base ] && timeout 900 python -m pytest tests/test_parity_gpu.py -x -q 2>&1 | tail -3
  timeout 120 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1 | cut -c1-600
  timeout 120 python tools/lat_frame.py synth_1m_4k 50 2>&1 | tail -1
  for wl in ${EXTRA_WL:-tiger@3840x2160 synth_16k}; do timeout 120 python tools/lat_frame.py $wl 20 2>&1 | tail -1; done
done
unset SLPR_LIB
