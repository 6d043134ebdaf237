#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for v in $(ls vkscanlinepr_b200/variants | sed "s/libslpr_//;s/.so//"); do
  export SLPR_LIB=$PWD/vkscanlinepr_b200/variants/libslpr_$v.so
  echo "=== variant $v"
  timeout 120 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1 | cut -c100-
  for wl in synth_1m_4k synth_16k tiger@3840x2160; do timeout 120 python tools/lat_frame.py $wl 30 2>&1 | tail -1; done
done
