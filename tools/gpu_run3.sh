#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "=== stages"; timeout 300 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1
for wl in synth_1m_4k tiger@3840x2160 reschart@1920x1080 synth_16k; do timeout 120 python tools/lat_frame.py $wl 30 2>&1 | tail -1; done
