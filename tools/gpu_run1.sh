#!/bin/bash
# parity tests with the main build, then stage timings of every variant build
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== main"; timeout 300 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1
bash tools/gpu_variants.sh
for wl in synth_16k; do timeout 120 python tools/prof_frame.py $wl 4 2>&1 | tail -1; done
