#!/bin/bash
# stage timings of every variant build only (no parity tests: variants may be timing experiments)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
echo "=== main"; timeout 300 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -1
bash tools/gpu_variants.sh
