#!/bin/bash
# quick GPU iteration: parity tests + stage timings of the bench workload
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "=== stages (segmented default)"; timeout 300 python tools/prof_frame.py synth_1m_4k 6 2>&1 | tail -2
echo "=== stages (radix forced)"; timeout 300 python tools/prof_frame.py synth_1m_4k 6 12 2>&1 | tail -2
for wl in tiger@3840x2160 reschart@1920x1080 synth_16k; do timeout 120 python tools/prof_frame.py $wl 4 2>&1 | tail -1; done
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
