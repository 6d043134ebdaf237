#!/bin/bash
# band mode on one GPU: the GPU test suite, the stage times of single bands of the 16K frame, the full 16K frame
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
for b in 0 3; do echo "band $b/8:"; SLPR_BAND=$b/8 timeout 300 python tools/prof_frame.py synth_16k 4 2>&1 | tail -1; done
timeout 120 python tools/lat_frame.py synth_16k 30 2>&1 | tail -1
timeout 120 python tools/lat_frame.py synth_1m_4k 30 2>&1 | tail -1
timeout 300 python tools/prof_frame.py synth_16k 4 2>&1 | tail -1
