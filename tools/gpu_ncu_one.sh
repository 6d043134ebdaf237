#!/bin/bash
# ncu --set full for one kernel regex, one launch of the 2nd frame. usage: gpu_ncu_one.sh <regex> <tag> [skip] [count]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-1} -c ${4:-1} \
   -f -o gpurun_out/prof_$2 python tools/prof_frame.py ${5:-synth_1m_4k} 2 > gpurun_out/prof_$2.log 2>&1
tail -2 gpurun_out/prof_$2.log
