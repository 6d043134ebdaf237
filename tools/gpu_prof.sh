#!/bin/bash
# ncu launch list + full capture of the hot kernels for one frame of synth_1m_4k
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 21 -c 40 --csv \
   --log-file gpurun_out/launches_$TAG.csv python tools/prof_frame.py synth_1m_4k 3 > gpurun_out/launches_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:"k_onesweep|k_intersect|k_lookback_scan|k_fill_cells|k_gen_fragment|k_radix_hist|k_resolve|k_walk" -s 15 -c 15 \
   -f -o gpurun_out/prof_$TAG python tools/prof_frame.py synth_1m_4k 2 > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/launches_$TAG.log; tail -3 gpurun_out/prof_$TAG.log; ls -la gpurun_out
