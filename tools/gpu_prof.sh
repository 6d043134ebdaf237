#!/bin/bash
# profiles for the round: (1) ncu launch list of the bench command, (2) --set full capture of the hot kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:"k_onesweep|k_walk|k_spans|k_fill_cells|k_radix_hist|k_resolve|k_piece_emit|k_lookback_scan|k_wsum|k_transform|k_monotonize" -s 14 -c 16 \
   -f -o gpurun_out/prof_$TAG python tools/prof_frame.py synth_1m_4k 2 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/launches_$TAG.log | cut -c1-300; tail -2 gpurun_out/prof_$TAG.log; ls -la gpurun_out | tail -5
