#!/bin/bash
# profiles for the round: (1) ncu launch list of the bench command, (2) --set full capture of every
# kernel of one frame (segmented sort, the default), (3) the same for the radix sort kernels
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
   --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-radix-leg > gpurun_out/launches_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_" -s 20 -c 15 \
   -f -o gpurun_out/prof_$TAG python tools/prof_frame.py synth_1m_4k 2 > gpurun_out/prof_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_onesweep|k_radix" -s 7 -c 7 \
   -f -o gpurun_out/prof_${TAG}_radix python tools/prof_frame.py synth_1m_4k 2 12 > gpurun_out/prof_${TAG}_radix.log 2>&1
tail -2 gpurun_out/launches_$TAG.log | cut -c1-300; tail -2 gpurun_out/prof_$TAG.log; tail -2 gpurun_out/prof_${TAG}_radix.log; ls -la gpurun_out | tail -5
