#!/bin/bash
# Round-end captures on one B200 (run under gpurun): the bench line, the ncu launch list over bench.py and one
# ncu --set full capture of a direct-launch frame; tools/summarize_profile.py <tag> turns them into profiles/<tag>_*.
# usage: bash tools/capture_profiles.sh <tag>
tag=${1:-final}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-radix-leg --no-scenes > gpurun_out/launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_ -c 40 -f -o gpurun_out/prof_${tag} \
    python tools/prof_frame.py synth_1m_4k 2 > gpurun_out/prof_${tag}.log 2>&1
tail -c 600 gpurun_out/${tag}_bench.err; tail -n 3 gpurun_out/launches_${tag}.log; tail -n 3 gpurun_out/prof_${tag}.log
