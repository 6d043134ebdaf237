#!/bin/bash
# final single-GPU evidence of the round: tests, smoke, bench (both arms), launch list + ncu full capture, scene table
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r1v5}
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/gpu_prof.sh $TAG > gpurun_out/prof_${TAG}_run.log 2>&1; tail -c 600 gpurun_out/prof_${TAG}_run.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_$TAG.json; cut -c1-300 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_${TAG}_reference.json; cut -c1-200 gpurun_out/bench_${TAG}_reference.json
timeout 900 python bench.py --mode anim --steps 256 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_${TAG}_anim_synth.json; cut -c1-300 gpurun_out/bench_${TAG}_anim_synth.json
timeout 900 python bench.py --mode anim --workload test@3840x2160 --steps 256 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_${TAG}_anim_test.json; cut -c1-300 gpurun_out/bench_${TAG}_anim_test.json
timeout 600 python tools/bench_scenes.py > gpurun_out/scenes_$TAG.txt 2>&1; tail -17 gpurun_out/scenes_$TAG.txt
timeout 120 python tools/lat_frame.py synth_16k 30 2>&1 | tail -1
timeout 300 python tools/prof_frame.py synth_16k 4 2>&1 | tail -1
