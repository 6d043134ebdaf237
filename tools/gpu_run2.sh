#!/bin/bash
# full GPU suite + the bench lines of the round (default cfg3, cfg5 animation, a small shipped scene)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-1500
echo "=== bench anim synth_1m_4k"; timeout 900 python bench.py --mode anim --steps 256 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_anim_synth.json | cut -c1-1500
echo "=== bench anim test@4K"; timeout 900 python bench.py --mode anim --workload test@3840x2160 --steps 256 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_anim_test.json | cut -c1-1500
for wl in tiger@3840x2160 reschart@1920x1080 synth_16k; do timeout 120 python tools/lat_frame.py $wl 20 2>&1 | tail -1; done
