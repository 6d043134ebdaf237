#!/bin/bash
# first GPU bring-up: smoke, parity tests, short bench; everything under its own timeout
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20
echo "=== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_first.json
