#!/bin/bash
# 2-GPU legs: frame-parallel default (what the driver's scaling run launches) or, with "bands", the 16K frame in two exact row bands
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 "$@" 2>gpurun_out/scale_err.log | tail -1; }
if [ "$1" = bands ]; then run --steps 15 --warmup 3 --workload synth_16k --mode bands --no-gather --no-cpu-baseline | tee gpurun_out/scale_bands_nogather_2.json | cut -c1-250
else run --steps 30 --warmup 5 | tee gpurun_out/scale_frames_2.json | cut -c1-400; fi
grep -v "Warning\|^\*\*\*\|^$\|OMP_NUM_THREADS" gpurun_out/scale_err.log | tail -3
