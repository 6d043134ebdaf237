#!/bin/bash
# multi-GPU legs: frame-parallel (bench default) and row bands of the 16K frame. usage: gpu_scale.sh N [bands-only]
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-2}
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/scale_err.log | tail -1; }
if [ -z "$2" ]; then echo "=== frames x$N (synth_1m_4k per rank)"; run --steps 30 --warmup 5 --no-cpu-baseline | tee gpurun_out/scale_frames_$N.json | cut -c1-200; fi
echo "=== bands x$N (synth_16k)"; run --steps 10 --warmup 3 --workload synth_16k --mode bands --no-cpu-baseline | tee gpurun_out/scale_bands_$N.json | cut -c1-200
echo "=== bands x$N, no gather"; run --steps 10 --warmup 3 --workload synth_16k --mode bands --no-gather --no-cpu-baseline | tee gpurun_out/scale_bands_nogather_$N.json | cut -c1-200
grep -v "Warning\|^\*\*\*\|^$" gpurun_out/scale_err.log | tail -3
