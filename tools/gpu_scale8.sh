#!/bin/bash
# 8-GPU legs of BASELINE cfg4: row bands of the 16K frame, band-resident and gathered to rank 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
N=${1:-8}
mkdir -p gpurun_out
export TORCH_NCCL_SHOW_EAGER_INIT_P2P_SERIALIZATION_WARNING=false
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" 2>gpurun_out/scale_err.log | tail -1; }
echo "=== bands x$N, no gather"; run --steps 20 --warmup 4 --workload synth_16k --mode bands --no-gather --no-cpu-baseline | tee gpurun_out/scale_bands_nogather_$N.json | cut -c1-260
echo "=== bands x$N gathered"; run --steps 20 --warmup 4 --workload synth_16k --mode bands --no-cpu-baseline | tee gpurun_out/scale_bands_$N.json | cut -c1-260
grep -o '"bands_vs_full_frame": {[^}]*}' gpurun_out/scale_bands_$N.json
grep -v "Warning\|^\*\*\*\|^$" gpurun_out/scale_err.log | tail -3
