#!/bin/bash
# round-2 GPU call 16 (2 GPUs, the last of the budget): BASELINE config 4 at its stated GPU count (512 chains over 2 GPUs, NCCL),
# then the Radon bench at N = 2 with the end-to-end leg (the path that failed in the 8-GPU call before the fix)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run2() { local tag=$1; shift; timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus 2 "$@" > gpurun_out/r2_scale_${tag}_n2.json 2> gpurun_out/r2_scale_${tag}_n2.err; echo "--- $tag rc=$? $(grep '^{' gpurun_out/r2_scale_${tag}_n2.json | tail -n 1 | cut -c1-220)"; }
run2 stochvol --workload stochvol --steps 1 --warmup 1 --no-cpu-baseline
run2 radon --steps 1 --warmup 1 --no-cpu-baseline
