#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus 8): the BASELINE matrix at its stated GPU counts:
#   #5 dense Gaussian n=1e4, 256 chains, strong scaling over 1/2/4/8 GPUs
#   #4 stochastic volatility, 512 chains on 2 GPUs
#   #3 logistic GLM 1e6 x 128, 4096 chains on 8 GPUs (fp64 DMMA parity mode and the tcgen05 performance mode)
# every line lands in gpurun_out/r2_scale_*.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_CACHE_DIR=/dev/shm/b200_cache
run() {  # n tag args...
  local n=$1 tag=$2; shift 2
  local devs=$(seq -s, 0 $((n-1)))
  if [ "$n" = 1 ]; then
    CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/r2_scale_${tag}_n${n}.json 2> gpurun_out/r2_scale_${tag}_n${n}.err
  else
    CUDA_VISIBLE_DEVICES=$devs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n "$@" > gpurun_out/r2_scale_${tag}_n${n}.json 2> gpurun_out/r2_scale_${tag}_n${n}.err
  fi
  echo "== $tag n=$n rc=$?"; tail -n 1 gpurun_out/r2_scale_${tag}_n${n}.json | cut -c1-330
}
nvidia-smi -L | head -8
for n in 1 2 4 8; do run $n mvgauss --workload mvgauss --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e; done
for n in 1 8; do run $n mvgauss_tc --workload mvgauss --precision tc_fp16x2 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e; done
run 2 stochvol --workload stochvol --steps 1 --warmup 1 --no-cpu-baseline
run 8 logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline
run 8 logistic_fp64 --workload logistic --tune 30 --draws 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
du -sh gpurun_out
