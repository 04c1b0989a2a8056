#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus 8): the BASELINE matrix at its stated GPU counts (VERDICT r1 "next round" item 6):
#   #5 dense Gaussian n=1e4, 256 chains in total, strong scaling over 1/2/4/8 GPUs (fp64 DMMA and the tcgen05 mode)
#   #4 stochastic volatility, 512 chains on 2 GPUs
#   #3 logistic GLM 1e6 x 128, 4096 chains on 8 GPUs (the tcgen05 performance mode and fp64 DMMA parity mode)
#   (#2 Radon at N = 1..8 is the driver's own scaling run and is not repeated here: an 8-GPU box is charged 8x)
# An 8-GPU box is charged 8x, so runs that need fewer GPUs share a phase on DISJOINT GPUs (no data-path collective in
# any of them; every rank has its own host process); every N of one scaling curve is measured on this same box.
# Every line lands in gpurun_out/r2_scale_<tag>_n<N>.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_CACHE_DIR=/dev/shm/b200_cache
run() {  # devs tag args...   (devs = comma list of GPU ids; N = its length)
  local devs=$1 tag=$2; shift 2
  local n=$(echo "$devs" | tr ',' '\n' | wc -l)
  local out=gpurun_out/r2_scale_${tag}_n${n}
  if [ "$n" = 1 ]; then
    CUDA_VISIBLE_DEVICES=$devs timeout 300 python bench.py --gpus 1 "$@" > $out.json 2> $out.err
  else
    CUDA_VISIBLE_DEVICES=$devs timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
      --master-port $((29600 + n + 10 * ${devs%%,*})) bench.py --gpus $n "$@" > $out.json 2> $out.err
  fi
  echo "== $tag n=$n devs=$devs rc=$? $(tail -n 1 $out.json | cut -c1-200)"
}
MV="--workload mvgauss --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
nvidia-smi -L | head -8; nproc
t0=$SECONDS
echo "### phase 0: build the Gaussian's matrices once (cache in /dev/shm)"
python -c "
import os, time; t=time.time()
from pymc_b200 import models; models.mvgauss(cache_dir=os.environ['B200_CACHE_DIR']); print('mvgauss matrices', round(time.time()-t,1), 's')"
echo "### phase A ($((SECONDS-t0)) s): mvgauss fp64 N=4 | N=2 | N=1, mvgauss tc N=1"
run 0,1,2,3 mvgauss $MV &
run 4,5 mvgauss $MV &
run 6 mvgauss $MV &
run 7 mvgauss_tc $MV --precision tc_fp16x2 &
wait
echo "### phase B ($((SECONDS-t0)) s): mvgauss tc N=4, stochvol N=2, logistic fp64 N=1, logistic tc N=1"
run 0,1,2,3 mvgauss_tc $MV --precision tc_fp16x2 &
run 4,5 stochvol --workload stochvol --steps 2 --warmup 1 --no-cpu-baseline &
run 6 logistic_fp64 --workload logistic --tune 30 --draws 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e &
run 7 logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline &
wait
echo "### phase C ($((SECONDS-t0)) s): mvgauss N=8 (fp64, then tc)"
run 0,1,2,3,4,5,6,7 mvgauss $MV
run 0,1,2,3,4,5,6,7 mvgauss_tc $MV --precision tc_fp16x2
echo "### phase D ($((SECONDS-t0)) s): logistic N=8, 4096 chains (tc with e2e, then fp64 short)"
run 0,1,2,3,4,5,6,7 logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline
run 0,1,2,3,4,5,6,7 logistic_fp64 --workload logistic --tune 30 --draws 10 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
echo "### done ($((SECONDS-t0)) s)"
du -sh gpurun_out
