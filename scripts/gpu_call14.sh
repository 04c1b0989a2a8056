#!/bin/bash
# round-2 GPU call 14: stochastic volatility with a chain spread over 16 warps (A/B against 8), launch lists of the dense
# Gaussian loop with deferred momentum service, the default bench command at this commit
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_CACHE_DIR=/dev/shm/b200_cache
t0=$SECONDS
b() { local tag=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r2n_bench_$tag.json 2> gpurun_out/r2n_bench_$tag.err; echo "--- $tag rc=$? $(tail -n 1 gpurun_out/r2n_bench_$tag.json | cut -c1-230)"; }
echo "=== [$((SECONDS-t0)) s] stochvol, 8 warps per chain"; b stochvol_w8 --workload stochvol --steps 2 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] stochvol, 16 warps per chain"; B200_TEAM_W16=1 b stochvol_w16 --workload stochvol --steps 2 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] stochvol parity tests with 16 warps per chain"; B200_TEAM_W16=1 timeout 300 python -m pytest tests -m gpu -q -rf -k "stochvol" 2>&1 | tail -6
echo "=== [$((SECONDS-t0)) s] launch lists: dense Gaussian"
NCU="ncu --target-processes application-only --clock-control none --metrics gpu__time_duration.sum -c 600 --csv"
timeout 300 $NCU --log-file gpurun_out/r2n_launches_mvgauss.csv python scripts/ncu_target3.py mvgauss 6 3 > gpurun_out/r2n_launches_mvgauss.log 2>&1; tail -n 1 gpurun_out/r2n_launches_mvgauss.log
timeout 300 $NCU --log-file gpurun_out/r2n_launches_mvgauss_tc.csv python scripts/ncu_target3.py mvgauss_tc 6 3 > gpurun_out/r2n_launches_mvgauss_tc.log 2>&1; tail -n 1 gpurun_out/r2n_launches_mvgauss_tc.log
echo "=== [$((SECONDS-t0)) s] default bench command"; b radon_default
echo "=== [$((SECONDS-t0)) s] done"; du -sh gpurun_out
