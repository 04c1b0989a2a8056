#!/bin/bash
# round-2 GPU call 11 (second session): QuadPotentialFullAdapt on the GPU, version 2 of the tcgen05 logistic pass, deferred
# momentum service in the lock-step engine; then the whole GPU suite, smoke, bench lines and one ncu capture.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_CACHE_DIR=/dev/shm/b200_cache
t0=$SECONDS
echo "=== [$((SECONDS-t0)) s] GPU suite without the tensor-core tests"; timeout 700 python -m pytest tests -m gpu -q -rf --ignore=tests/test_gpu_tc.py 2>&1 | tail -25
echo "=== [$((SECONDS-t0)) s] tc (version 2 of the logistic kernel)"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -rf 2>&1 | tail -15
cp gpurun_out/parity_report.json gpurun_out/r2k_parity_report.json 2>/dev/null
echo "=== [$((SECONDS-t0)) s] smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
b() { local tag=$1; shift; timeout 500 python bench.py "$@" > gpurun_out/r2k_bench_$tag.json 2> gpurun_out/r2k_bench_$tag.err; echo "--- $tag rc=$? $(tail -n 1 gpurun_out/r2k_bench_$tag.json | cut -c1-260)"; }
echo "=== [$((SECONDS-t0)) s] bench logistic tc v2"; b logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] bench mvgauss tc"; b mvgauss_tc --workload mvgauss --precision tc_fp16x2 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== [$((SECONDS-t0)) s] bench mvgauss tc, momentum served every round"; B200_LS_MOM_DEFER=0 b mvgauss_tc_nodefer --workload mvgauss --precision tc_fp16x2 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== [$((SECONDS-t0)) s] bench mvgauss fp64"; b mvgauss --workload mvgauss --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== [$((SECONDS-t0)) s] bench mvgauss fp64 32 chains (the per-GPU share of the 8-GPU strong-scaling point)"; b mvgauss_c32 --workload mvgauss --chains-per-gpu 32 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== [$((SECONDS-t0)) s] bench radon (default command)"; b radon --steps 3 --warmup 3
echo "=== [$((SECONDS-t0)) s] ncu tc2"; timeout 400 ./scripts/profile_round.sh r2k tc2 2>&1 | tail -4
echo "=== [$((SECONDS-t0)) s] done"; du -sh gpurun_out
