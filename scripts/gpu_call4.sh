#!/bin/bash
# round-2 GPU call 4: tensor-core kernels after the accumulate-rounding fixes (logistic: 8 epilogue warps, per-slab drains,
# small products first; new tcgen05 GEMM of the dense Gaussian), remaining test fixes, their bench lines and ncu captures
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tc"; timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -30
echo "=== rest"; timeout 600 python -m pytest tests/test_f3_variants.py tests/test_ir.py tests/test_gpu_state.py -m gpu -q 2>&1 | tail -8
echo "=== bench logistic tc"; timeout 400 python bench.py --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2d_bench_logistic_tc.json 2> gpurun_out/r2d_bench_logistic_tc.err; head -c 300 gpurun_out/r2d_bench_logistic_tc.json; echo
echo "=== bench mvgauss tc"; timeout 500 python bench.py --workload mvgauss --precision tc_fp16x2 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2d_bench_mvgauss_tc.json 2> gpurun_out/r2d_bench_mvgauss_tc.err; head -c 300 gpurun_out/r2d_bench_mvgauss_tc.json; echo
echo "=== profiles"; timeout 600 ./scripts/profile_round.sh r2d tc 2>&1 | tail -6
