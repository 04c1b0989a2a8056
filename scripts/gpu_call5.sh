#!/bin/bash
# round-2 GPU call 5: tensor-core logistic pass with the coalesced fp64 drain; lock-step arena; bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tc"; timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -12
echo "=== lockstep parity"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "logistic or mvgauss or dense" 2>&1 | tail -4
echo "=== bench logistic tc"; timeout 400 python bench.py --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_bench_logistic_tc.json 2> gpurun_out/r2e_bench_logistic_tc.err; head -c 300 gpurun_out/r2e_bench_logistic_tc.json; echo
echo "=== bench mvgauss tc"; timeout 500 python bench.py --workload mvgauss --precision tc_fp16x2 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_mvgauss_tc.json 2> gpurun_out/r2e_bench_mvgauss_tc.err; head -c 300 gpurun_out/r2e_bench_mvgauss_tc.json; echo
echo "=== bench mvgauss fp64"; timeout 500 python bench.py --workload mvgauss --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2e_bench_mvgauss.json 2> gpurun_out/r2e_bench_mvgauss.err; head -c 300 gpurun_out/r2e_bench_mvgauss.json; echo
echo "=== ncu tc"; timeout 300 ./scripts/profile_round.sh r2e tc 2>&1 | tail -3
