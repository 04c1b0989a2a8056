#!/bin/bash
# round-2 GPU call 6: the complete GPU suite on the final kernels + final single-GPU bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench radon"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 400 gpurun_out/r2f_bench.json; echo
echo "=== bench logistic tc"; timeout 400 python bench.py --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2f_bench_logistic_tc.json 2> gpurun_out/r2f_bench_logistic_tc.err; head -c 300 gpurun_out/r2f_bench_logistic_tc.json; echo
echo "=== bench mvgauss tc"; timeout 500 python bench.py --workload mvgauss --precision tc_fp16x2 --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2f_bench_mvgauss_tc.json 2> gpurun_out/r2f_bench_mvgauss_tc.err; head -c 300 gpurun_out/r2f_bench_mvgauss_tc.json; echo
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; tail -c 700 gpurun_out/r2f_bench_ref.json; echo
