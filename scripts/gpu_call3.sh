#!/bin/bash
# round-2 GPU call 3: full GPU suite, tensor-core mode tests, bench lines of all configs at real sizes, profiles
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_tc.py 2>&1 | tail -25
echo "=== tc"; timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q 2>&1 | tail -25
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench radon"; timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -c 300 gpurun_out/r2c_bench.json
echo "=== probe"; AB_FAST=1 timeout 200 ./scripts/ab_run.sh default@8,2,50 default@4,2,50 default@8,2,75
echo "=== bench logistic tc"; timeout 400 python bench.py --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench_logistic_tc.json 2> gpurun_out/r2c_bench_logistic_tc.err; tail -c 1200 gpurun_out/r2c_bench_logistic_tc.json | cut -c1-1200
echo "=== bench logistic fp64"; timeout 500 python bench.py --workload logistic --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_logistic.json 2> gpurun_out/r2c_bench_logistic.err; head -c 400 gpurun_out/r2c_bench_logistic.json
echo "=== bench stochvol"; timeout 300 python bench.py --workload stochvol --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c_bench_stochvol.json 2> gpurun_out/r2c_bench_stochvol.err; head -c 400 gpurun_out/r2c_bench_stochvol.json
echo "=== bench mvgauss"; timeout 500 python bench.py --workload mvgauss --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2c_bench_mvgauss.json 2> gpurun_out/r2c_bench_mvgauss.err; head -c 400 gpurun_out/r2c_bench_mvgauss.json
echo "=== profiles"; timeout 600 ./scripts/profile_round.sh r2 bench 2>&1 | tail -4; timeout 400 ./scripts/profile_round.sh r2 tc 2>&1 | tail -4
