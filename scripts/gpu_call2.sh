#!/bin/bash
# round-2 GPU call 2: GPU tests (new default: scheduled kernel + (m,s) weights + analytic padding; IR), scheduling/occupancy A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q --ignore=tests/test_gpu_tc.py 2>&1 | tail -15
echo "=== tc"; timeout 240 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -15
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== variants"; AB_FAST=1 timeout 900 ./scripts/ab_run.sh default@4,2 default@4,2,10 default@4,2,50 default@4,2,100 default@4,1 default@4,3 default@8,2 m224@3,2 m224@3,1 m200@5,2 2>&1
echo "=== m224 tests"; B200_LIB=$PWD/variants/lib_m224.so B200_NUTS_WPB=3 timeout 200 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_fullsize.py --ignore=tests/test_ir.py 2>&1 | tail -2
echo "=== bench"; timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 600 gpurun_out/r2b_bench.json
echo "=== logistic default"; timeout 300 python bench.py --workload logistic --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-400
echo "=== logistic fast-log"; B200_LIB=$PWD/variants/lib_fl.so timeout 300 python bench.py --workload logistic --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-400
