#!/bin/bash
# round-2 GPU call 12: version 2 of the tcgen05 logistic pass at 96 registers (18 warps: the 112-register build of call 11
# could not launch), its tests, bench lines of both versions on the same box, launch list and one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$SECONDS
echo "=== [$((SECONDS-t0)) s] tc tests (version 2) + the Radon full_adapt case"; timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_dense_adapt.py -m gpu -q -rf 2>&1 | tail -15
cp gpurun_out/parity_report.json gpurun_out/r2l_parity_report.json 2>/dev/null
b() { local tag=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r2l_bench_$tag.json 2> gpurun_out/r2l_bench_$tag.err; echo "--- $tag rc=$? $(tail -n 1 gpurun_out/r2l_bench_$tag.json | cut -c1-260)"; }
echo "=== [$((SECONDS-t0)) s] bench logistic tc v2"; b logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] bench logistic tc v1 (same box)"; B200_LOGI_TC_V=1 b logistic_tc_v1 --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
echo "=== [$((SECONDS-t0)) s] ncu tc2"; timeout 400 ./scripts/profile_round.sh r2l tc2 2>&1 | tail -4
echo "=== [$((SECONDS-t0)) s] done"; du -sh gpurun_out
