#!/bin/bash
# round-2 GPU call 15 (last): the whole GPU suite and smoke() on the final code, the tensor-core logistic pass with the fp64
# sums moved off the drain's critical path (bench line + ncu capture of the final kernel)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$SECONDS
echo "=== [$((SECONDS-t0)) s] whole GPU suite"; timeout 600 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8
cp gpurun_out/parity_report.json gpurun_out/r2o_parity_report.json 2>/dev/null
echo "=== [$((SECONDS-t0)) s] smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
b() { local tag=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r2o_bench_$tag.json 2> gpurun_out/r2o_bench_$tag.err; echo "--- $tag rc=$? $(tail -n 1 gpurun_out/r2o_bench_$tag.json | cut -c1-260)"; }
echo "=== [$((SECONDS-t0)) s] bench logistic tc"; b logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] ncu tc2"; timeout 400 ./scripts/profile_round.sh r2o tc2 2>&1 | tail -4
echo "=== [$((SECONDS-t0)) s] done"; du -sh gpurun_out
