#!/bin/bash
# round-2 GPU call 1: full GPU test suite, smoke, prepared-variant A/B, bench line, StochVol ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== variants"; timeout 900 ./scripts/ab_run.sh default sched ms u2 pk@4,0 pk144@14,1 2>&1
echo "=== bench"; timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
echo "=== stochvol ncu"; timeout 400 ./scripts/profile_round.sh r2 stochvol 2>&1 | tail -5
