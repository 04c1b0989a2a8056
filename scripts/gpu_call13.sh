#!/bin/bash
# round-2 GPU call 13: whole GPU suite on the final code (coalesced feature-major finish kernel of the tensor-core logistic
# pass), bench line of config 3 in performance mode, and the N > 1 host logic of bench.py with two ranks sharing the GPU (gloo)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$SECONDS
echo "=== [$((SECONDS-t0)) s] whole GPU suite"; timeout 600 python -m pytest tests -m gpu -q -rf 2>&1 | tail -8
cp gpurun_out/parity_report.json gpurun_out/r2m_parity_report.json 2>/dev/null
b() { local tag=$1; shift; timeout 400 python bench.py "$@" > gpurun_out/r2m_bench_$tag.json 2> gpurun_out/r2m_bench_$tag.err; echo "--- $tag rc=$? $(tail -n 1 gpurun_out/r2m_bench_$tag.json | cut -c1-260)"; }
echo "=== [$((SECONDS-t0)) s] bench logistic tc"; b logistic_tc --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] launch list logistic tc"; timeout 200 ncu --target-processes application-only --clock-control none --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r2m_launches_logistic_tc.csv python scripts/ncu_target3.py logistic_tc 30 10 > gpurun_out/r2m_launches_logistic_tc.log 2>&1; tail -n 1 gpurun_out/r2m_launches_logistic_tc.log
two() { local tag=$1; shift; B200_BENCH_ONE_GPU=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 "$@" > gpurun_out/r2m_tworank_$tag.json 2> gpurun_out/r2m_tworank_$tag.err; echo "--- two ranks on one GPU: $tag rc=$? $(grep '^{' gpurun_out/r2m_tworank_$tag.json | tail -n 1 | cut -c1-200)"; grep -n "Error" gpurun_out/r2m_tworank_$tag.err | head -3; }
echo "=== [$((SECONDS-t0)) s] N > 1 host logic: radon"; two radon --chains-per-gpu 256 --tune 150 --draws 100 --steps 1 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] N > 1 host logic: stochvol"; two stochvol --workload stochvol --chains-per-gpu 32 --tune 50 --draws 20 --steps 1 --warmup 1 --no-cpu-baseline
echo "=== [$((SECONDS-t0)) s] done"; du -sh gpurun_out
