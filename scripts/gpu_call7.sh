#!/bin/bash
# round-2 GPU call 7: lock-step row compaction (tests + A/B), statistics staged in HBM instead of single-value PCIe writes (A/B)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest (lock-step + tc + sampling seams)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_sampling_seam.py tests/test_step_seam.py -m gpu -q 2>&1 | tail -8
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
echo "=== radon e2e A/B: stats staged (default) | stats direct (round-2 calls 1-6) | everything staged (copy after the kernel)"
timeout 300 $B > gpurun_out/r2g_radon_stats_staged.json 2> gpurun_out/r2g_radon_stats_staged.err
B200_DIRECT_STATS=1 timeout 300 $B > gpurun_out/r2g_radon_stats_direct.json 2> gpurun_out/r2g_radon_stats_direct.err
B200_NO_DIRECT_HOST_WRITES=1 timeout 300 $B > gpurun_out/r2g_radon_all_staged.json 2> gpurun_out/r2g_radon_all_staged.err
timeout 300 $B > gpurun_out/r2g_radon_stats_staged2.json 2> gpurun_out/r2g_radon_stats_staged2.err
python - <<'P'
import json
for t in ["stats_staged", "stats_direct", "all_staged", "stats_staged2"]:
    try:
        d = json.loads(open(f"gpurun_out/r2g_radon_{t}.json").read().strip().splitlines()[-1])
        print(t, "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (d["e2e"]["value"] / 1e6), "e2e ms %.1f" % d["e2e"]["ms_per_step"],
              "kernel ms %.1f" % d["roofline"]["kernel_ms"], d["e2e"].get("host_link"))
    except Exception as e:
        print(t, "FAILED", e)
P
L="python bench.py --workload logistic --precision tc_fp16x2 --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline"
echo "=== logistic tc: compaction on | off"
timeout 400 $L > gpurun_out/r2g_logistic_tc_compact.json 2> gpurun_out/r2g_logistic_tc_compact.err
B200_LS_COMPACT=0 timeout 400 $L > gpurun_out/r2g_logistic_tc_nocompact.json 2> gpurun_out/r2g_logistic_tc_nocompact.err
echo "=== logistic fp64: compaction on"
timeout 600 python bench.py --workload logistic --tune 100 --draws 50 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2g_logistic_fp64_compact.json 2> gpurun_out/r2g_logistic_fp64_compact.err
python - <<'P'
import json
for t in ["logistic_tc_compact", "logistic_tc_nocompact", "logistic_fp64_compact"]:
    try:
        d = json.loads(open(f"gpurun_out/r2g_{t}.json").read().strip().splitlines()[-1])
        print(t, "value %.1fk" % (d["value"] / 1e3), "ms/step %.0f" % d["ms_per_step"], "launches", d["gpu_launches"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(t, "FAILED", e)
P
