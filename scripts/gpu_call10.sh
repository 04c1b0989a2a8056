#!/bin/bash
# round-2 GPU call 10: gemm_tc epilogue A/B -- compensated fp32 sums (default) vs plain fp32 sums; error (test_gpu_tc) + time
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_CACHE_DIR=/dev/shm/b200_cache
for tag in default gtplain; do
  lib=pymc_b200/libb200nuts.so; [ "$tag" != default ] && lib=variants/lib_$tag.so
  export B200_LIB=$PWD/$lib
  echo "=== [$tag] tc tests (dense Gaussian)"
  rm -f gpurun_out/parity_report.json
  timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -q -k "gemm" 2>&1 | tail -4
  python - "$tag" <<'P'
import json, sys
try:
    r = json.load(open("gpurun_out/parity_report.json"))
    for k, v in r.items():
        if "mvgauss" in k:
            print(" ", sys.argv[1], k, {a: (round(b, 10) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("no report", e)
P
  cp gpurun_out/parity_report.json gpurun_out/r2j_parity_$tag.json 2>/dev/null
  timeout 400 python bench.py --workload mvgauss --precision tc_fp16x2 --tune 30 --draws 15 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2j_mvgauss_tc_$tag.json 2> gpurun_out/r2j_mvgauss_tc_$tag.err
  python - "$tag" <<'P'
import json, sys
t = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2j_mvgauss_tc_{t}.json").read().strip().splitlines()[-1])
    print(" ", t, "value %.1fk" % (d["value"] / 1e3), "ms/step %.0f" % d["ms_per_step"])
except Exception as e:
    print(" ", t, "FAILED", e)
P
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2j_launches_$tag.csv \
    python bench.py --workload mvgauss --precision tc_fp16x2 --tune 4 --draws 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2j_ncu_$tag.log 2>&1
  python - "$tag" <<'P'
import csv, collections, sys
rows = [r for r in csv.reader(open(f"gpurun_out/r2j_launches_{sys.argv[1]}.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); v = v / 1e3 if r[ui] in ("ns", "nsecond") else v
    agg[r[ki][:40]].append(v)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))[:4]:
    print("   %-42s n=%4d avg %.1f us" % (k, len(v), sum(v) / len(v)))
P
done
