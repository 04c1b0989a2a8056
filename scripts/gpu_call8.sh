#!/bin/bash
# round-2 GPU call 8: lock-step advance kernel with tiled loads (all loads of a tile before its stores): parity + A/B of team width / tile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200_CACHE_DIR=/dev/shm/b200_cache
echo "=== pytest (lock-step users)"
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_sampling_seam.py -m gpu -q 2>&1 | tail -6
M="python bench.py --workload mvgauss --precision tc_fp16x2 --tune 30 --draws 15 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
for tag in default w8u4 w16u2 w8u8 w4u4; do
  lib=pymc_b200/libb200nuts.so; [ "$tag" != default ] && lib=variants/lib_$tag.so
  B200_LIB=$PWD/$lib timeout 400 $M > gpurun_out/r2h_mvgauss_tc_$tag.json 2> gpurun_out/r2h_mvgauss_tc_$tag.err
  python - "$tag" <<'P'
import json, sys
t = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2h_mvgauss_tc_{t}.json").read().strip().splitlines()[-1])
    print(t, "value %.1fk" % (d["value"] / 1e3), "ms/step %.0f" % d["ms_per_step"], "launches", d["gpu_launches"])
except Exception as e:
    print(t, "FAILED", e)
P
done
echo "=== launch list of the default build (mvgauss tc, short)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches_mvgauss_tc.csv \
  python bench.py --workload mvgauss --precision tc_fp16x2 --tune 6 --draws 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2h_ncu_bench.log 2>&1
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2h_launches_mvgauss_tc.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1e3 if u in ("ns", "nsecond") else v
    agg[r[ki][:50]].append(v)
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("%-52s n=%4d avg %.1f us total %.1f ms" % (k, len(v), sum(v) / len(v), sum(v) / 1e3))
P
echo "=== mvgauss fp64 (default build)"
timeout 600 python bench.py --workload mvgauss --tune 60 --draws 30 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2h_mvgauss_fp64.json 2> gpurun_out/r2h_mvgauss_fp64.err; head -c 250 gpurun_out/r2h_mvgauss_fp64.json; echo
echo "=== radon e2e: where the time of a public-API call goes (B200_TRACE) -- stats staged (default) vs direct, twice each"
for rep in 1 2; do
  for mode in staged direct; do
    env=""; [ $mode = direct ] && env="B200_DIRECT_STATS=1"
    env $env B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_radon_${mode}_$rep.json 2> gpurun_out/r2h_radon_${mode}_$rep.err
    python - "$mode" "$rep" <<'P'
import json, sys
m, r = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/r2h_radon_{m}_{r}.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print(m, r, "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (e["value"] / 1e6), "calls", [round(x) for x in e["call_ms"]], "kernels", [round(x) for x in e["call_kernel_ms"]])
except Exception as ex:
    print(m, r, "FAILED", ex)
P
    grep "b200_nuts_run" gpurun_out/r2h_radon_${mode}_$rep.err | tail -4
  done
done
