#!/bin/bash
# round-2 GPU call 9: staging arena for every host array (no allocator calls in a steady-state call), pinned buffers near the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
for rep in 1 2 3; do
  B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_radon_$rep.json 2> gpurun_out/r2i_radon_$rep.err
  python - "$rep" <<'P'
import json, sys
r = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2i_radon_{r}.json").read().strip().splitlines()[-1])
    e = d["e2e"]
    print(r, "value %.1fM" % (d["value"] / 1e6), "e2e %.1fM" % (e["value"] / 1e6), "calls", [round(x) for x in e["call_ms"]], "kernels", [round(x) for x in e["call_kernel_ms"]], e.get("host_numa"), e.get("host_link"))
except Exception as ex:
    print(r, "FAILED", ex)
P
  grep "b200_nuts_run" gpurun_out/r2i_radon_$rep.err | tail -3
done
echo "=== stochvol"; timeout 300 python bench.py --workload stochvol --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2i_stochvol.json 2> gpurun_out/r2i_stochvol.err; head -c 200 gpurun_out/r2i_stochvol.json; echo
