#!/bin/bash
# Run ON THE GPU BOX (inside ONE gpurun call: every call costs ~90 s of budget before the command starts):
#   gpurun --timeout 600 -- './scripts/ab_run.sh default tag1 tag2'
# For each variant: GPU parity tests (fail fast, 120 s cap), then the bench-shaped probe (gpu_probe9: per-chain work
# distribution + kernel ms) and the micro probe (gpu_probe10: leapfrog throughput, solo-warp latency).  One block of lines
# per variant; nothing is written to gpurun_out/ (keep it under the 64 MiB merge limit).
cd "$(dirname "$0")/.."
for tag in "$@"; do
  if [ "$tag" = default ]; then unset B200_LIB; else export B200_LIB=$PWD/variants/lib_$tag.so; fi
  echo "=== $tag"
  timeout 120 python -m pytest tests -q -m gpu -x 2>&1 | tail -1
  timeout 90 python scripts/gpu_probe9.py 4,2 2>&1 | tail -2 | cut -c1-260
  timeout 60 python scripts/gpu_probe10.py 2>&1 | tail -4
done
