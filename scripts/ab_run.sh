#!/bin/bash
# Run ON THE GPU BOX (inside ONE gpurun call: every call costs ~90 s of budget before the command starts):
#   gpurun --timeout 900 -- './scripts/ab_run.sh default ms u2 pk@4,0 pk144@14,1'
# Each argument is  tag[@wpb,hot]  (tag = variants/lib_<tag>.so built by scripts/ab_variants.sh; `default` = the in-tree
# library; wpb,hot = chains per CTA and hot stack levels of the persistent kernel, default 4,2).
# For each variant: GPU parity tests (fail fast, 120 s cap), then the bench-shaped probe (gpu_probe9: per-chain work
# distribution + kernel ms) and the micro probe (gpu_probe10: leapfrog throughput, solo-warp latency).  One block of lines
# per variant; nothing is written to gpurun_out/ (keep it under the 64 MiB merge limit).
cd "$(dirname "$0")/.."
for arg in "$@"; do
  tag=${arg%%@*}; cfg=4,2; [[ "$arg" == *@* ]] && cfg=${arg#*@}
  if [ "$tag" = default ]; then unset B200_LIB; else export B200_LIB=$PWD/variants/lib_$tag.so; fi
  export B200_NUTS_WPB=${cfg%%,*} B200_NUTS_HOT=${cfg#*,}
  echo "=== $tag (wpb,hot = $cfg)"
  timeout 150 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_fullsize.py 2>&1 | tail -1
  timeout 90 python scripts/gpu_probe9.py $cfg 2>&1 | tail -1 | cut -c1-260
  timeout 60 python scripts/gpu_probe10.py 2>&1 | tail -4
done
