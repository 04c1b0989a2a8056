#!/bin/bash
# Run ON THE GPU BOX (inside ONE gpurun call: every call costs ~90 s of budget before the command starts):
#   gpurun --timeout 900 -- './scripts/ab_run.sh default t96@3,2 default@4,2,50'
# Each argument is  tag[@wpb,hot[,seg]]  (tag = variants/lib_<tag>.so built by scripts/ab_variants.sh; `default` = the
# in-tree library; wpb,hot = chains per CTA and hot stack levels of the persistent kernel, default 4,2; seg = iterations
# per scheduling unit, default 25).  Per variant: GPU parity tests (fail fast), the bench-shaped probe (gpu_probe9:
# per-chain work distribution + kernel ms) and the micro probe (gpu_probe10: leapfrog throughput, solo-warp latency).
cd "$(dirname "$0")/.."
FAST=${AB_FAST:-0}
for arg in "$@"; do
  tag=${arg%%@*}; cfg=4,2; [[ "$arg" == *@* ]] && cfg=${arg#*@}
  IFS=, read wpb hot seg <<< "$cfg"
  if [ "$tag" = default ]; then unset B200_LIB; else export B200_LIB=$PWD/variants/lib_$tag.so; fi
  export B200_NUTS_WPB=$wpb B200_NUTS_HOT=${hot:-2}
  if [ -n "$seg" ]; then export B200_NUTS_SEG=$seg; else unset B200_NUTS_SEG; fi
  echo "=== $tag (wpb,hot,seg = $cfg)"
  if [ "$FAST" = 0 ]; then timeout 150 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_fullsize.py --ignore=tests/test_ir.py 2>&1 | tail -1; fi
  timeout 90 python scripts/gpu_probe9.py $wpb,${hot:-2} 2>&1 | tail -1 | cut -c1-260
  if [ "$FAST" = 0 ]; then timeout 60 python scripts/gpu_probe10.py 2>&1 | tail -4; fi
done
