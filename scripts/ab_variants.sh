#!/bin/bash
# Build A/B variants of libb200nuts.so locally (nvcc cross-compiles without a GPU) into variants/ (git-ignored, but
# shipped to the GPU box by gpurun).  usage: scripts/ab_variants.sh "tag1:-DFLAG=1 -DOTHER=2" "tag2:..."
# Prints registers/spills of the bench kernel per variant.  Then on the box: scripts/ab_run.sh tag1 tag2 ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for v in "$@"; do
  tag=${v%%:*}; defs=${v#*:}
  ( B200_DEFS="$defs" B200_OUT=variants/lib_$tag.so B200_PTXAS_V=1 ./build.sh 2>&1 \
      | grep -A2 "nuts_warp_kernelINS_10RadonModelELi6ELi1" | grep -E "registers|spill" | sed "s/^/[$tag] /" ) &
done
wait
ls -la variants/
