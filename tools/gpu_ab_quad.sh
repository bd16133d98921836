#!/bin/bash
# same-box A/B of 2x2 clusters with A multicast (SB200_QUAD) on the L2-bound GEMM shapes; each case in its own process
for shape in "8192 1280 1280 0" "8192 1280 1280 5" "8192 1280 1280 21" "8192 1280 5120 5" "8192 3840 1280 16" "32768 640 640 5" "32768 1920 640 16" "2048 1280 1280 21" "8192 10240 1280 9"; do
  for q in 0 2; do
    SB200_QUAD=$q timeout 60 python tools/gpu_prof_gemm.py $shape 2>&1 | tail -1 | sed "s/^/quad=$q /" | cut -c1-120
  done
done
