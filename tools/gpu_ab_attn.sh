#!/bin/bash
# same-box A/B of the forward attention kernel options (cross-attention shapes: query-tile loop on / off)
run() { for shape in "8 20 1024 77" "8 10 4096 77" "2 20 1024 77" "2 10 4096 77" "8 8 4096 77 40" "8 20 1024 1024"; do
  timeout 120 python tools/gpu_prof_attn.py $shape 2>&1 | tail -1 | sed "s/^/$1 /" | cut -c1-100; done; }
SB200_ATTN_QLOOP=0 run "qloop=0"
SB200_ATTN_QLOOP=1 run "qloop=1"
