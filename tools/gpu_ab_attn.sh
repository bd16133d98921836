#!/bin/bash
# same-box A/B of the forward attention kernels (two-CTA vs ping-pong), of the polynomial-exp2 share and of the
# explicit MUFU hand-over in the ping-pong kernel
run() { for shape in "8 10 4096 4096" "8 20 1024 1024" "2 10 4096 4096"; do
  timeout 120 python tools/gpu_prof_attn.py $shape 2>&1 | tail -1 | sed "s/^/$1 /" | cut -c1-100; done; }
SB200_ATTN_PP=0 SB200_ATTN_POLY=0 run "two-cta poly=0        "
SB200_ATTN_PP=0 SB200_ATTN_POLY=4 run "two-cta poly=4        "
for tok in 0 1; do for poly in 0 8 4; do
  SB200_ATTN_PP=1 SB200_ATTN_POLY=$poly SB200_ATTN_TOKEN=$tok run "pingpong poly=$poly token=$tok"
done; done
