#!/bin/bash
# same-box A/B of the forward attention kernel options
run() { for shape in "8 10 4096 4096" "8 20 1024 1024" "2 10 4096 4096" "2 20 1024 1024"; do
  timeout 120 python tools/gpu_prof_attn.py $shape 2>&1 | tail -1 | sed "s/^/$1 /" | cut -c1-100; done; }
SB200_ATTN_POLY=4 SB200_ATTN_SPLIT_EXP=0 run "two-cta poly=4 split_exp=0"
SB200_ATTN_POLY=4 SB200_ATTN_SPLIT_EXP=1 run "two-cta poly=4 split_exp=1"
SB200_ATTN_POLY=0 SB200_ATTN_SPLIT_EXP=1 run "two-cta poly=0 split_exp=1"
SB200_ATTN_POLY=8 SB200_ATTN_SPLIT_EXP=1 run "two-cta poly=8 split_exp=1"
