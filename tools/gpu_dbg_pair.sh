#!/bin/bash
# each case in its own process under a short timeout
run() { timeout 25 python tools/gpu_dbg_pair.py "$@" 2>&1 | tail -2; rc=${PIPESTATUS[0]}; [ "$rc" != "0" ] && echo "  rc=$rc for $*"; }
for dbg in 0 3 11; do
  run 1024 1024 1024 128 0x1000 $dbg 3
  run 8192 1024 1024 128 0x1000 $dbg 3
  run 8192 8192 1024 128 0x1000 $dbg 3
  run 8192 8192 8192 128 0x1000 $dbg 3
  run 8192 8192 8192 64 0x1000 $dbg 3
  run 8192 8192 8192 192 0x1000 $dbg 3
done
