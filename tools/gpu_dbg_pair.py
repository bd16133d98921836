"""Diagnostic: one (shape, bn, kernel, debug) GEMM call per process, so a hang only costs its own timeout."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_exp_fill import t  # noqa: E402

M, N, K, bn, force, debug, iters = (int(a, 0) for a in sys.argv[1:8])
us, tf = t(M, N, K, bn, debug, iters=iters, force=force)
print(f"OK M{M} N{N} K{K} bn{bn} force{force:#x} debug{debug}: {us:.1f}us {tf:.0f}TF", flush=True)
