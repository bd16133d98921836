"""Experiment: UMMA rate of the single-CTA kernel vs the cta_group::2 CTA-pair kernel with the TMA loads (and the
epilogue) suppressed through the debug bits of `bn` — is the pair's MMA stream itself faster?  Timing only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_exp_fill import t  # noqa: E402

for (M, N, K) in ((8192, 8192, 8192), (8192, 10240, 1280)):
    for bn in (256, 128):
        for force, name in ((0x2000, "single"), (0x1000, "pair  ")):
            row = []
            for debug, dn in ((0, "full"), (3, "no loads"), (11, "no loads, no epilogue"), (0, "full again")):
                try:
                    us, tf = t(M, N, K, bn, debug, force=force)
                    row.append(f"{dn}: {us:7.1f} us {tf:5.0f} TF")
                except Exception as e:  # noqa: BLE001
                    row.append(f"{dn}: ERR {str(e)[:40]}")
            print(f"M{M} N{N} K{K} bn{bn} {name}: " + " | ".join(row), flush=True)
