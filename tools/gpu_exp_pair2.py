"""Follow-up: MMA-only stream (no loads) with and without the per-k-block mbarrier round trip, single vs CTA pair."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_exp_fill import t  # noqa: E402

for (M, N, K) in ((8192, 8192, 8192),):
    for bn in (256, 192, 128, 64):
        for force, name in ((0x2000, "single"), (0x1000, "pair  ")):
            row = []
            for debug, dn in ((3, "no loads"), (7, "no loads/sync"), (15, "no loads/sync/epi"), (0, "full")):
                try:
                    us, tf = t(M, N, K, bn, debug, force=force)
                    row.append(f"{dn}: {us:7.1f} us {tf:5.0f} TF")
                except Exception as e:  # noqa: BLE001
                    row.append(f"{dn}: ERR {str(e)[:40]}")
            print(f"M{M} N{N} K{K} bn{bn} {name}: " + " | ".join(row), flush=True)
