"""Same-box matrix: {single, pair} x {lane0 issue, elect issue} x {full, no TMA loads, no loads + no epilogue} over the
GEMM shapes that dominate the SDXL forward.  Timing only (suppressed variants compute garbage)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gpu_exp_fill import t  # noqa: E402

shapes = ((8192, 8192, 8192, 256), (8192, 1280, 1280, 256), (8192, 10240, 1280, 256), (8192, 1280, 5120, 256),
          (8192, 3840, 1280, 256), (32768, 640, 640, 160), (2048, 1280, 1280, 128))
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for (M, N, K, bn) in shapes:
    for force, kname in ((0x2000, "single"), (0x1000, "pair  ")):
        for el, ename in ((0, "lane0"), (32, "elect")):
            row = []
            for debug, dn in ((0, "full"), (3, "noload"), (11, "noload+noepi")):
                try:
                    us, tf = t(M, N, K, bn, debug | el, force=force)
                    row.append(f"{dn} {us:7.1f}us {tf:5.0f}TF")
                except Exception as e:  # noqa: BLE001
                    row.append(f"{dn} ERR {str(e)[:30]}")
            print(f"M{M} N{N} K{K} bn{bn} {kname} {ename}: " + " | ".join(row), flush=True)
