"""Stall samples of one kernel grouped by code region (split at barrier / TMEM / MMA instructions):
    python tools/ncu_phases.py report.ncu-rep [min_samples]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
isrc, iss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(int(x[iss] or 0), x[isrc].strip(), int(x[ie] or 0)) for x in rows[2:] if len(x) > iss]
tot = sum(d[0] for d in data)
print(rows[0][1], "total samples", tot)
acc, start = 0, 0
for i, d in enumerate(data):
    acc += d[0]
    if re.search(r"SYNCS|BAR\.|UTCHMMA|UTCBAR|LDTM|STTM|UTMALDG|EXIT|ELECT", d[1]):
        if acc >= thr:
            print(f"#{start:5d}-{i:5d} {acc:6d} {100 * acc / tot:5.1f}%  exec {d[2]:8d} | {d[1][:72]}")
        acc, start = 0, i + 1
