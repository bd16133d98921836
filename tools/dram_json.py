"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch) -> per-kernel means as JSON
(bench.py reads `per_kernel.gemm_kernel` for `roofline.traffic`):   python tools/dram_json.py in.csv "source text" > out.json"""
import collections
import csv
import json
import sys

rows = collections.defaultdict(lambda: collections.defaultdict(list))
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0].replace("sb200::", "").replace("void ", "").split("<")[0]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    else:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rows[name][r["Metric Name"]].append(v)
out = {"source": sys.argv[2] if len(sys.argv) > 2 else "", "per_kernel": {}}
for k, m in rows.items():
    n = len(m["gpu__time_duration.sum"])
    out["per_kernel"][k] = {"launches": n,
                            "dram_read_bytes_per_launch": sum(m["dram__bytes_read.sum"]) / n,
                            "dram_write_bytes_per_launch": sum(m["dram__bytes_write.sum"]) / n,
                            "time_us_per_launch": sum(m["gpu__time_duration.sum"]) / n}
json.dump(out, sys.stdout, indent=1)
