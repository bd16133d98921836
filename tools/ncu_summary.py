"""Key metrics of every kernel in .ncu-rep files -> text (the committed summaries under profiles/):
    python tools/ncu_summary.py gpurun_out/a.ncu-rep [b.ncu-rep ...] > profiles/xyz.txt"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "gpc__cycles_elapsed.max.per_second",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor"]
for path in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print(f"== {path.split('/')[-1]} :: {r[idx['Kernel Name']][:90]}")
        for w in WANT:
            if w in idx:
                print(f"{w:78s} {r[idx[w]]:>14s} {units[idx[w]]}")
        st = [(h, float(r[idx[h]].replace(',', ''))) for h in hdr
              if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')
              and r[idx[h]] not in ('', 'n/a')]
        top = sorted(st, key=lambda x: -x[1])[:5]
        print("top stall reasons (warps per issue): " + ", ".join(
            f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}"
            for h, v in top))
