#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py launches <launches.csv> <out.md>        per-kernel totals and shares of a --metrics gpu__time_duration.sum pass
  python profiles/summarize.py full <prof.ncu-rep> <out.csv> [regex]   key raw metrics of every profiled launch of a --set full capture
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEY = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
       "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
       "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__maximum_warps_per_active_cycle_pct", "launch__shared_mem_per_block_static"]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        agg.setdefault(r[ki], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): gpu__time_duration.sum per launch, --clock-control none\n\n")
        f.write("Per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n\n")
        f.write(f"| kernel | launches | total {data[0][ui]} | avg {data[0][ui]} | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:110]}` | {len(v)} | {sum(v):.0f} | {sum(v) / len(v):.0f} | {sum(v) / tot:.3f} |\n")
        ours = sum(sum(v) for k, v in agg.items() if "onk::" in k or "onika" in k)
        f.write(f"\nShare of this library's kernels (onk::*) in the listed device time: {ours / tot:.3f}\n")


def full(rep, dst, pattern=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(k) for k in KEY if k in hdr]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            if pattern is None or re.search(pattern, r[hdr.index("Kernel Name")]):
                w.writerow([r[i] for i in idx])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
