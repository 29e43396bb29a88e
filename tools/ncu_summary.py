#!/usr/bin/env python
"""Summarise ncu outputs into the text files kept under profiles/.
  launch list : python tools/ncu_summary.py launches <csv>            (gpu__time_duration.sum per launch)
  full report : python tools/ncu_summary.py full <file.ncu-rep>       (key counters per captured kernel)
"""
import collections
import csv
import statistics
import subprocess
import sys


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    cols = rows[hdr]
    ki, vi = cols.index("Kernel Name"), cols.index("Metric Value")
    mi = cols.index("Metric Name") if "Metric Name" in cols else None
    agg = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > vi and (mi is None or r[mi] == "gpu__time_duration.sum"):
            agg[r[ki].split("(")[0]].append(float(r[vi].replace(",", "")) / 1e6)
    tot = sum(sum(v) for v in agg.values())
    print("# per-kernel device time, cold-cache and serialised under ncu: compare SHARES, not absolutes")
    print("%-44s %6s %12s %10s %10s %7s" % ("kernel", "n", "total_ms", "avg_ms", "median_ms", "share"))
    for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
        print("%-44s %6d %12.3f %10.4f %10.4f %6.1f%%" % (k, len(v), sum(v), sum(v) / len(v), statistics.median(v),
                                                          100 * sum(v) / tot))
    print("%-44s %6d %12.3f" % ("TOTAL", sum(len(v) for v in agg.values()), tot))


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s" % r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                print("   %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
        st = [(h, float(r[i])) for i, h in enumerate(hdr)
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and r[i]]
        print("   top stalls (warps per issue-active): " +
              ", ".join("%s=%.2f" % (h[34:-23], v) for h, v in sorted(st, key=lambda x: -x[1])[:5]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
