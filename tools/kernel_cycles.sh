#!/bin/bash
# Per-kernel SM cycles (clock-independent A/B metric; boxes differ in power-capped clocks) for one chain batch of the C3 bench.
#   tools/kernel_cycles.sh <tag> [extra bench args]   -> gpurun_out/cycles_<tag>.csv and a per-kernel summary on stdout
tag=$1; shift
ncu --metrics sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"tc_g|tc_fused|k_layer2<" -s 9 -c 12 --csv --log-file gpurun_out/cycles_$tag.csv \
    python bench.py --chains 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "$@" > gpurun_out/cycles_$tag.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/cycles_$tag.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
c = rows[h]; ki, mi, vi = c.index("Kernel Name"), c.index("Metric Name"), c.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[h + 1:]:
    agg[r[ki][:40]][r[mi]].append(float(r[vi].replace(",", "")))
for k, m in agg.items():
    print("%-42s n=%d cycles=%.3fM time=%.3fms tensor=%.1f%%" % (k, len(m["sm__cycles_elapsed.avg"]),
          sum(m["sm__cycles_elapsed.avg"]) / len(m["sm__cycles_elapsed.avg"]) / 1e6,
          sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]) / 1e6,
          sum(m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]) / len(m["sm__cycles_elapsed.avg"])))
PY
