#!/bin/bash
# round 2: per-kernel SM cycles / time / DRAM + L2<->SM bytes of one 148-chain batch with int8 slices (tc_i8=2),
# then one --set full capture of the three tensor-core kernels
ncu --metrics sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum \
    --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r2_cycles_i8.csv \
    python bench.py --chains 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --opt tc_i8=2 > gpurun_out/r2_cycles_i8.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_cycles_i8.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
c = rows[h]; ki, mi, vi = c.index("Kernel Name"), c.index("Metric Name"), c.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[h + 1:]:
    try: agg[r[ki][:48]][r[mi]].append(float(r[vi].replace(",", "")))
    except ValueError: pass
tot = sum(sum(m["gpu__time_duration.sum"]) for m in agg.values())
for k, m in sorted(agg.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
    n = len(m["gpu__time_duration.sum"]); a = lambda x: sum(m[x]) / max(1, len(m[x]))
    print("%-50s n=%2d share=%5.1f%% cyc=%.3fM t=%.3fms tensor=%.1f%% dramR=%.2fGB dramW=%.2fGB l2->sm=%.2fGB sm->l2=%.2fGB" % (
        k, n, 100 * sum(m["gpu__time_duration.sum"]) / tot, a("sm__cycles_elapsed.avg") / 1e6, a("gpu__time_duration.sum") / 1e6,
        a("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), a("dram__bytes_read.sum") / 1e9, a("dram__bytes_write.sum") / 1e9,
        a("l1tex__m_xbar2l1tex_read_bytes.sum") / 1e9, a("l1tex__m_l1tex2xbar_write_bytes.sum") / 1e9))
PY
ncu --set full --clock-control none --import-source on -k regex:"tc_g" -s 9 -c 3 -f -o gpurun_out/r2_full_i8 \
    python bench.py --chains 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --opt tc_i8=2 > gpurun_out/r2_full_i8.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_full_i8.ncu-rep > gpurun_out/r2_ncu_full_i8.txt 2>&1; cat gpurun_out/r2_ncu_full_i8.txt
