#!/bin/bash
# Round-2 first GPU call for the operand-split prototypes: the gated tests, the un-profiled timing,
# then SM cycles, tensor-pipe activity and DRAM bytes of the three 1-CTA GEMM variants on the same shape and a --set full capture.
#   tools/proto_cycles.sh   -> gpurun_out/proto_tests.log, proto_timing.json, proto_cycles.csv (+ a summary on stdout)
PYB_TEST_I8=1 python -m pytest tests/test_gpu_tensor.py -k "int8 or mixed" -q -s > gpurun_out/proto_tests.log 2>&1
tail -3 gpurun_out/proto_tests.log
python tools/bench_mixed_proto.py 20 --i8 > gpurun_out/proto_timing.json 2> gpurun_out/proto_timing.err || exit 1
ncu --metrics sm__cycles_elapsed.avg,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum \
    --clock-control none -k regex:"tc_gemm_bf16x3" -s 2 -c 9 --csv --log-file gpurun_out/proto_cycles.csv \
    python tools/bench_mixed_proto.py 2 --i8 > gpurun_out/proto_cycles.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/proto_cycles.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
c = rows[h]; ki, mi, vi = c.index("Kernel Name"), c.index("Metric Name"), c.index("Metric Value")
agg = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[h + 1:]:
    agg[r[ki][-40:]][r[mi]].append(float(r[vi].replace(",", "")))
for k, m in agg.items():
    print(k, {n: round(sum(v) / len(v), 1) for n, v in m.items()})
PY
# L2 <-> SM bytes and stall reasons: one --set full capture of the same nine launches (they are ~0.1 ms each)
ncu --set full --clock-control none --import-source on -k regex:"tc_gemm_bf16x3" -c 9 -f -o gpurun_out/proto_full \
    python tools/bench_mixed_proto.py 2 --i8 > gpurun_out/proto_full.log 2>&1
python tools/ncu_summary.py full gpurun_out/proto_full.ncu-rep > gpurun_out/proto_full.txt 2>&1; tail -60 gpurun_out/proto_full.txt
