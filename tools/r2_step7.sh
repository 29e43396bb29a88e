#!/bin/bash
timeout 90 python tools/dbg_mma.py 2>&1 | tail -3; test ${PIPESTATUS[0]} -eq 0 || { echo SMOKE FAILED; exit 1; }
FLAGS=${FLAGS:-0} timeout 300 python tools/timeline_fused.py 2>&1 | tail -8 | tee gpurun_out/r2_timeline.txt
tools/run_each.sh gpurun_out/r2_mma_tests.log 150 2 tests/test_gpu_i8.py
grep -E "mma epilogue vs|converged|step [0-9]+:" gpurun_out/r2_mma_tests.log | cut -c1-300
for em in 1 0; do
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --opt tc_epi_mma=$em > gpurun_out/r2_bench_mma_$em.json 2> gpurun_out/r2_bench_mma_$em.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_bench_mma_$em.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("epi_mma=$em value %.0f ms/step %.0f frac %.3f accept %.3f clocks %s" % (j["value"], j["ms_per_step"], j["roofline"]["frac"], j["accept_rate"], j["clocks"]["sm_mhz"]))
except Exception as e:
    print("bench failed", e)
PY
done
