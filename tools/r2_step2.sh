#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_i8.py tests/test_gpu_fullsize.py tests/test_gpu_dataset.py tests/test_gpu_tensor.py -m gpu -q -s --timeout=200 --timeout-method=thread > gpurun_out/r2_step2_tests.log 2>&1
grep -E "passed|failed|rc=|^FAILED|^ERROR|additivity|relu masks|full-size relu|unnormalised" gpurun_out/r2_step2_tests.log | tail -30
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --opt tc_i8=2 > gpurun_out/r2_bench_i8_2b.json 2> gpurun_out/r2_bench_i8_2b.err
python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_bench_i8_2b.json").read().strip().splitlines()[-1])
    print("tc_i8=2 value %.0f ms/step %.0f frac %.3f e2e %.0f accept %.3f clocks %s" % (j["value"], j["ms_per_step"], j["roofline"]["frac"], j["e2e"]["value"], j["accept_rate"], j["clocks"]["sm_mhz"]))
except Exception as e:
    print("bench failed", e)
PY
PYB_TC_I8=2 timeout 600 python -m pytest tests -m gpu -q --timeout=200 --timeout-method=thread > gpurun_out/r2_suite_i8_default.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_suite_i8_default.log | tail -30
