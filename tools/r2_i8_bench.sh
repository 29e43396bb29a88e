#!/bin/bash
# round 2: the C3 headline with the three operand splits, then the whole GPU suite with int8 slices as the default
tools/run_each.sh gpurun_out/i8_stress.log 100 1 tests/test_gpu_i8.py::test_int8_slices_zero_weights_and_unnormalised_data
grep "unnormalised" gpurun_out/i8_stress.log
for m in 0 1 2; do
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --opt tc_i8=$m > gpurun_out/r2_bench_i8_$m.json 2> gpurun_out/r2_bench_i8_$m.err
  python - <<PY
import json
try:
    j = json.loads(open("gpurun_out/r2_bench_i8_$m.json").read().strip().splitlines()[-1])
    print("tc_i8=$m value %.0f ms/step %.0f frac %.3f share %.3f accept %.3f clocks %s" % (j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel_share_of_step"], j["accept_rate"], j["clocks"]["sm_mhz"]))
except Exception as e:
    print("tc_i8=$m failed", e)
PY
done
PYB_TC_I8=2 timeout 600 python -m pytest tests -m gpu -q -x --timeout=120 --timeout-method=thread > gpurun_out/r2_suite_i8_default.log 2>&1
tail -15 gpurun_out/r2_suite_i8_default.log
