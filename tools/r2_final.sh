#!/bin/bash
# final 1-GPU validation of the round: smoke(), the whole GPU suite, the default bench line
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 --timeout-method=thread > gpurun_out/r2_final_tests.log 2>&1; tail -5 gpurun_out/r2_final_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_bench_default_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1 value %.0f e2e %.0f frac %.3f cpu %s clocks %s launches %s" % (j["value"], j["e2e"]["value"], j["roofline"]["frac"], j["cpu_baseline"]["value"], j["clocks"], j["gpu_launches"]))
    for k, v in (j.get("extra") or {}).items():
        print(" ", k, json.dumps(v)[:300])
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/r2_bench_default_n1.err
