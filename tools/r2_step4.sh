#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -s --timeout=300 --timeout-method=thread > gpurun_out/r2_suite.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|converged chain|additivity|relu masks|full-size relu" gpurun_out/r2_suite.log | tail -20
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_bench_n1.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=1 value %.0f frac %.3f e2e %.0f cpu %s clocks %s dtype %s" % (j["value"], j["roofline"]["frac"], j["e2e"]["value"], j["cpu_baseline"]["value"], j["clocks"]["sm_mhz"], j["dtype"][:60]))
    for k, v in (j.get("extra") or {}).items():
        print(" ", k, json.dumps(v)[:260])
except Exception as e:
    print("bench failed", e)
PY
tail -3 gpurun_out/r2_bench_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
