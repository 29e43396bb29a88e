#!/bin/bash
# small-width path (C1 / C2): parity tests that run on it, then the timings
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py -m gpu -q -x --timeout=300 2>&1 | tail -4
timeout 300 python tools/bench_small.py 2>&1 | cut -c1-420 | tee gpurun_out/r2_bench_small.jsonl
