#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py tests/test_gpu_tensor.py tests/test_gpu_reference_goldens.py -m gpu -q -k "svgd or sharded or predictive" --timeout=300 --timeout-method=thread > gpurun_out/r2_multi2_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/r2_multi2_tests.log | tail -8
timeout 300 python tools/bench_svgd_sharded.py --world 1,2 --steps 4 2>/dev/null | grep "^{" | cut -c1-700
