#!/bin/bash
# 2 GPUs: sharded SVGD tests on the tensor path + the C4 sharded timing (pipeline version)
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s --timeout=300 --timeout-method=thread > gpurun_out/r2_multi3_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|sharded \(pshard|Error|error" gpurun_out/r2_multi3_tests.log | tail -12
timeout 400 python tools/bench_svgd_sharded.py --world ${WORLDS:-1,2} --particles ${PARTICLES:-4096} --steps 4 > gpurun_out/r2_svgd_c4_sharded_12b.jsonl 2> gpurun_out/r2_svgd_c4_sharded_12b.err; cat gpurun_out/r2_svgd_c4_sharded_12b.jsonl | cut -c1-600; tail -3 gpurun_out/r2_svgd_c4_sharded_12b.err
