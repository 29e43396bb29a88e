#!/bin/bash
# round 2, 2 GPUs: the multi-GPU tests (sharded SVGD on the tensor path, multi-device API), the C4 sharded timing and the
# bench line with its sub-records at N = 2
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_api.py -m gpu -q -s --timeout=300 --timeout-method=thread > gpurun_out/r2_multi_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|sharded \(pshard" gpurun_out/r2_multi_tests.log | tail -12
timeout 400 python tools/bench_svgd_sharded.py --world 1,2 --steps 4 > gpurun_out/r2_svgd_c4_sharded_12.jsonl 2> gpurun_out/r2_svgd_c4_sharded_12.err; cat gpurun_out/r2_svgd_c4_sharded_12.jsonl | cut -c1-250; tail -3 gpurun_out/r2_svgd_c4_sharded_12.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_bench_n2.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=2 value %.0f frac %.3f e2e %.0f" % (j["value"], j["roofline"]["frac"], j["e2e"]["value"]))
    for k, v in (j.get("extra") or {}).items():
        print(" ", k, json.dumps(v)[:300])
except Exception as e:
    print("bench failed", e)
PY
tail -5 gpurun_out/r2_bench_n2.err
