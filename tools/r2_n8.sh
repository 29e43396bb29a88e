#!/bin/bash
# round 2, 8 GPUs: C4 SVGD 1/2/4/8 with the per-phase split (+ one A/B of the Gram all-reduce placement at 8), then the
# bench line with its sub-records at N = 8
timeout 400 python tools/bench_svgd_sharded.py --world 1,2,4,8 --steps 6 > gpurun_out/r2_svgd_c4_sharded.jsonl 2> gpurun_out/r2_svgd_c4_sharded.err
cat gpurun_out/r2_svgd_c4_sharded.jsonl | cut -c1-800; tail -3 gpurun_out/r2_svgd_c4_sharded.err
PYB_SVGD_GRAM_SYNC=1 timeout 200 python tools/bench_svgd_sharded.py --world 8 --steps 6 2>/dev/null | sed 's/^/gram_sync=1 /' | tee gpurun_out/r2_svgd_c4_sharded_gram_sync.jsonl | cut -c1-800
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/r2_bench_n8.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("N=8 value %.0f frac %.3f clocks %s" % (j["value"], j["roofline"]["frac"], j["clocks"]["sm_mhz"]))
    for k, v in (j.get("extra") or {}).items():
        print(" ", k, json.dumps(v)[:400])
except Exception as e:
    print("bench failed", e)
PY
tail -4 gpurun_out/r2_bench_n8.err
