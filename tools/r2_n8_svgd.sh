#!/bin/bash
# 8 GPUs: C4 SVGD 1/2/4/8 with the per-phase split
timeout 400 python tools/bench_svgd_sharded.py --world ${WORLDS:-1,2,4,8} --steps 6 > gpurun_out/r2_svgd_c4_sharded.jsonl 2> gpurun_out/r2_svgd_c4_sharded.err
cat gpurun_out/r2_svgd_c4_sharded.jsonl | cut -c1-700; tail -3 gpurun_out/r2_svgd_c4_sharded.err
