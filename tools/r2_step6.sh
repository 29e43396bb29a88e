#!/bin/bash
timeout 90 python tools/dbg_mma.py 2>&1 | tail -3; test ${PIPESTATUS[0]} -eq 0 || { echo SMOKE FAILED; exit 1; }
FLAGS=0,2,4,8,16,32,62 timeout 300 python tools/timeline_fused.py 2>&1 | tail -8 | tee gpurun_out/r2_timeline.txt
