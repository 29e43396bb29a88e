#!/bin/bash
timeout 90 python tools/dbg_mma.py 2>&1 | tail -3; test ${PIPESTATUS[0]} -eq 0 || { echo SMOKE FAILED; exit 1; }
timeout 200 python tools/timeline_fused.py 2>&1 | tail -4 | tee gpurun_out/r2_timeline.txt
