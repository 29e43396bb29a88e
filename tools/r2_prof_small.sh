#!/bin/bash
timeout 120 python tools/prof_small.py || exit 1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fs_hmc -s 2 -c 1 -f -o gpurun_out/r2_full_small python tools/prof_small.py > gpurun_out/r2_full_small.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_full_small.ncu-rep > gpurun_out/r2_ncu_full_small.txt 2>&1; cat gpurun_out/r2_ncu_full_small.txt | head -30
