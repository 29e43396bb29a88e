#!/bin/bash
# ncu --set full of (a) the C4 Gram / K*Y tcgen05 pair GEMM and the element passes around it, (b) the C1 iteration kernel
timeout 120 python tools/prof_c4.py 2>&1 | tail -2 || exit 1
timeout 400 ncu --set full --clock-control none -k regex:"tc_gemm_pair_bf16x3|k_stein_rhs_split_t|k_split_rows2|k_phi_finish_adam|k_d2_from_gram|k_kernel_rowsum" -s 12 -c 10 -f -o gpurun_out/r2_full_c4 python tools/prof_c4.py > gpurun_out/r2_full_c4.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_full_c4.ncu-rep > gpurun_out/r2_ncu_full_c4_kernels.txt 2>&1; grep -E "^==|duration|dram_throughput|tensor_cycles|issue_active" gpurun_out/r2_ncu_full_c4_kernels.txt | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fs_hmc -s 2 -c 1 -f -o gpurun_out/r2_full_small2 python tools/prof_small.py > gpurun_out/r2_full_small2.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_full_small2.ncu-rep > gpurun_out/r2_ncu_full_c1_fused_small.txt 2>&1; cat gpurun_out/r2_ncu_full_c1_fused_small.txt | head -20
