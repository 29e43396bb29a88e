#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:"tc_fused_i8_mma" -s 3 -c 1 -f -o gpurun_out/r2_full_mma \
    python bench.py --chains 148 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r2_full_mma.log 2>&1
python tools/ncu_summary.py full gpurun_out/r2_full_mma.ncu-rep > gpurun_out/r2_ncu_full_mma.txt 2>&1; cat gpurun_out/r2_ncu_full_mma.txt
ncu -i gpurun_out/r2_full_mma.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[2]
for k in ['lts__t_bytes.sum.per_second','lts__t_sectors.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__m_l1tex2xbar_write_bytes.sum','smsp__inst_executed.sum','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed']:
    for i,n in enumerate(h):
        if n==k: print(k, r[i])
for i,n in enumerate(h):
    if 'pipe' in n and 'pct_of_peak_sustained_active' in n and 'inst_executed' in n: print(n, r[i])
"
