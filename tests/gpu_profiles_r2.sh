#!/bin/bash
# Round-2 profile capture (run under gpurun; everything lands in gpurun_out/, summaries are copied to profiles/ by hand)
set -x
mkdir -p gpurun_out
python tests/gpu_hbm_probe.py > gpurun_out/r2_hbm_probe.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"fakequant|a_self|select_hist" -c 60 --csv --log-file gpurun_out/r2_ncu_dram_membound.csv \
    python tests/gpu_hbm_probe.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv \
    python tests/gpu_ncu_step.py > gpurun_out/r2_ncu_step.log 2>&1
python tests/gpu_ncu_step.py 2 > gpurun_out/r2_step_live.log 2>&1
ADALOG_B200_LIN_FUSED=force ncu --set full --clock-control none --import-source on -k regex:lin_fused -c 1 \
    -o gpurun_out/r2_full_linf_i8 python tests/gpu_lin_bench.py small > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_fused -s 18 -c 1 \
    -o gpurun_out/r2_full_linf_log python tests/gpu_lin_bench.py small > /dev/null 2>&1
WSIDE=1 ncu --set full --clock-control none --import-source on -k regex:cand_gemm_err -c 1 \
    -o gpurun_out/r2_full_wside_i8 python tests/gpu_lin_bench.py small > /dev/null 2>&1
python tests/gpu_torchprof.py 3 128 deit_small_patch16_224 2>&1 | grep -v Warning > gpurun_out/r2_kernel_time_table.log
ls -la gpurun_out | tail -12
