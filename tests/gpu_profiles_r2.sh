#!/bin/bash
# Round-2 profile capture (run under gpurun; everything lands in gpurun_out/, summaries are made here with
# tests/ncu_summarise.py and copied to profiles/ by hand).  Never a bench number: everything below runs under ncu or
# torch.profiler except the two probes.
set -x
mkdir -p gpurun_out
./tests/microbench/tmem_ld_bench > gpurun_out/r2_tmem_ld_bench.log 2>&1
./tests/microbench/mma_n_bench > gpurun_out/r2_mma_n_bench.log 2>&1
python tests/gpu_hbm_probe.py > gpurun_out/r2_hbm_probe.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"fakequant|a_self|select_hist" -c 60 --csv --log-file gpurun_out/r2_ncu_dram_membound.csv \
    python tests/gpu_hbm_probe.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv \
    python tests/gpu_ncu_step.py > gpurun_out/r2_ncu_step.log 2>&1
python tests/gpu_ncu_step.py 2 > gpurun_out/r2_step_live.log 2>&1
# the three tcgen05 kernels, one launch each (DeiT-S shapes): lin_fused int8 (qkv), lin_fused AdaLog (fc2), W-side int8
ncu --set full --clock-control none --import-source on -k regex:lin_fused -c 1 \
    -o gpurun_out/r2_full_linf_i8 python tests/gpu_lin_bench.py small > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:lin_fused -s 18 -c 1 \
    -o gpurun_out/r2_full_linf_log python tests/gpu_lin_bench.py small > /dev/null 2>&1
WSIDE=1 ncu --set full --clock-control none --import-source on -k regex:cand_gemm_err -c 1 \
    -o gpurun_out/r2_full_wside_i8 python tests/gpu_lin_bench.py small > /dev/null 2>&1
SKIP_TILE=1 ncu --set full --clock-control none --import-source on -k regex:fused_cand_gemm -c 1 \
    -o gpurun_out/r2_full_fattn_qk python tests/gpu_diag.py matmul > /dev/null 2>&1
python tests/gpu_lin_bench.py small > gpurun_out/r2_lin_bench_small.log 2>&1
python tests/gpu_lin_bench.py base > gpurun_out/r2_lin_bench_base.log 2>&1
SKIP_TILE=1 python tests/gpu_diag.py matmul > gpurun_out/r2_attention_bench.log 2>&1
python tests/gpu_torchprof.py 3 128 deit_small_patch16_224 2>&1 | grep -v Warning > gpurun_out/r2_kernel_time_table.log
python tests/gpu_torchprof.py 4 128 deit_base_patch16_224 2>&1 | grep -v Warning > gpurun_out/r2_kernel_time_table_deit_base.log
ls -la gpurun_out | tail -16
