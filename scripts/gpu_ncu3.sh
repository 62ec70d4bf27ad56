#!/usr/bin/env bash
# ncu --set full of the BF16x3 GEMM and its split pre-pass at 4096^3 (third launch of each: warm)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"sgemm_tf32_kernel|split_bf16" -s 4 -c 2 -o gpurun_out/prof_r1d \
    python scripts/profile_targets.py gemm_bf16 > gpurun_out/ncu_r1d.log 2>&1
tail -2 gpurun_out/ncu_r1d.log
