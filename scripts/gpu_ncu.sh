#!/usr/bin/env bash
# ncu evidence for profiles/: (1) launch list of the bench command, (2) --set full captures of the hot kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sgemm_tf32_kernel|split_tf32" -c 6 -o gpurun_out/prof_gemm \
    python scripts/profile_targets.py gemm > gpurun_out/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ew_flat_vec|ew_bcast2d|reduce_cols_kernel|reduce_rows_kernel" -c 14 -o gpurun_out/prof_ew \
    python scripts/profile_targets.py ew > gpurun_out/ncu_ew.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"reduce_rows_kernel|arg_rows_kernel" -c 4 -o gpurun_out/prof_reduce \
    python scripts/profile_targets.py reduce > gpurun_out/ncu_reduce.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_gemm.log
