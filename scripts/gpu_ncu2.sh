#!/usr/bin/env bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"ew_flat_vec|ew_bcast2d|split_tf32" -c 10 -o gpurun_out/prof_r1c \
    python scripts/profile_targets.py gemm ew > gpurun_out/ncu_r1c.log 2>&1
tail -2 gpurun_out/ncu_r1c.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 3 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
