#!/usr/bin/env bash
# usage: gpu_ncu.sh <tag> <kernel regex> <skip> <count> <profile_targets args...>   -> gpurun_out/prof_<tag>.ncu-rep (+ raw csv)
tag=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s $skip -c $count -f -o gpurun_out/prof_$tag \
    python scripts/profile_targets.py "$@" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/prof_$tag.ncu-rep gpurun_out/prof_${tag}_raw.csv
