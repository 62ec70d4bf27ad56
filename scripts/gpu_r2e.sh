#!/usr/bin/env bash
mkdir -p gpurun_out
for pdl in 1 0; do
  NB200_PDL=$pdl timeout 300 python scripts/gemm_timeline.py 3 4096x4096x4096 2048x2048x2048 1024x1024x1024 8192x8192x8192 > gpurun_out/r2e_timeline_pdl$pdl.jsonl 2> gpurun_out/r2e_timeline_pdl$pdl.err
  tail -3 gpurun_out/r2e_timeline_pdl$pdl.err; cat gpurun_out/r2e_timeline_pdl$pdl.jsonl | cut -c1-900
done
