#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider -k "matmul or sgemm or fp16" > gpurun_out/r2i_pytest_matmul.log 2>&1; tail -12 gpurun_out/r2i_pytest_matmul.log | cut -c1-300
for pdl in 1 0; do
  NB200_PDL=$pdl timeout 300 python scripts/gemm_timeline.py 3 4096x4096x4096 2048x2048x2048 1024x1024x1024 8192x8192x8192 4096x512x4096 > gpurun_out/r2i_timeline_pdl$pdl.jsonl 2> gpurun_out/r2i_timeline_pdl$pdl.err
  tail -3 gpurun_out/r2i_timeline_pdl$pdl.err
  python - <<PY
import json
for l in open("gpurun_out/r2i_timeline_pdl$pdl.jsonl"):
    d = json.loads(l); t = d["timeline_us"]
    t2 = d["timeline_in_loop_us"]; print("   in loop: prep_end", t2["prep_last_cta_done"], "gemm", t2["gemm_first_cta_past_wait"], t2["gemm_last_cta_done"], "fb_done", t2["fallback_done"], "next", t2["next_call_prep_start"]); print("pdl=$pdl", d["M"], d["K"], d["N"], "ms/call", round(d["ms_per_call_back_to_back"], 4), "A", t["prep_phaseA_done"], "B1", t["prep_phaseB1_done"], "bar", t["prep_barrier_passed"], "prep_end", t["prep_last_cta_done"], "gemm", t["gemm_first_cta_enter"], t["gemm_first_cta_past_wait"], t["gemm_last_cta_done"], "post", t["post_enter"], t["post_done"], "fb", t["fallback_enter"], t["fallback_done"])
PY
done
