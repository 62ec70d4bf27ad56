#!/usr/bin/env bash
mkdir -p gpurun_out
for bw in 100 130 150 170 200; do
  NB200_PREP_BW=$bw timeout 300 python scripts/gemm_timeline.py 3 4096x4096x4096 2048x2048x2048 > gpurun_out/r2j_timeline_bw$bw.jsonl 2> gpurun_out/r2j_timeline_bw$bw.err
  tail -3 gpurun_out/r2j_timeline_bw$bw.err
  python - <<PY
import json
for l in open("gpurun_out/r2j_timeline_bw$bw.jsonl"):
    d = json.loads(l); t = d["timeline_us"]; t2 = d["timeline_in_loop_us"]
    print("bw=$bw", d["M"], "ms/call", round(d["ms_per_call_back_to_back"], 4), "A", t["prep_phaseA_done"], "B1", t["prep_phaseB1_done"], "prep_end", t["prep_last_cta_done"], "| loop: prep_end", t2["prep_last_cta_done"], "gemm", t2["gemm_first_cta_past_wait"], t2["gemm_last_cta_done"], "next", t2["next_call_prep_start"])
PY
done
