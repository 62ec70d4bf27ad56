#!/usr/bin/env bash
# One parameterised GPU run script (replaces the per-experiment gpu_r2*.sh files).
#   usage: gpu_stage.sh <tag> <stage> [<stage> ...]
#   stages: tests | tests_matmul | shard | bench | launches | ncu_matmul | ncu_hbm | probe:<precision number,...> (0 tf32x3, 2 bf16x3, 3 auto, 4 fp16x3, ...) | timeline:<precision,...> | memcheck | multi:<N>
tag=$1; shift
mkdir -p gpurun_out
for stage in "$@"; do
  case $stage in
    tests) timeout 900 python -m pytest tests -m gpu -x -q --timeout=600 -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${tag}_pytest_gpu.log | cut -c1-300 ;;
    tests_matmul) timeout 600 python -m pytest tests/test_gpu_parity.py -x -q --timeout=600 -p no:cacheprovider -k "host or matmul or graph" > gpurun_out/${tag}_pytest_matmul.log 2>&1; tail -4 gpurun_out/${tag}_pytest_matmul.log | cut -c1-300 ;;
    shard) timeout 600 python -m pytest tests/test_shard_gpu.py tests/test_multi_gpu.py -x -q --timeout=600 -p no:cacheprovider > gpurun_out/${tag}_pytest_shard.log 2>&1; tail -4 gpurun_out/${tag}_pytest_shard.log | cut -c1-300 ;;
    bench) timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; tail -3 gpurun_out/${tag}_bench_n1.err
       python scripts/bench_summary.py gpurun_out/${tag}_bench_n1.json ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_bench_under_ncu.log 2>&1; wc -l gpurun_out/${tag}_launches_bench.csv ;;
    ncu_matmul) bash scripts/gpu_ncu.sh ${tag}_matmul_auto "prep16|sgemm_tf32_kernel|fp16_post" 4 4 gemm_auto
       python scripts/ncu_extract.py gpurun_out/prof_${tag}_matmul_auto.ncu-rep gpurun_out/${tag}_ncu_matmul_auto.csv ;;
    ncu_merged) bash scripts/gpu_ncu.sh ${tag}_merged "sgemm_tf32_kernel" 1 6 gemm_bf16 gemm_fp16u
       python scripts/ncu_extract.py gpurun_out/prof_${tag}_merged.ncu-rep gpurun_out/${tag}_ncu_merged.csv ;;
    ncu_hbm) bash scripts/gpu_ncu.sh ${tag}_hbm "ew_flat_vec|ew_bcast2d|reduce_rows_kernel|arg_rows_kernel|reduce_cols" 0 14 ew reduce
       python scripts/ncu_extract.py gpurun_out/prof_${tag}_hbm.ncu-rep gpurun_out/${tag}_ncu_hbm.csv ;;
    probe:*) for prec in $(echo ${stage#probe:} | tr , ' '); do
         timeout 300 python scripts/gemm_probe.py child auto $prec 4096x4096x4096 8192x8192x8192 2048x2048x2048 1024x1024x1024 1000x520x776 > gpurun_out/${tag}_probe_$prec.jsonl 2> gpurun_out/${tag}_probe_$prec.err; cut -c1-330 gpurun_out/${tag}_probe_$prec.jsonl; tail -2 gpurun_out/${tag}_probe_$prec.err; done ;;
    timeline:*) for prec in $(echo ${stage#timeline:} | tr , ' '); do
         timeout 300 python scripts/gemm_timeline.py $prec 4096x4096x4096 8192x8192x8192 2048x2048x2048 > gpurun_out/${tag}_timeline_$prec.jsonl 2> gpurun_out/${tag}_timeline_$prec.err; cut -c1-1200 gpurun_out/${tag}_timeline_$prec.jsonl; tail -2 gpurun_out/${tag}_timeline_$prec.err; done ;;
    memcheck) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitizer_targets.py > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/${tag}_memcheck.log ;;
    prepbw) for bw in 100 150 200 250; do
         NB200_PREP_BW=$bw timeout 200 python scripts/gemm_timeline.py 5 4096x4096x4096 2048x2048x2048 > gpurun_out/${tag}_timeline_bw$bw.jsonl 2> gpurun_out/${tag}_timeline_bw$bw.err
         python - <<PY
import json
for ln in open("gpurun_out/${tag}_timeline_bw$bw.jsonl"):
    d = json.loads(ln); t = d["timeline_us"]
    print("bw $bw", d["M"], "period", round(d["ms_per_call_back_to_back"], 4), "A done", t["prep_phaseA_done"], "B1", t["prep_phaseB1_done"], "prep done", t["prep_last_cta_done"], "gemm", t["gemm_first_cta_past_wait"], "->", t["gemm_last_cta_done"])
PY
       done ;;
    synccheck) timeout 1500 compute-sanitizer --tool synccheck --error-exitcode 9 python scripts/sanitizer_targets.py > gpurun_out/${tag}_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/${tag}_synccheck.log ;;
    multi:*) bash scripts/gpu_multi.sh ${stage#multi:} $tag ;;
    *) echo "unknown stage $stage" ;;
  esac
done
