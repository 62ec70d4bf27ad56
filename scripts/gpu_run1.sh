#!/usr/bin/env bash
# first GPU bring-up: non-GEMM parity, then GEMM probe (one variant per subprocess), then GEMM tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; ls /root/reference >> gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout=300 -k "not matmul and not dot and not config2" -p no:cacheprovider > gpurun_out/pytest_nogemm.log 2>&1
echo "pytest_nogemm exit $?" >> gpurun_out/gpu.txt
timeout 1500 python scripts/gemm_probe.py > gpurun_out/gemm_probe.log 2>&1
echo "gemm_probe exit $?" >> gpurun_out/gpu.txt
timeout 600 python -m pytest tests -m gpu -q --timeout=300 -k "matmul or dot or config2" -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1
echo "pytest_gemm exit $?" >> gpurun_out/gpu.txt
tail -5 gpurun_out/pytest_nogemm.log; cat gpurun_out/gemm_probe.log | tail -20; tail -5 gpurun_out/pytest_gemm.log
