#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_gpu.py -x -q --timeout=600 -p no:cacheprovider > gpurun_out/r2p_pytest_shard.log 2>&1; tail -5 gpurun_out/r2p_pytest_shard.log | cut -c1-300
for blocks in 4 16; do
 NB200_HOST_BLOCKS=$blocks NB200_HOST_TRACE=1 python scripts/host_probe.py 2> gpurun_out/r2p_host_blocks$blocks.log; grep -a "ms/call" gpurun_out/r2p_host_blocks$blocks.log
done
