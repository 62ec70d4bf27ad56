"""GPU test of the N > 1 path on real devices (NCCL over NVLink): needs >= 2 GPUs, skipped otherwise.
Spawns torchrun with scripts/multi_gpu_check.py (2 ranks) and checks its verdict line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_sharded_matmul_and_reductions_over_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, NB200_CHECK_BATCH_PER_RANK="4")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29577", os.path.join(ROOT, "scripts", "multi_gpu_check.py")], capture_output=True, text=True,
                       timeout=550, env=env)
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0 and lines, p.stderr[-2000:]
    r = json.loads(lines[-1])
    assert r["matmul_ok"] and r["sum_ok"] and r["argmax_ok"], r
