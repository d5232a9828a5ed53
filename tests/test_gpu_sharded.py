"""BASELINE config 4 on real devices: one registration with its queries sharded over 2 GPUs, the 28-value exchange fused
into the persistent kernel (P2P stores into cudaIpc-mapped mailboxes) and, for comparison, as an NCCL all-reduce.
Skipped unless the box has at least 2 GPUs (the driver's single-GPU tier skips it; `gpurun --gpus 2` runs it)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
def test_sharded_registration_two_gpus():
    env = dict(os.environ, LLB_SENSOR="hdl32e", LLB_RAW_CORNER="300000", LLB_RAW_SURF="200000")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "sharded_check.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == 2 and r["pose_match"] and r["bit_identical_across_ranks"] and r["fused_matches_nccl_bitwise"]
    assert r["iterations_sharded"] == r["iterations_single"]
