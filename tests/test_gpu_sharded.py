"""Multi-GPU selector (NCCL): needs >= 2 GPUs on the box, skipped otherwise.  Launches
tools/sharded_check.py under torchrun and requires bit-identical results to the single-GPU path."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["planner", "fused", "nccl"])
def test_sharded_selector_equals_single_gpu(mode):
    """planner: the library shares only the problems where sharding pays (the others run on one GPU on every rank);
    fused / nccl: every problem is shared, over the peer-memory mailboxes / with one ncclAllGather per round."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 2 if n < 4 else 4
    env = dict(os.environ)
    if mode != "planner":
        env["BVIO_SEL_FORCE_SHARD"] = "1"
    if mode == "nccl":
        env["BVIO_SEL_NCCL"] = "1"
    port = {"planner": "29533", "fused": "29534", "nccl": "29535"}[mode]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "sharded_check.py")],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and "SHARDED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
