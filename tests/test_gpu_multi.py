"""Landmark-sharded solve on 2 GPUs (torchrun, NCCL) against the single-GPU result.  Skipped on boxes with one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tail", ["default", "split", "fused"])
def test_sharded_two_gpus_matches_single(tail):
    """`tail`: how the slab PCG over peer memory runs its recurrences - the default (one-launch cluster tail below 8192 unknowns, the exchange
    kernel + three-kernel tail above), always the three-kernel path, always the cluster tail (both inside CUDA graphs)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(here, "_multi_gpu_worker.py")]
    env = dict(os.environ)
    if tail == "split":
        env["G2OCU_PCG_TAIL"] = "split"
    elif tail == "fused":
        env["G2OCU_PCG_FUSED_MAX"] = "65536"
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_OK" in out.stdout
