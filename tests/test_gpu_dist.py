"""GPU (>= 2 devices): the C-ABI NCCL wrappers (wesep_b200_nccl_*) — two processes, one GPU each, all-reduce of a flat buffer
through `GradAllReducer(direct=True)` equals the torch.distributed result.  Skipped on single-GPU boxes."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

WORKER = r"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
import torch.distributed as dist
from wesep_b200.distributed import GradAllReducer, init_from_env
rank, world, local = init_from_env("nccl")
dev = torch.device("cuda", local)
g = torch.Generator(device=dev).manual_seed(10 + rank)
flat = torch.randn(1_000_003, device=dev, generator=g)
ref = flat.clone()
dist.all_reduce(ref)
red = GradAllReducer(flat, n_buckets=3, direct=True)
assert red.comm is not None
red.all_reduce()
torch.cuda.synchronize()
err = float((flat - ref).abs().max())
assert err <= 1e-6, err
red.comm.close()
dist.barrier()
if rank == 0:
    print("NCCL_DIRECT_OK", err)
"""


def test_nccl_wrappers_world2(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, WESEP_DIST_TIMEOUT_S="120")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29621", str(script)], capture_output=True, text=True, timeout=300, env=env,
                         cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert out.returncode == 0 and "NCCL_DIRECT_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
