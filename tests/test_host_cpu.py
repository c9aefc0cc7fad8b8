"""CPU: host logic, C-ABI surface, and the world_size-2 gloo path of the gradient all-reduce."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    from wesep_b200 import _lib
    L = _lib.lib()
    assert L.wesep_b200_version() == 100
    hdr = open(os.path.join(ROOT, "include", "wesep_b200.h")).read()
    declared = set(re.findall(r"\b(wesep_b200_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(L, name), name
    assert set(_lib.FUNCTIONS) == declared


def test_struct_layout_matches_c(tmp_path):
    """sizeof() of every ctypes struct generated from the header == the C compiler's."""
    from wesep_b200 import _lib
    names = list(_lib.STRUCTS)
    src = '#include "wesep_b200.h"\n#include <stdio.h>\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    c = tmp_path / "sz.c"
    c.write_text(src)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    for n in names:
        assert int(out[n]) == ctypes.sizeof(_lib.STRUCTS[n]), n


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of silently computing elsewhere."""
    from wesep_b200 import ops
    with pytest.raises(RuntimeError):
        ops.sisdr_losses([torch.zeros(2, 100)], torch.zeros(2, 100))
    with pytest.raises(RuntimeError):
        ops.conv1x1(torch.zeros(1, 8, 16), torch.zeros(4, 8))
    from wesep_b200.utils.optim import ParamArena
    with pytest.raises(RuntimeError):
        ParamArena([torch.nn.Parameter(torch.zeros(3))])


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wesep_b200")):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_model_registry_and_state_dict_contract():
    from oracle import spexplus as ospex
    from wesep_b200.models import get_model
    cfg = dict(ospex.DEFAULT_CFG)
    cfg.update(B=64, H=128, X=3, R=2)
    m = get_model("ConvTasNet")(N=256, L=20, B=64, H=128, P=3, X=3, R=2, spk_emb_dim=256, norm="gLN", activate="relu",
                                causal=False, skip_con=False, spk_fuse_type="concatConv", multi_fuse=True,
                                use_spk_transform=False, encoder_type="Multi", decoder_type="Multi", joint_training=True,
                                multi_task=True, spksInTrain=251)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == ospex.state_dict_spec(cfg)
    with pytest.raises(NotImplementedError):
        get_model("BSRNN_Feats")
    with pytest.raises(RuntimeError):                  # pBSRNN: CUDA only (no fallback)
        get_model("BSRNN")(joint_training=False, use_spk_transform=False, feature_dim=16, num_repeat=1,
                           spk_fuse_type="multiply")(torch.zeros(1, 2000), torch.zeros(1, 256))
    with pytest.raises(NotImplementedError):           # wespeaker's ResNet18 / 34 and ECAPA-TDNN families are built, nothing else
        get_model("BSRNN")(joint_training=True, use_spk_transform=False, spk_feat=True, spk_model="CAMPPlus",
                           spk_args=dict(feat_dim=80, embed_dim=192))
    e = get_model("BSRNN")(joint_training=True, use_spk_transform=False, spk_feat=True, spk_model="ECAPA_TDNN_GLOB_c512",
                           spk_args=dict(feat_dim=80, embed_dim=192, pooling_func="ASTP"), spk_emb_dim=192, feature_dim=16, num_repeat=1)
    assert sum(p.numel() for p in e.spk_model.parameters()) == 6190720      # wespeaker's published size of ECAPA_TDNN_GLOB_c512
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 100), torch.zeros(1, 100))     # >= 3-D input, convtasnet.py:163-166


def test_scheduler_matches_golden():
    """Closed-form schedule vs learning rates written by the reference's own ExponentialDecrease (tests/golden/sched.npz)."""
    import numpy as np
    from wesep_b200.utils.lr import exponential_decrease_lr
    z = np.load(os.path.join(ROOT, "tests", "golden", "sched.npz"))
    for it, lr in zip(z["its"], z["lrs"]):
        got = exponential_decrease_lr(int(it), 150 * 1000, 1e-3, 2.5e-5)
        assert abs(got - lr) <= 1e-12 + 1e-9 * lr


def test_shard_rows():
    from wesep_b200.distributed import shard_rows
    for n, w in ((32, 8), (10, 4), (3, 8)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from wesep_b200.distributed import init_from_env, GradAllReducer, broadcast_params
rank, world, _ = init_from_env(backend="gloo")
torch.manual_seed(rank)
flat = torch.arange(1003, dtype=torch.float32) * (rank + 1)
red = GradAllReducer(flat, n_buckets=3)
red.all_reduce()
expect = torch.arange(1003, dtype=torch.float32) * sum(r + 1 for r in range(world))
assert torch.equal(flat, expect), (rank, flat[:4])
assert abs(red.grad_scale - 1.0 / world) < 1e-12
p = torch.full((17,), float(rank))
broadcast_params(p)
assert float(p.sum()) == 0.0
dist.barrier()
print("rank", rank, "ok")
"""


def test_gloo_world2_allreduce(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(w), ROOT],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_dpccn_state_dict_contract():
    """DPCCN constructs on the host with the reference's keys, order and shapes (oracle.dpccn.make_state_dict restates
    wesep/models/dpccn.py:58-204), rejects configurations that are not built, and refuses CPU tensors."""
    import pytest
    import torch
    from oracle import dpccn as od
    from wesep_b200.models import get_model
    m = get_model("DPCCN")(joint_training=False, tcn_blocks=3, tcn_layers=1)
    ref = od.make_state_dict(tcn_blocks=3, tcn_layers=1)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(sd[k].shape == ref[k].shape for k in ref)
    with pytest.raises(NotImplementedError):
        get_model("DPCCN")(joint_training=False, spk_fuse_type="concat")
    with pytest.raises(NotImplementedError):
        get_model("DPCCN")(joint_training=False, causal=True)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4000), torch.zeros(1, 256))
