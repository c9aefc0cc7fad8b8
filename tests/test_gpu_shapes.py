"""GPU: shape robustness of the four model families against the fp64 oracles (which are pinned to goldens of the real reference):
one row, odd row counts, lengths that are not multiples of the hop, the shortest lengths the architectures accept — eval-mode
forward under no_grad (the whole-utterance inference path, infer.py:108-122) and, for one case each, the gradients."""
import numpy as np
import pytest
import torch

from oracle import bsrnn as ob
from oracle import dpccn as od
from oracle import tfgridnet as ot
from tests.test_gpu_kernels import check
from wesep_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _inputs(n, L, seed):
    b = synth.make_batch(n, T=L, Te=8, seed=seed)
    emb = torch.from_numpy(np.random.default_rng(seed + 1).standard_normal((n, 256)).astype(np.float32))
    return b["wav_mix"], emb


def _run(model, sd_fn, oracle_fn, n, L, seed, tol, grads):
    sd = sd_fn()
    synth.fill_state_dict_(sd, seed=seed)
    m = model.to(DEV)
    m.load_state_dict({k: v for k, v in sd.items()}, strict=True)
    mix, emb = _inputs(n, L, seed)
    sd64 = {k: v.double().requires_grad_(grads) for k, v in sd.items()}
    if grads:
        m.train()
        est = m(mix.to(DEV), emb.to(DEV))[0]
        ref = oracle_fn(sd64, mix.double(), emb.double())
        w = torch.from_numpy(np.random.default_rng(seed + 2).standard_normal((n, L)).astype(np.float32))
        (est * w.to(DEV)).sum().backward()
        (ref * w.double()).sum().backward()
        check("est", est.detach(), ref.detach().to(DEV), tol)
        worst = 0.0
        for k, p in m.named_parameters():
            g64 = sd64[k].grad
            if float(g64.norm()) < 1e-9:
                continue
            e = float((p.grad.double().cpu() - g64).norm() / g64.norm())
            worst = max(worst, e)
            assert e <= 3e-3, (k, e)
        return worst
    m.eval()
    with torch.no_grad():
        est = m(mix.to(DEV), emb.to(DEV))[0]
        ref = oracle_fn(sd64, mix.double(), emb.double())
    assert est.shape == (n, L)
    check("est", est, ref.to(DEV), tol)
    return 0.0


@pytest.mark.parametrize("n,L,grads", [(1, 1000, False), (3, 2049, False), (1, 640, False), (2, 1531, True)])
def test_bsrnn_shapes(n, L, grads):
    from wesep_b200.models import get_model
    from tests.test_oracle_bsrnn import _state_dict_like
    args = dict(spk_emb_dim=256, sr=16000, win=512, stride=128, use_spk_transform=False, joint_training=False, feature_dim=16,
                num_repeat=2, spk_fuse_type="multiply", multi_fuse=False)
    _run(get_model("BSRNN")(**args), lambda: _state_dict_like(args),
         lambda sd, x, e: ob.bsrnn_forward(sd, x, e, num_repeat=2, spk_fuse_type="multiply", multi_fuse=False), n, L, 100 + n, 2e-3, grads)


@pytest.mark.parametrize("n,L,grads", [(1, 4096, False), (3, 4100 + 77, False), (1, 5001, True)])
def test_dpccn_shapes(n, L, grads):
    """T = 1 + L // 128 >= 32 frames (the 32 x 32 average pooling of the pyramid needs one full window)."""
    from wesep_b200.models import get_model
    m = get_model("DPCCN")(win=512, stride=128, feature_dim=257, use_spk_transform=False, spk_fuse_type="multiply",
                           multi_fuse=False, joint_training=False, causal=False, tcn_blocks=2, tcn_layers=1)
    _run(m, lambda: od.make_state_dict(tcn_blocks=2, tcn_layers=1),
         lambda sd, x, e: od.dpccn_forward(sd, x, e, tcn_blocks=2, tcn_layers=1), n, L, 200 + n, 1e-3, grads)


@pytest.mark.parametrize("n,L,grads", [(1, 999, False), (3, 1601, False), (1, 130, False), (2, 777, True)])
def test_tfgridnet_shapes(n, L, grads):
    from wesep_b200.models import get_model
    kw = dict(n_layers=1, emb_dim=16, hidden=32, n_head=2, approx_qk_dim=260)
    m = get_model("TFGridNet")(n_srcs=1, n_fft=128, stride=64, window="hann", n_imics=1, n_layers=1, emb_dim=16, emb_ks=1, emb_hs=1,
                               lstm_hidden_units=32, attn_n_head=2, attn_approx_qk_dim=260, activation="prelu", eps=1e-5,
                               use_spk_transform=False, spk_fuse_type="multiply", joint_training=False)
    _run(m, lambda: ot.make_state_dict(**kw),
         lambda sd, x, e: ot.tfgridnet_forward(sd, x, e, n_layers=1, n_head=2), n, L, 300 + n, 2e-3, grads)
