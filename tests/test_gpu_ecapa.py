"""GPU: wespeaker ECAPA-TDNN speaker encoder on libwesep_b200 vs the fp64 oracle restatement (oracle/ecapa.py —
PARITY UNPINNED by the reference: wespeaker is an external package), its building blocks, and pDPCCN trained jointly with it."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ecapa as oe
from tests.test_gpu_dpccn import _act
from tests.test_gpu_kernels import check, rnd

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("K,dil,T", [(5, 1, 201), (3, 2, 97), (3, 4, 300)])
def test_conv1d_k(K, dil, T):
    """Conv1d(kernel K, dilation d, 'same' padding) = im2col1d + pointwise GEMM (+ ReLU epilogue): forward and all gradients."""
    from wesep_b200 import ops
    n, Ci, Co = 2, 12, 16
    x0, w0, b0 = rnd(n, Ci, T, seed=1), rnd(Co, Ci, K, seed=2, scale=0.3), rnd(Co, seed=3)
    x = _act(x0)
    w, b = w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    y = ops.conv1d_k(x, w, b, dil, act="relu")
    x64, w64, b64 = (t.double().requires_grad_(True) for t in (x0, w0, b0))
    r = F.relu(F.conv1d(x64, w64, b64, padding=dil * (K - 1) // 2, dilation=dil))
    g0 = rnd(n, Co, T, seed=4)
    y.backward(g0)
    r.backward(g0.double())
    check("y", y.detach(), r.detach(), 2e-5)
    check("gx", x.grad, x64.grad, 5e-5)
    check("gw", w.grad, w64.grad, 5e-5)
    check("gb", b.grad, b64.grad, 5e-5)


def test_astp_and_unary():
    from wesep_b200 import ops
    n, C, T = 3, 20, 157
    x0, a0 = rnd(n, C, T, seed=1), rnd(n, C, T, seed=2)
    x, a = _act(x0), _act(a0)
    alpha = ops.SoftmaxFn.apply(a, 1.0)
    out = ops.AstpFn.apply(x, alpha)
    g0 = rnd(n, 2 * C, seed=3)
    out.backward(g0)
    x64, a64 = x0.double().requires_grad_(True), a0.double().requires_grad_(True)
    al = torch.softmax(a64, 2)
    mean = (al * x64).sum(2)
    ref = torch.cat([mean, torch.sqrt(((al * x64 ** 2).sum(2) - mean ** 2).clamp(min=1e-10))], 1)
    ref.backward(g0.double())
    check("out", out.detach(), ref.detach(), 1e-5)
    check("gx", x.grad, x64.grad, 5e-5)
    check("ga", a.grad, a64.grad, 1e-4)
    for mode, fn in ((0, torch.relu), (1, torch.sigmoid)):
        v0 = rnd(7, 33, seed=5 + mode)
        v = v0.clone().requires_grad_(True)
        y = ops.UnaryFn.apply(v, mode)
        y.backward(torch.ones_like(y))
        v64 = v0.double().requires_grad_(True)
        fn(v64).sum().backward()
        check("unary", y.detach(), fn(v0.double()), 1e-6)
        check("unary grad", v.grad, v64.grad, 1e-6)


@pytest.mark.parametrize("glob", [True, False])
def test_ecapa_vs_oracle(glob):
    """ECAPA_TDNN(_GLOB)_c512 forward (train-mode BatchNorm) + every gradient vs the fp64 oracle."""
    from wesep_b200 import synth
    from wesep_b200.modules.speaker.resnet import get_speaker_model
    m = get_speaker_model("ECAPA_TDNN_GLOB_c512" if glob else "ECAPA_TDNN_c512")(feat_dim=80, embed_dim=192, pooling_func="ASTP")
    synth.fill_state_dict_(m.state_dict(), seed=11)
    sd64 = {k: v.detach().clone().double().requires_grad_(v.dtype.is_floating_point and "running" not in k)
            for k, v in m.state_dict().items()}
    m = m.to(DEV).train()
    feats = rnd(3, 120, 80, seed=5)
    emb = m(feats)
    assert emb.shape == (3, 192)
    ref = oe.ecapa_forward(sd64, feats.double().cpu(), global_context=glob, training=True)
    check("emb", emb.detach(), ref.detach().to(DEV), 2e-4)
    w = rnd(3, 192, seed=6)
    (emb * w).sum().backward()
    (ref * w.double().cpu()).sum().backward()
    worst = 0.0
    for k, p in m.named_parameters():
        g64 = sd64[k].grad
        if k == "pool.linear2.bias":           # constant over time in front of a softmax over time: the gradient is exactly 0
            assert float(p.grad.abs().max()) <= 1e-4
            continue
        gp = p.grad.double().cpu()
        e = float((gp - g64).norm() / (g64.norm() + 1e-12))
        worst = max(worst, e)
        # ReLU kinks: 29 conv -> ReLU -> BN layers; a pre-activation within fp32 round-off of 0 takes the other branch in the
        # fp64 oracle and moves single gradient elements (measured up to 6e-3 relative on a bias vector)
        assert e <= 2e-2, (k, e)
        assert float((gp * g64).sum() / (gp.norm() * g64.norm() + 1e-30)) >= 0.9998, k
    print("worst relative gradient error", worst)
    assert int(m.bn.num_batches_tracked) == 1 and float(m.layer1.bn.running_mean.abs().sum()) > 0


def test_dpccn_with_ecapa_constructs_and_steps():
    """dpccn.yaml with its ECAPA alternative (dpccn.yaml:59-64: spk_model ECAPA_TDNN_GLOB_c512, embed_dim 192, ASTP)."""
    from wesep_b200.models import get_model
    args = dict(win=512, stride=128, feature_dim=257, tcn_blocks=2, tcn_layers=1, causal=False, spk_fuse_type="multiply",
                use_spk_transform=False, multi_fuse=False, joint_training=True, spk_model="ECAPA_TDNN_GLOB_c512", spk_model_init=False,
                spk_args=dict(embed_dim=192, feat_dim=80, pooling_func="ASTP"), spk_emb_dim=192, spk_model_freeze=False,
                spk_feat=True, feat_type="consistent")
    m = get_model("DPCCN")(**args).to(DEV).train()
    g = torch.Generator().manual_seed(0)
    mix = (torch.randn(2, 4500, generator=g) * 0.1).to(DEV)
    fb = torch.randn(2, 90, 80, generator=g).to(DEV)
    est, emb = m(mix, fb)
    assert est.shape == (2, 4500) and emb.shape == (2, 192)
    est.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
