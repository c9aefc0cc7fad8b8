"""GPU parity tests: every CUDA entry point (through the C-ABI via wesep_b200.ops) against the
oracle (plain torch, fp64) on the same seeded inputs.  Run on the B200 box: pytest -m gpu."""
import math

import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import optim as ooptim
from oracle import spexplus as ospex
from tests.util import rel_l2
from wesep_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from wesep_b200 import ops
    return ops


def check(name, got, ref, tol):
    e = rel_l2(got, ref)
    assert e <= tol, f"{name}: rel_l2 {e:.3e} > {tol:.1e} (|ref| {float(ref.double().norm()):.3e})"
    return e


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(*shape, generator=g)).to(DEV)


# --------------------------------------------------------------------------- SI-SDR
@pytest.mark.parametrize("n,L", [(1, 64000), (2, 48000), (5, 63999), (32, 8000)])
def test_sisdr_fwd_bwd(n, L):
    ops = _ops()
    rng = np.random.default_rng(L + n)
    t = torch.from_numpy((0.1 * rng.standard_normal((n, L)) + 0.03).astype(np.float32)).to(DEV)   # DC offset
    ests = []
    for i, snr in enumerate((-30.0, 5.0, 60.0)):
        e = torch.from_numpy(rng.standard_normal((n, L)).astype(np.float32)).to(DEV)
        e = e * (t.norm(dim=1, keepdim=True) / e.norm(dim=1, keepdim=True)) * 10 ** (-snr / 20)
        ests.append((0.7 * t + e + 0.01 * i).requires_grad_(True))
    losses, rows = ops.sisdr_losses(ests, t)
    w = torch.tensor([0.8, 0.1, 0.1], device=DEV)
    (losses * w).sum().backward()
    ests64 = [e.detach().double().requires_grad_(True) for e in ests]
    ref_rows = torch.stack([olosses.sisdr_per_row(e, t.double()) for e in ests64])
    ref_losses = torch.stack([olosses.sisdr_loss(e, t.double()) for e in ests64])
    (ref_losses * w.double()).sum().backward()
    dmax = float((rows.double() - ref_rows).abs().max())
    assert dmax <= 1e-3, f"per-row SI-SDR differs by {dmax:.2e} dB"
    assert float((losses.double() - ref_losses).abs().max()) <= 1e-3
    for i in range(3):
        check(f"sisdr grad est{i}", ests[i].grad, ests64[i].grad, 2e-4)


def test_sisdr_zero_target_row_and_single():
    ops = _ops()
    t = rnd(3, 4000, seed=1, scale=0.1)
    t[1].zero_()                      # eps path: all-zero target row
    x = rnd(3, 4000, seed=2, scale=0.1).requires_grad_(True)
    losses, rows = ops.sisdr_losses([x], t)
    losses[0].backward()
    x64 = x.detach().double().requires_grad_(True)
    ref = olosses.sisdr_loss(x64, t.double())
    ref.backward()
    assert torch.isfinite(rows).all() and torch.isfinite(x.grad).all()
    assert abs(float(losses[0]) - float(ref)) <= 1e-3
    check("grad", x.grad, x64.grad, 1e-3)


# --------------------------------------------------------------------------- clip + Adam
def test_clip_adam_matches_oracle_and_golden():
    from wesep_b200.utils.optim import FusedClipAdam
    z = np.load("tests/golden/optim.npz")
    n = 5
    params = [torch.nn.Parameter(torch.from_numpy(z[f"p0_{i}"].copy()).to(DEV)) for i in range(n)]
    opt = FusedClipAdam(params, lr=1e-3, weight_decay=1e-4, clip=5.0)
    for step in range(3):
        opt.zero_grad()
        for i, p in enumerate(params):
            p.grad.copy_(torch.from_numpy(z[f"g{step}_{i}"]).to(DEV))
        opt.param_groups[0]["lr"] = float(z["lrs"][step])
        norms = opt.step()
        assert np.allclose(norms.cpu().numpy(), z[f"norms{step}"], rtol=1e-5), "per-tensor norms"
        for i, p in enumerate(params):
            ref = torch.from_numpy(z[f"p{step + 1}_{i}"])
            assert torch.allclose(p.detach().cpu(), ref, rtol=3e-5, atol=1e-7), f"param {i} after step {step + 1}"


def test_clip_adam_large_random():
    from wesep_b200.utils.optim import FusedClipAdam
    shapes = [(512, 256, 1), (512,), (1,), (512, 1), (512, 1, 3), (251, 256), (9001,)]
    g = torch.Generator().manual_seed(5)
    p0 = [torch.randn(*s, generator=g) for s in shapes]
    params = [torch.nn.Parameter(p.clone().to(DEV)) for p in p0]
    opt = FusedClipAdam(params, lr=1e-3, weight_decay=1e-4, clip=5.0)
    ref_p = [p.clone().double() for p in p0]
    m = [torch.zeros_like(p) for p in ref_p]
    v = [torch.zeros_like(p) for p in ref_p]
    for step in range(1, 4):
        grads = [torch.randn(*s, generator=g) * (10.0 if i % 2 else 0.01) for i, s in enumerate(shapes)]
        opt.zero_grad()
        for p, gr in zip(params, grads):
            p.grad.copy_(gr.to(DEV))
        opt.grad_scale = 0.5
        opt.step()
        g64 = [0.5 * gr.double() for gr in grads]
        ooptim.clip_gradients(g64, 5.0)
        ooptim.adam_step(ref_p, g64, m, v, step, 1e-3)
        for i, p in enumerate(params):
            check(f"param{i} step{step}", p.detach().cpu(), ref_p[i], 1e-5)


# --------------------------------------------------------------------------- GEMMs
@pytest.mark.parametrize("n,Kd,M,T", [(2, 64, 128, 1000), (1, 256, 512, 6399), (3, 20, 256, 130), (2, 512, 256, 999),
                                      (2, 160, 768, 517)])
@pytest.mark.parametrize("w_trans", [False, True])
def test_conv1x1_fwd_bwd(n, Kd, M, T, w_trans):
    ops = _ops()
    x = ops.new_act(n, Kd, T, DEV)
    x.copy_(rnd(n, Kd, T, seed=1))
    x.requires_grad_(True)
    W = rnd(Kd, M, seed=2, scale=1 / math.sqrt(Kd)) if w_trans else rnd(M, Kd, seed=2, scale=1 / math.sqrt(Kd))
    W.requires_grad_(True)
    b = rnd(M, seed=3).requires_grad_(True)
    for act in (None, "relu"):
        y = ops.conv1x1(x, W, b, w_trans=w_trans, act=act)
        gy = rnd(n, M, T, seed=4)
        gx, gW, gb = torch.autograd.grad(y, (x, W, b), gy)
        x64, W64, b64 = (t.detach().double().requires_grad_(True) for t in (x, W, b))
        Wm = W64.t() if w_trans else W64
        y64 = torch.einsum("mk,nkt->nmt", Wm, x64) + b64[None, :, None]
        if act == "relu":   # pin the branch at |y| ~ 0 to the CUDA result (ReLU's derivative jumps there)
            y64 = torch.where(y.detach() > 0, y64, torch.zeros_like(y64))
        rx, rW, rb = torch.autograd.grad(y64, (x64, W64, b64), gy.double())
        check("y", y, y64, 1e-5)
        check("dx", gx, rx, 1e-5)
        check("dW", gW, rW, 1e-5)
        check("db", gb, rb, 1e-5)


def test_conv1x1_tf32_mode_is_close():
    from wesep_b200 import _lib
    ops = _ops()
    x = ops.new_act(2, 256, 2000, DEV)
    x.copy_(rnd(2, 256, 2000, seed=1))
    W = rnd(512, 256, seed=2, scale=1 / 16)
    try:
        _lib.set_gemm_mode(1)
        y1 = ops.conv1x1(x, W)
    finally:
        _lib.set_gemm_mode(0)
    y64 = torch.einsum("mk,nkt->nmt", W.double(), x.double())
    e = rel_l2(y1, y64)
    assert 1e-5 < e < 2e-3, f"single-pass TF32 error {e:.2e} outside the expected band"


# --------------------------------------------------------------------------- cLN
@pytest.mark.parametrize("n,C,T", [(2, 768, 1000), (1, 96, 37), (3, 768, 131)])
def test_cln(n, C, T):
    ops = _ops()
    x = ops.new_act(n, C, T, DEV)
    x.copy_(torch.relu(rnd(n, C, T, seed=1)) + 0.1)
    x.requires_grad_(True)
    g = (1 + 0.1 * rnd(C, seed=2)).requires_grad_(True)
    b = (0.1 * rnd(C, seed=3)).requires_grad_(True)
    y = ops.cln(x, g, b)
    gy = rnd(n, C, T, seed=4)
    gx, gg, gb = torch.autograd.grad(y, (x, g, b), gy)
    x64, g64, b64 = (t.detach().double().requires_grad_(True) for t in (x, g, b))
    y64 = ospex.cln(x64, g64, b64)
    rx, rg, rb = torch.autograd.grad(y64, (x64, g64, b64), gy.double())
    check("y", y, y64, 2e-6)
    check("dx", gx, rx, 5e-6)
    check("dgamma", gg, rg, 1e-5)
    check("dbeta", gb, rb, 1e-5)


# --------------------------------------------------------------------------- TCN blocks
def _block_case(fuse, n, B, H, T, dil, seed, alpha=None, E=32):
    from wesep_b200.modules.tasnet.convs import Conv1DBlock, Conv1DBlock4Fuse
    ops = _ops()
    if fuse:
        blk = Conv1DBlock4Fuse(in_channels=B, spk_embed_dim=E, conv_channels=H, kernel_size=3, dilation=dil, norm="gLN")
    else:
        blk = Conv1DBlock(B, H, 3, dil, "gLN", False, False)
    synth.fill_state_dict_(blk.state_dict(), seed=seed)
    if alpha is not None:
        with torch.no_grad():
            for k, p in blk.named_parameters():
                if "prelu" in k.lower():
                    p.fill_(alpha)
    blk = blk.to(DEV)
    x = ops.new_act(n, B, T, DEV)
    x.copy_(rnd(n, B, T, seed=seed + 1))
    x.requires_grad_(True)
    aux = rnd(n, E, 1, seed=seed + 2).requires_grad_(True) if fuse else None
    out = blk(x, aux) if fuse else blk(x)
    gy = rnd(n, B, T, seed=seed + 3)
    params = list(blk.parameters())
    ins = [x] + ([aux] if fuse else [])
    ops.DEBUG_STASH = {}
    grads = torch.autograd.grad(out, ins + params, gy)
    stash, ops.DEBUG_STASH = ops.DEBUG_STASH, None
    # PReLU's derivative jumps at 0: an element with |pre-activation| ~ 1e-7 takes different branches in
    # fp32 and fp64 and moves one gradient element by O(1).  Pin the oracle's branches to the ones the
    # CUDA path took (its saved pre-activations u, d), so the comparison tests the arithmetic.
    masks = [stash["u"] > 0, stash["d"] > 0]

    def prelu_pinned(v, a):
        m = masks.pop(0)
        return torch.where(m, v, a * v)

    sd64 = {k: v.detach().double().requires_grad_(v.is_floating_point()) for k, v in blk.state_dict().items()}
    x64 = x.detach().double().requires_grad_(True)
    if fuse:
        a64 = aux.detach().double().requires_grad_(True)
        o64 = ospex.conv1d_block4fuse(sd64, "", x64, a64, dil, prelu=prelu_pinned)
        ins64 = [x64, a64]
    else:
        o64 = ospex.conv1d_block(sd64, "", x64, dil, prelu=prelu_pinned)
        ins64 = [x64]
    names = [k for k, _ in blk.named_parameters()]
    ref = torch.autograd.grad(o64, ins64 + [sd64[k] for k in names], gy.double())
    errs = {"out": check("out", out, o64, 3e-6)}
    labels = ["dx"] + (["daux"] if fuse else []) + names
    for lab, g, r in zip(labels, grads, ref):
        errs[lab] = check(lab, g, r, 3e-5 if lab == "dx" else 2e-4)
    return errs


@pytest.mark.parametrize("dil", [1, 2, 4, 8, 16, 32, 64, 128])
def test_tcn_block_dilations(dil):
    _block_case(False, n=2, B=64, H=128, T=1000, dil=dil, seed=10 + dil)


@pytest.mark.parametrize("T", [17, 130, 2049, 4799])
def test_tcn_block_lengths(T):
    _block_case(False, n=3, B=64, H=128, T=T, dil=4, seed=3)


@pytest.mark.parametrize("alpha", [0.25, -0.1, 0.0])
def test_tcn_block_prelu_slopes(alpha):
    _block_case(False, n=1, B=64, H=128, T=700, dil=2, seed=5, alpha=alpha)


def test_tcn_block_full_size():
    """Spex+ shapes: B=256, H=512, K=6399 frames (4 s @ 16 kHz), dilation 128."""
    _block_case(False, n=2, B=256, H=512, T=6399, dil=128, seed=7)


@pytest.mark.parametrize("T,E", [(1000, 32), (6399, 256)])
def test_tcn_fuse_block(T, E):
    B, H = (256, 512) if T == 6399 else (64, 128)
    _block_case(True, n=2, B=B, H=H, T=T, dil=1, seed=9, E=E)


def test_unsupported_configs_raise_cleanly():
    from wesep_b200.modules.tasnet.convs import Conv1DBlock
    x = torch.zeros(1, 64, 100, device=DEV)
    for kw in (dict(causal=True), dict(skip_con=True), dict(norm="cLN"), dict(norm="BN")):
        args = dict(in_channels=64, out_channels=128, kernel_size=3, dilation=1, norm="gLN", causal=False, skip_con=False)
        args.update(kw)
        with pytest.raises(NotImplementedError):
            Conv1DBlock(**args).to(DEV)(x)
    with pytest.raises(RuntimeError):
        from wesep_b200.modules.common.norm import select_norm
        select_norm("xLN", 4)


# --------------------------------------------------------------------------- encoder / decoder
@pytest.mark.parametrize("n,T", [(2, 3200), (1, 1999), (3, 64000)])
def test_multi_encoder(n, T):
    from wesep_b200.modules.tasnet import MultiEncoder
    enc = MultiEncoder(1, 256, 64, 20, 10)
    synth.fill_state_dict_(enc.state_dict(), seed=2)
    enc = enc.to(DEV)
    x = rnd(n, T, seed=3, scale=0.1)
    e, w1, w2, w3 = enc(x)
    outs = [e, w1, w2, w3]
    gys = [rnd(*o.shape, seed=4 + i) for i, o in enumerate(outs)]
    params = list(enc.parameters())
    grads = torch.autograd.grad(outs, params, gys)
    sd64 = {k: v.detach().double().requires_grad_(True) for k, v in enc.state_dict().items()}
    masks = [w1.detach() > 0, w2.detach() > 0, w3.detach() > 0]     # pin ReLU branches to the CUDA path's
    r = ospex.multi_encoder(sd64, "", x.double(), relu=lambda v: torch.where(masks.pop(0), v, torch.zeros_like(v)))
    names = [k for k, _ in enc.named_parameters()]
    ref = torch.autograd.grad(r, [sd64[k] for k in names], [g.double() for g in gys])
    for nm, a, b in zip(("e", "w1", "w2", "w3"), outs, r):
        check(nm, a, b, 1e-5)
    for nm, a, b in zip(names, grads, ref):
        check(nm, a, b, 1e-4)


@pytest.mark.parametrize("n,K", [(2, 319), (1, 199), (2, 6399)])
def test_multi_decoder(n, K):
    from wesep_b200.modules.tasnet import MultiDecoder
    ops = _ops()
    dec = MultiDecoder(64, 256, 1, 20, 10)
    synth.fill_state_dict_(dec.state_dict(), seed=2)
    dec = dec.to(DEV)
    e = ops.new_act(n, 64, K, DEV)
    e.copy_(rnd(n, 64, K, seed=3))
    e.requires_grad_(True)
    w = ops.new_act(n, 768, K, DEV)
    w.copy_(torch.relu(rnd(n, 768, K, seed=4)))
    w.requires_grad_(True)
    ops.DEBUG_STASH = {}
    ests = dec.forward_cat(e, w)
    stash, ops.DEBUG_STASH = ops.DEBUG_STASH, None
    m_on = stash["decoder_m"] > 0           # the CUDA path's ReLU branches: pinned in the oracle (PReLU / ReLU kink)
    gys = [rnd(*o.shape, seed=5 + i) for i, o in enumerate(ests)]
    params = list(dec.parameters())
    grads = torch.autograd.grad(ests, [e, w] + params, gys)
    sd64 = {k: v.detach().double().requires_grad_(True) for k, v in dec.state_dict().items()}
    e64 = e.detach().double().requires_grad_(True)
    w64 = w.detach().double().requires_grad_(True)
    r = ospex.multi_decoder(sd64, "", e64, w64[:, :256], w64[:, 256:512], w64[:, 512:],
                            relu_on=[m_on[:, :256], m_on[:, 256:512], m_on[:, 512:]])
    names = [k for k, _ in dec.named_parameters()]
    ref = torch.autograd.grad(r, [e64, w64] + [sd64[k] for k in names], [g.double() for g in gys])
    for i, (a, b) in enumerate(zip(ests, r)):
        assert a.shape == b.shape
        check(f"est{i + 1}", a, b, 3e-6)
    for nm, a, b in zip(["de", "dw"] + names, grads, ref):
        check(nm, a, b, 1e-4)


# --------------------------------------------------------------------------- speaker ResBlock / linear / CE
@pytest.mark.parametrize("ci,co,n,T,train", [(64, 64, 2, 300, True), (64, 128, 3, 301, True), (256, 512, 2, 2133, True),
                                             (64, 64, 2, 299, False)])
def test_resblock(ci, co, n, T, train):
    from wesep_b200.modules.tasnet.speaker import ResBlock
    ops = _ops()
    blk = ResBlock(ci, co)
    synth.fill_state_dict_(blk.state_dict(), seed=ci + co + T)
    with torch.no_grad():   # non-trivial running statistics so that eval mode is a real test
        for k, v in blk.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(0.1 * torch.randn(v.shape, generator=torch.Generator().manual_seed(1)))
            if k.endswith("running_var"):
                v.copy_(1.0 + 0.2 * torch.rand(v.shape, generator=torch.Generator().manual_seed(2)))
    sd0 = {k: v.detach().clone() for k, v in blk.state_dict().items()}
    blk = blk.to(DEV).train(train)
    x = ops.new_act(n, ci, T, DEV)
    x.copy_(rnd(n, ci, T, seed=5))
    x.requires_grad_(True)
    y = blk(x)
    gy = rnd(*y.shape, seed=6)
    params = list(blk.parameters())
    grads = torch.autograd.grad(y, [x] + params, gy)
    sd64 = {k: (v.double().to(DEV).requires_grad_(v.is_floating_point() and "running" not in k)) for k, v in sd0.items()}
    x64 = x.detach().double().requires_grad_(True)
    bufs = {}
    y64 = ospex.resblock(sd64, "", x64, train, bufs)
    names = [k for k, _ in blk.named_parameters()]
    ref = torch.autograd.grad(y64, [x64] + [sd64[k] for k in names], gy.double())
    assert y.shape == y64.shape
    check("y", y, y64, 1e-5)
    for nm, a, b in zip(["dx"] + names, grads, ref):
        # PReLU / max-pool branch flips at |v| ~ 1e-7 bound this comparison, not the arithmetic (scalar slopes most)
        check(nm, a, b, 3e-3 if "prelu" in nm else 5e-4)
    if train:
        for k, v in bufs.items():
            check(k, blk.state_dict()[k], v, 1e-5)
        assert int(blk.batch_norm1.num_batches_tracked) == 1


def test_linear_and_cross_entropy():
    ops = _ops()
    x = rnd(7, 256, seed=1).requires_grad_(True)
    W = rnd(251, 256, seed=2, scale=1 / 16).requires_grad_(True)
    b = rnd(251, seed=3).requires_grad_(True)
    lab = torch.randint(0, 251, (7,), generator=torch.Generator().manual_seed(4)).to(DEV)
    z = ops.LinearFn.apply(x, W, b)
    loss = ops.cross_entropy(z, lab)
    gx, gW, gb = torch.autograd.grad(0.5 * loss, (x, W, b))
    x64, W64, b64 = (t.detach().double().requires_grad_(True) for t in (x, W, b))
    z64 = torch.nn.functional.linear(x64, W64, b64)
    l64 = torch.nn.functional.cross_entropy(z64, lab)
    rx, rW, rb = torch.autograd.grad(0.5 * l64, (x64, W64, b64))
    check("logits", z, z64, 1e-6)
    assert abs(float(loss) - float(l64)) <= 1e-5
    check("dx", gx, rx, 1e-5)
    check("dW", gW, rW, 1e-5)
    check("db", gb, rb, 1e-5)


@pytest.mark.parametrize("mode", ["film", "multiply", "additive", "identity"])
@pytest.mark.parametrize("n,C,T", [(2, 64, 319), (3, 256, 6399)])
def test_fuse_prelu_gln(mode, n, C, T):
    ops = _ops()
    x = ops.new_act(n, C, T, DEV)
    x.copy_(rnd(n, C, T, seed=1))
    x.requires_grad_(True)
    ra = (1 + 0.3 * rnd(n, C, seed=2)).requires_grad_(True) if mode in ("film", "multiply") else None
    rb = (0.3 * rnd(n, C, seed=3)).requires_grad_(True) if mode in ("film", "additive") else None
    al = torch.tensor([0.2], device=DEV, requires_grad=True)
    gm = (1 + 0.1 * rnd(C, 1, seed=4)).requires_grad_(True)
    bt = (0.1 * rnd(C, 1, seed=5)).requires_grad_(True)
    z = ops.FusePreluGlnFn.apply(x, ra, rb, al, gm, bt, 0.0)
    gz = rnd(n, C, T, seed=6)
    ins = [t for t in (x, ra, rb, al, gm, bt) if t is not None]
    grads = torch.autograd.grad(z, ins, gz)
    ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    it = iter(ins64)
    x64 = next(it)
    ra64 = next(it) if ra is not None else None
    rb64 = next(it) if rb is not None else None
    al64, gm64, bt64 = next(it), next(it), next(it)
    v = x64
    if ra64 is not None:
        v = v * ra64[:, :, None]
    if rb64 is not None:
        v = v + rb64[:, :, None]
    pre = v.detach()
    y = torch.where(pre > 0, v, al64 * v)
    z64 = ospex.gln(y, gm64, bt64)
    ref = torch.autograd.grad(z64, ins64, gz.double())
    check("z", z, z64, 2e-6)
    for i, (a, b) in enumerate(zip(grads, ref)):
        check(f"grad{i}", a, b, 2e-4)
