"""GPU: the tcgen05/TMA/TMEM GEMM backend (default) vs the legacy mma.sync backend and the fp64 oracle."""
import math

import pytest
import torch

from tests.test_gpu_kernels import _block_case, check, rnd

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(params=[0, 1], ids=["mma_sync", "tcgen05"])
def backend(request):
    from wesep_b200 import _lib
    _lib.set_gemm_backend(request.param)
    yield request.param
    _lib.set_gemm_backend(1)


@pytest.mark.parametrize("n,Kd,M,T,w_trans", [(1, 128, 128, 256, False), (2, 256, 512, 6399, False), (2, 512, 256, 6399, True),
                                              (3, 64, 128, 100, False), (2, 256, 384, 517, True)])
def test_conv1x1_backends(backend, n, Kd, M, T, w_trans):
    from wesep_b200 import ops
    x = ops.new_act(n, Kd, T, DEV)
    x.copy_(rnd(n, Kd, T, seed=1))
    W = rnd(Kd, M, seed=2, scale=1 / math.sqrt(Kd)) if w_trans else rnd(M, Kd, seed=2, scale=1 / math.sqrt(Kd))
    b = rnd(M, seed=3)
    R = ops.new_act(n, M, T, DEV)
    R.copy_(rnd(n, M, T, seed=4))
    y = ops.conv1x1_raw(x, W, w_trans, M, bias=b, epi=2, R=R)
    Wm = W.double().t() if w_trans else W.double()
    ref = torch.einsum("mk,nkt->nmt", Wm, x.double()) + b.double()[None, :, None] + R.double()
    check("y", y, ref, 1e-5)


def test_tc_is_used_and_counts_launches():
    """With the default backend an eligible shape must go through the tcgen05 kernel (split_w + gemm = 2 launches)."""
    from wesep_b200 import _lib, ops
    _lib.set_gemm_backend(1)
    x = ops.new_act(1, 128, 512, DEV)
    x.normal_()
    W = rnd(128, 128, seed=1)
    before = _lib.launch_count()
    ops.conv1x1_raw(x, W, False, 128)
    assert _lib.launch_count() - before == 2
    _lib.set_gemm_backend(0)
    before = _lib.launch_count()
    ops.conv1x1_raw(x, W, False, 128)
    assert _lib.launch_count() - before == 1
    _lib.set_gemm_backend(1)


def test_gemm_backend_and_mode_per_call():
    """`mode_sel` / `backend_sel` of the GEMM argument structs choose the kernel for ONE launch without touching the
    process-wide default: launch counts tell the backends apart, the error level tells the precision modes apart."""
    from wesep_b200 import _lib, ops
    x = ops.new_act(2, 128, 600, DEV)
    x.copy_(rnd(2, 128, 600, seed=1))
    W = rnd(128, 128, seed=2, scale=1 / math.sqrt(128))
    ref = torch.einsum("mk,nkt->nmt", W.double(), x.double())
    before = _lib.launch_count()
    y1 = ops.conv1x1_raw(x, W, False, 128, backend=1)
    assert _lib.launch_count() - before == 2                    # weight split + tcgen05 GEMM
    before = _lib.launch_count()
    y0 = ops.conv1x1_raw(x, W, False, 128, backend=0)
    assert _lib.launch_count() - before == 1                    # mma.sync GEMM
    before = _lib.launch_count()
    ops.conv1x1_raw(x, W, False, 128)
    assert _lib.launch_count() - before == 2                    # the process default (tcgen05) is untouched
    check("tcgen05", y1, ref, 1e-5)
    check("mma.sync", y0, ref, 1e-5)
    y_fast = ops.conv1x1_raw(x, W, False, 128, mode=1)          # single-pass TF32 for this call only
    e = float((y_fast.double() - ref).norm() / ref.norm())
    assert 1e-5 < e < 2e-3, e
    check("default again", ops.conv1x1_raw(x, W, False, 128), ref, 1e-5)
    dW = torch.zeros(128, 128, device=DEV)
    ops.conv1x1_dw_raw(y1, x, dW, backend=0, mode=0)
    check("dw", dW, torch.einsum("nmt,nkt->mk", y1.double(), x.double()), 1e-5)


def test_tcn_block_full_size_backends(backend):
    _block_case(False, n=2, B=256, H=512, T=6399, dil=16, seed=21)


def test_tcn_fuse_block_full_size_backends(backend):
    _block_case(True, n=3, B=256, H=512, T=4799, dil=1, seed=22, E=256)


def test_tcn_block_small_mixed_backends(backend):
    # H=128 is eligible for tcgen05 (M=128), B=64 is not: the block mixes both kernels
    _block_case(False, n=2, B=64, H=128, T=700, dil=8, seed=23)


@pytest.mark.parametrize("n,M,N,T,pro,per_row", [(2, 256, 512, 6399, 1, True), (3, 512, 256, 1000, 0, False),
                                                   (1, 128, 256, 50, 0, False), (32, 512, 256, 6399, 0, False)])
def test_conv1x1_dw_backends(backend, n, M, N, T, pro, per_row):
    from wesep_b200 import ops
    A = ops.new_act(n, M, T, DEV)
    A.copy_(rnd(n, M, T, seed=1))
    B = ops.new_act(n, N, T, DEV)
    B.copy_(rnd(n, N, T, seed=2))
    alpha = torch.tensor([0.3], device=DEV)
    C = torch.zeros((n, M, N) if per_row else (M, N), device=DEV)
    ops.conv1x1_dw_raw(A, B, C, per_row=per_row, pro_b=pro, alpha_b=alpha if pro else None)
    Bd = B.double()
    if pro:
        Bd = torch.where(Bd > 0, Bd, 0.3 * Bd)
    ref = torch.einsum("nmt,nkt->nmk", A.double(), Bd)
    if not per_row:
        ref = ref.sum(0)
    check("C", C, ref, 1e-4 if n * T > 100000 else 1e-5)   # 2e5-term fp32 sums (tensor-core accumulation truncates)


@pytest.mark.parametrize("n,M,N,T,pro,stats", [(2, 256, 256, 2133, 3, False), (3, 256, 512, 711, 3, False),
                                                (2, 128, 256, 1000, 2, True), (2, 512, 512, 2133, 3, False)])
def test_conv1x1_dw_scale_shift_prologues(backend, n, M, N, T, pro, stats):
    """pro_b 2: sc*prelu(b)+sh (gLN apply, per-row statistics); pro_b 3: prelu(sc*b+sh) (BatchNorm apply, then PReLU)
    — the ResBlock weight gradients (wesep/modules/tasnet/speaker.py:31-45)."""
    from wesep_b200 import ops
    A = ops.new_act(n, M, T, DEV)
    A.copy_(rnd(n, M, T, seed=1))
    B = ops.new_act(n, N, T, DEV)
    B.copy_(rnd(n, N, T, seed=2))
    alpha = torch.tensor([0.3], device=DEV)
    gm = (1.0 + 0.1 * rnd(N, seed=3)).to(DEV)
    bt = (0.1 * rnd(N, seed=4)).to(DEV)
    st = None
    Bd = B.double()
    prelu = lambda x: torch.where(x > 0, x, 0.3 * x)
    if stats:
        y = prelu(Bd)
        cnt = float(N * T)
        st = torch.stack([y.sum((1, 2)), (y * y).sum((1, 2))], 1).contiguous()
        mu = (st[:, 0] / cnt).view(n, 1, 1)
        r = 1.0 / torch.sqrt(st[:, 1].view(n, 1, 1) / cnt - mu * mu + 1e-5)
    else:
        mu, r, cnt = 0.0, 1.0, 1.0
    g, b_ = gm.double().view(1, N, 1), bt.double().view(1, N, 1)
    if pro == 2:
        f = g * (prelu(Bd) - mu) * r + b_
    else:
        f = prelu(g * r * (Bd - mu) + b_) if stats else prelu(g * Bd + b_)
    C = torch.zeros((M, N), device=DEV)
    ops.conv1x1_dw_raw(A, B, C, pro_b=pro, alpha_b=alpha, ch_scale_b=gm, ch_shift_b=bt, row_stats_b=st,
                       stat_count=cnt, stat_eps=1e-5 if stats else 0.0)
    ref = torch.einsum("nmt,nkt->mk", A.double(), f)
    check("C", C, ref, 1e-5)


@pytest.mark.parametrize("n,Kd,M,T", [(2, 256, 768, 6399), (3, 256, 256, 517), (2, 128, 512, 2133)])
def test_conv1x1_relu_mask_and_channel_stats_epilogues(backend, n, Kd, M, T):
    """epi 1 (ReLU, encoder.py:99), epi 3 (decoder masks: Y2 = relu(v), Y = aux * relu(v), decoder.py:96-102) and the
    BatchNorm channel statistics by-product of epi 0 (speaker.py:31-45), on channel counts the 2-CTA kernel takes."""
    from wesep_b200 import ops
    x = ops.new_act(n, Kd, T, DEV)
    x.copy_(rnd(n, Kd, T, seed=1))
    W = rnd(M, Kd, seed=2, scale=1 / math.sqrt(Kd))
    b = rnd(M, seed=3)
    v = torch.einsum("mk,nkt->nmt", W.double(), x.double()) + b.double()[None, :, None]
    y1 = ops.conv1x1_raw(x, W, False, M, bias=b, epi=1)
    check("relu", y1, v.clamp_min(0), 1e-5)
    aux = ops.new_act(n, M, T, DEV)
    aux.copy_(rnd(n, M, T, seed=5))
    Y2 = ops.new_act(n, M, T, DEV)
    y3 = ops.conv1x1_raw(x, W, False, M, bias=b, epi=3, R=aux, Y2=Y2)
    check("masks", Y2, v.clamp_min(0), 1e-5)
    check("masked", y3, aux.double() * v.clamp_min(0), 1e-5)
    chs = torch.zeros(M, 2, dtype=torch.float64, device=DEV)
    y0 = ops.conv1x1_raw(x, W, False, M, bias=b, epi=0, ch_stats=chs)
    check("y", y0, v, 1e-5)
    check("ch_sum", chs[:, 0], v.sum((0, 2)), 1e-4)
    check("ch_sumsq", chs[:, 1], (v * v).sum((0, 2)), 1e-5)


def test_tcn_block_direct_param_grads_match_autograd():
    """ops.direct_param_grads(): the block's kernels accumulate into the live .grad buffers (what train_step uses);
    the result must equal the ordinary autograd path (temporary gradients + AccumulateGrad), including accumulation
    on top of a non-zero .grad."""
    from wesep_b200 import ops, synth
    from wesep_b200.modules.tasnet.convs import Conv1DBlock
    torch.manual_seed(0)
    blk = Conv1DBlock(256, 512, 3, 4, "gLN", False, False)
    synth.fill_state_dict_(blk.state_dict(), seed=3)
    blk = blk.to(DEV)
    x = ops.new_act(2, 256, 1500, DEV)
    x.copy_(rnd(2, 256, 1500, seed=1))
    g = ops.new_act(2, 256, 1500, DEV)
    g.copy_(rnd(2, 256, 1500, seed=2))
    params = list(blk.parameters())
    base = [0.01 * torch.randn_like(p) for p in params]

    def run(direct):
        for p, b in zip(params, base):
            p.grad = b.clone()
        xx = x.clone().requires_grad_(True)
        y = blk(xx)
        if direct:
            with ops.direct_param_grads():
                y.backward(g)
        else:
            y.backward(g)
        return [p.grad.clone() for p in params], xx.grad.clone()

    ref_g, ref_dx = run(False)
    got_g, got_dx = run(True)
    check("dx", got_dx, ref_dx, 1e-6)
    for (name, _), a, b in zip(blk.named_parameters(), got_g, ref_g):
        check(name, a, b, 2e-5)
