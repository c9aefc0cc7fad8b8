"""torch.autograd.Function wrappers over the C-ABI kernels (include/wesep_b200.h).

PyTorch here is plumbing only: device memory (caching allocator), streams and the autograd
tape.  All arithmetic on the hot path runs in libwesep_b200.so; there is no fallback.

Activation layout ("act"): fp32 ``[n, C, T]`` views of ``[n, C, ld]`` storage with the time
axis contiguous and ``ld = ceil32(T)`` so every row is a whole number of 128-byte TMA atoms (6399 frames -> 6400).
"""
import os

import torch

from . import _lib
from ._lib import STRUCTS


DEBUG_STASH = None   # tests may set this to a dict to inspect the TCN backward workspaces


# --------------------------------------------------------------------------- helpers
def _stream():
    return torch.cuda.current_stream().cuda_stream


def _args(name, **kw):
    s = STRUCTS[name]()
    for k, v in kw.items():
        if isinstance(v, torch.Tensor):
            v = v.data_ptr()
        elif v is None:
            v = 0
        setattr(s, k, v)
    return s


_WS = {}


def gemm_ws(device, nbytes=8 << 20):
    """Per-device scratch for the tcgen05 GEMMs' split-weight copies (all launches are stream-ordered)."""
    key = (device.type, device.index)
    t = _WS.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = t
    return t


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("wesep_b200 kernels need CUDA tensors (no CPU fallback)")


def ceil4(T):
    """row stride of an act tensor: a multiple of 32 floats (128 B) so rows are whole TMA swizzle atoms."""
    return (T + 31) // 32 * 32


def new_act(n, C, T, device, zero=False):
    ld = ceil4(T)
    buf = (torch.zeros if zero else torch.empty)((n, C, ld), dtype=torch.float32, device=device)
    return buf[:, :, :T] if ld != T else buf


def is_act(x):
    return (x.dim() == 3 and x.dtype == torch.float32 and x.is_cuda and x.stride(2) == 1
            and x.stride(1) == ceil4(x.shape[2]) and x.stride(0) == x.shape[1] * x.stride(1) and x.data_ptr() % 16 == 0)


def is_act_slice(x):
    """act layout, or a channel slice of one (batch stride larger than C*ld)."""
    return (x.dim() == 3 and x.dtype == torch.float32 and x.is_cuda and x.stride(2) == 1
            and x.stride(1) == ceil4(x.shape[2]) and x.stride(0) % 4 == 0 and x.stride(0) >= x.shape[1] * x.stride(1)
            and x.data_ptr() % 16 == 0)


def as_act(x):
    """Return x in act layout (copy only if its strides do not qualify)."""
    _check_cuda(x)
    if x.dim() != 3:
        raise RuntimeError("expected a 3-D [n, C, T] tensor")
    if is_act(x):
        return x
    y = new_act(x.shape[0], x.shape[1], x.shape[2], x.device)
    y.copy_(x)
    return y


def _w2d(w):
    """conv weight (O, I, 1) / linear weight (O, I) -> contiguous 2-D view."""
    w2 = w.reshape(w.shape[0], -1)
    return w2 if w2.is_contiguous() else w2.contiguous()


def _vec(p):
    v = p.reshape(-1)
    return v if v.is_contiguous() else v.contiguous()


def _sig(x):
    """[n, L] signal with contiguous samples -> (tensor, ld)."""
    if x.dim() != 2 or x.dtype != torch.float32:
        raise RuntimeError("expected a float32 [n, L] signal")
    if x.stride(1) != 1:
        x = x.contiguous()
    return x, x.stride(0)


# --------------------------------------------------------------------------- raw calls
def conv1x1_raw(x, W2d, w_trans, M, *, bias=None, row_bias=None, pro=0, alpha=None, ch_scale=None, ch_shift=None,
                row_stats=None, stat_count=1.0, stat_eps=0.0, epi=0, R=None, Y=None, Y2=None, out_stats=None,
                out_alpha=None, ch_stats=None, mode=None, backend=None):
    n, Kd, T = x.shape
    if Y is None:
        Y = new_act(n, M, T, x.device)
    a = _args("WesepGemmArgs", n=n, M=M, Kd=Kd, T=T, W=W2d, ldw=W2d.stride(0), w_trans=int(w_trans), X=x, ldx=x.stride(1),
              Y=Y, ldy=Y.stride(1), bias=bias, row_bias=row_bias, pro=pro, alpha=alpha, ch_scale=ch_scale,
              ch_shift=ch_shift, row_stats=row_stats, stat_count=float(stat_count), stat_eps=float(stat_eps), epi=epi,
              R=R, ldr=0 if R is None else R.stride(1), Y2=Y2, ldy2=0 if Y2 is None else Y2.stride(1),
              out_stats=out_stats, out_alpha=out_alpha, ch_stats=ch_stats, bsx=x.stride(0), bsy=Y.stride(0),
              bsr=0 if R is None else R.stride(0), bsy2=0 if Y2 is None else Y2.stride(0),
              mode_sel=0 if mode is None else int(mode) + 1, backend_sel=0 if backend is None else int(backend) + 1)
    ws = gemm_ws(x.device)
    a.ws, a.ws_bytes = ws.data_ptr(), ws.numel()
    for t_ in (x, Y, R, Y2):
        if t_ is not None and not is_act_slice(t_):
            raise RuntimeError("conv1x1: operand is not in act layout")
    _lib.call("wesep_b200_conv1x1", a, _stream())
    return Y


def conv1x1_dw_raw(A, B, C, *, per_row=False, pro_b=0, alpha_b=None, ch_scale_b=None, ch_shift_b=None,
                   row_stats_b=None, stat_count=1.0, stat_eps=0.0, mode=None, backend=None):
    """C[M][N] += sum_n sum_t A[n][M][t] * f(B[n][N][t])  (C dense, zero-initialised by the caller)."""
    n, M, T = A.shape
    N = B.shape[1]
    a = _args("WesepGemmDwArgs", n=n, M=M, N=N, T=T, A=A, lda=A.stride(1), B=B, ldb=B.stride(1), C=C, per_row=int(per_row),
              pro_b=pro_b, alpha_b=alpha_b, ch_scale_b=ch_scale_b, ch_shift_b=ch_shift_b, row_stats_b=row_stats_b,
              stat_count=float(stat_count), stat_eps=float(stat_eps), bsa=A.stride(0), bsb=B.stride(0),
              ldc=C.stride(0) if C.dim() == 2 else C.stride(1),
              mode_sel=0 if mode is None else int(mode) + 1, backend_sel=0 if backend is None else int(backend) + 1)
    if not (is_act_slice(A) and is_act_slice(B)) or C.stride(-1) != 1:
        raise RuntimeError("conv1x1_dw: operand layout")
    _lib.call("wesep_b200_conv1x1_dw", a, _stream())
    return C


def rowsum_raw(x):
    n, C, T = x.shape
    out = torch.empty((n, C), dtype=torch.float32, device=x.device)
    _lib.call("wesep_b200_rowsum", _args("WesepRowSumArgs", n=n, C=C, T=T, ld=x.stride(1), x=x, out=out), _stream())
    return out


def frames_raw(sig, J, K, hop):
    sig, ldx = _sig(sig)
    n, S = sig.shape
    F = new_act(n, J, K, sig.device)
    _lib.call("wesep_b200_frames", _args("WesepFrameArgs", n=n, J=J, K=K, hop=hop, S=S, x=sig, ldx=ldx, F=F,
                                         ldf=F.stride(1)), _stream())
    return F


# --------------------------------------------------------------------------- fused TCN block
# When set (wesep_b200.utils.executor.train_step does, around loss.backward()), TCNBlockFn.backward lets its kernels
# accumulate parameter gradients straight into the existing `.grad` buffers (the optimizer's flat arena) and returns None
# for them: no temporary gradient, no zero-fill, and no per-parameter `grad += new` kernel from autograd's AccumulateGrad
# (12 parameters x 32 blocks per step).  Off by default so torch.autograd.grad(...) keeps returning the gradients.
DIRECT_PARAM_GRADS = False


class direct_param_grads:
    """Context manager enabling in-place accumulation into existing `.grad` buffers during backward."""

    def __enter__(self):
        global DIRECT_PARAM_GRADS
        self.prev = DIRECT_PARAM_GRADS
        DIRECT_PARAM_GRADS = True
        return self

    def __exit__(self, *exc):
        global DIRECT_PARAM_GRADS
        DIRECT_PARAM_GRADS = self.prev
        return False


def _grad_targets(params):
    """Flat fp32 views of the params' existing .grad buffers, or None if any of them cannot take direct accumulation."""
    outs = []
    for p in params:
        g = getattr(p, "grad", None)
        if (g is None or not p.is_leaf or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device
                or g.shape != p.shape or g.data_ptr() % 16):
            return None
        outs.append(g.view(-1))
    return outs


class TCNBlockFn(torch.autograd.Function):
    """Conv1DBlock / Conv1DBlock4Fuse (wesep/modules/tasnet/convs.py:43-160) as one fused op."""

    @staticmethod
    def forward(ctx, x, aux, W1, b1, a1, g1, be1, wd, bd, a2, g2, be2, W3, b3, dil):
        x = as_act(x)
        n, B, T = x.shape
        W1c, W3c = _w2d(W1), _w2d(W3)
        H = W1c.shape[0]
        E = 0
        auxc = None
        if aux is not None:
            auxc = aux.reshape(n, -1).contiguous()
            E = auxc.shape[1]
        if W1c.shape[1] != B + E or W3c.shape != (B, H):
            raise RuntimeError("TCN block: weight shapes do not match the input")
        dev = x.device
        u = new_act(n, H, T, dev)
        d = new_act(n, H, T, dev)
        out = new_act(n, B, T, dev)
        stats = torch.empty((2, n, 2), dtype=torch.float64, device=dev)
        row_bias = torch.empty((n, H), dtype=torch.float32, device=dev) if E else None
        P = dict(b1=_vec(b1), a1=_vec(a1), g1=_vec(g1), be1=_vec(be1), wd=_vec(wd), bd=_vec(bd), a2=_vec(a2),
                 g2=_vec(g2), be2=_vec(be2), b3=_vec(b3))
        fa = _args("WesepTcnFwdArgs", n=n, B=B, H=H, T=T, dil=int(dil), E=E, ld=x.stride(1), x=x, aux=auxc, W1=W1c,
                   ldw1=W1c.stride(0), W3=W3c, ldw3=W3c.stride(0), u=u, d=d, out=out, stats1=stats[0], stats2=stats[1],
                   row_bias=row_bias, **P)
        ws = gemm_ws(dev)
        fa.ws, fa.ws_bytes = ws.data_ptr(), ws.numel()
        if u.stride(1) != x.stride(1):
            raise RuntimeError("TCN block: inconsistent row strides")
        _lib.call("wesep_b200_tcn_block_fwd", fa, _stream())
        ctx.dil = int(dil)
        ctx.has_aux = aux is not None
        ctx.shapes = (W1.shape, W3.shape, None if aux is None else aux.shape)
        ctx.param_refs = (W1, b1, a1, g1, be1, wd, bd, a2, g2, be2, W3, b3)
        ctx.save_for_backward(x, auxc, W1c, W3c, u, d, stats, *[P[k] for k in ("b1", "a1", "g1", "be1", "wd", "bd", "a2",
                                                                            "g2", "be2", "b3")])
        return out

    @staticmethod
    def backward(ctx, gout):
        x, auxc, W1c, W3c, u, d, stats, b1, a1, g1, be1, wd, bd, a2, g2, be2, b3 = ctx.saved_tensors
        n, B, T = x.shape
        H = W1c.shape[0]
        E = 0 if auxc is None else auxc.shape[1]
        dev = x.device
        gout = as_act(gout)
        if gout.stride(1) != x.stride(1):
            g2_ = new_act(n, B, T, dev)
            g2_.copy_(gout)
            gout = g2_
        dx = new_act(n, B, T, dev)
        dd = new_act(n, H, T, dev)
        du = new_act(n, H, T, dev)
        sizes = [H * (B + E), H, 1, H, H, 3 * H, H, 1, H, H, B * H, B]
        direct = _grad_targets(ctx.param_refs) if DIRECT_PARAM_GRADS else None
        if direct is not None and [t.numel() for t in direct] == sizes:
            parts = direct                       # kernels accumulate (+=) into the live .grad buffers
        else:
            direct = None
            flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
            parts = list(torch.split(flat, sizes))
        dW1, db1, da1, dg1, dbe1, dwd, dbd, da2, dg2, dbe2, dW3, db3 = parts
        daux = torch.empty((n, E), dtype=torch.float32, device=dev) if E else None
        # scratch the kernels zero themselves: one buffer [rowsc | sdu | sg | Gn] so the C side needs ONE memset
        scratch = torch.empty(16 * n + n * H + n * B + n * B * H, dtype=torch.float32, device=dev)
        rowsc = scratch[:16 * n].view(torch.float64).view(n, 8)
        sdu = scratch[16 * n:16 * n + n * H].view(n, H)
        sg = scratch[16 * n + n * H:16 * n + n * H + n * B].view(n, B)
        Gn = scratch[16 * n + n * H + n * B:].view(n, B, H)
        fa = _args("WesepTcnFwdArgs", n=n, B=B, H=H, T=T, dil=ctx.dil, E=E, ld=x.stride(1), x=x, aux=auxc, W1=W1c,
                   ldw1=W1c.stride(0), b1=b1, a1=a1, g1=g1, be1=be1, wd=wd, bd=bd, a2=a2, g2=g2, be2=be2, W3=W3c,
                   ldw3=W3c.stride(0), b3=b3, u=u, d=d, out=None, stats1=stats[0], stats2=stats[1], row_bias=None)
        ws = gemm_ws(dev)
        fa.ws, fa.ws_bytes = ws.data_ptr(), ws.numel()
        ba = _args("WesepTcnBwdArgs", gout=gout, dx=dx, dW1=dW1, db1=db1, da1=da1, dg1=dg1, dbe1=dbe1, dwd=dwd, dbd=dbd,
                   da2=da2, dg2=dg2, dbe2=dbe2, dW3=dW3, db3=db3, daux=daux, dd=dd, du=du, Gn=Gn, sg=sg, sdu=sdu,
                   rowsc=rowsc)
        ba.f = fa
        _lib.call("wesep_b200_tcn_block_bwd", ba, _stream())
        if DEBUG_STASH is not None:
            DEBUG_STASH.update(dd=dd, du=du, Gn=Gn, sg=sg, sdu=sdu, rowsc=rowsc, u=u, d=d, stats=stats)
        W1s, W3s, auxs = ctx.shapes
        g_aux = None if daux is None else daux.reshape(auxs)
        if direct is not None:
            return (dx, g_aux) + (None,) * 13
        return (dx, g_aux, dW1.view(W1s), db1, da1, dg1.view(H, 1), dbe1.view(H, 1), dwd.view(H, 1, 3), dbd, da2,
                dg2.view(H, 1), dbe2.view(H, 1), dW3.view(W3s), db3, None)


def tcn_block(x, aux, W1, b1, a1, g1, be1, wd, bd, a2, g2, be2, W3, b3, dil):
    return TCNBlockFn.apply(x, aux, W1, b1, a1, g1, be1, wd, bd, a2, g2, be2, W3, b3, dil)


# --------------------------------------------------------------------------- generic 1x1 conv
class Conv1x1Fn(torch.autograd.Function):
    """y = act(W x + b) for x [n, I, T]; W given 2-D as (O, I) or, w_trans, as (I, O). act in {None, 'relu'}."""

    @staticmethod
    def forward(ctx, x, W2d, bias, w_trans, act):
        x = x if is_act_slice(x) else as_act(x)            # channel slices of an act tensor are read in place
        W2d = W2d if W2d.is_contiguous() else W2d.contiguous()
        M = W2d.shape[1] if w_trans else W2d.shape[0]
        Kd = W2d.shape[0] if w_trans else W2d.shape[1]
        if Kd != x.shape[1]:
            raise RuntimeError("conv1x1: channel mismatch")
        y = conv1x1_raw(x, W2d, w_trans, M, bias=None if bias is None else _vec(bias), epi=1 if act == "relu" else 0)
        ctx.w_trans, ctx.act, ctx.has_bias = bool(w_trans), act, bias is not None
        ctx.save_for_backward(x, W2d, y if act == "relu" else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W2d, y = ctx.saved_tensors
        n, Kd, T = x.shape
        g = as_act(gy)
        if ctx.act == "relu":
            g2 = new_act(n, g.shape[1], T, g.device)
            if g.stride(1) != y.stride(1):
                raise RuntimeError("conv1x1 backward: stride mismatch")
            _lib.call("wesep_b200_relu_bwd", _args("WesepReluBwdArgs", n=n, C=g.shape[1], T=T, ld=g.stride(1), y=y, gy=g,
                                                   gx=g2), _stream())
            g = g2
        M = g.shape[1]
        dW = torch.zeros_like(W2d)
        if ctx.w_trans:   # W stored [Kd][M]: dW[k][m] = sum x[k,t] g[m,t]
            conv1x1_dw_raw(x, g, dW)
        else:
            conv1x1_dw_raw(g, x, dW)
        db = rowsum_raw(g).sum(0) if ctx.has_bias else None
        dx = None
        if ctx.needs_input_grad[0]:
            dx = conv1x1_raw(g, W2d, not ctx.w_trans, Kd)
        return dx, dW, db, None, None


def conv1x1(x, W2d, bias=None, w_trans=False, act=None):
    return Conv1x1Fn.apply(x, W2d, bias, w_trans, act)


# --------------------------------------------------------------------------- cLN
class ClnFn(torch.autograd.Function):
    """ChannelWiseLayerNorm (wesep/modules/common/norm.py:51-66)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = as_act(x)
        n, C, T = x.shape
        y = new_act(n, C, T, x.device)
        mean = torch.empty((n, T), dtype=torch.float32, device=x.device)
        rstd = torch.empty((n, T), dtype=torch.float32, device=x.device)
        gamma, beta = _vec(gamma), _vec(beta)
        _lib.call("wesep_b200_cln_fwd", _args("WesepClnFwdArgs", n=n, C=C, T=T, ldx=x.stride(1), ldy=y.stride(1), x=x, y=y,
                                              gamma=gamma, beta=beta, eps=float(eps), mean=mean, rstd=rstd), _stream())
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, mean, rstd = ctx.saved_tensors
        n, C, T = x.shape
        g = as_act(gy)
        dx = new_act(n, C, T, x.device)
        dgb = torch.zeros((2, C), dtype=torch.float32, device=x.device)
        _lib.call("wesep_b200_cln_bwd", _args("WesepClnBwdArgs", n=n, C=C, T=T, ldx=x.stride(1), ldg=g.stride(1),
                                              lddx=dx.stride(1), x=x, gy=g, dx=dx, gamma=gamma, mean=mean, rstd=rstd,
                                              dgamma=dgb[0], dbeta=dgb[1]), _stream())
        return dx, dgb[0], dgb[1], None


def cln(x, gamma, beta, eps=1e-5):
    return ClnFn.apply(x, gamma, beta, eps)


# --------------------------------------------------------------------------- encoder / decoder
def _relu_bwd(y, g):
    n, C, T = y.shape
    g = as_act(g)
    out = new_act(n, C, T, y.device)
    if not (is_act(y) and g.stride(1) == y.stride(1) == out.stride(1)):
        raise RuntimeError("relu backward: layout mismatch")
    _lib.call("wesep_b200_relu_bwd", _args("WesepReluBwdArgs", n=n, C=C, T=T, ld=y.stride(1), y=y, gy=g, gx=out), _stream())
    return out


class MultiEncoderConvFn(torch.autograd.Function):
    """cat_i relu(conv1d(x, w_i, b_i, stride=hop)) for the three filter lengths of MultiEncoder
    (wesep/modules/tasnet/encoder.py:95-110): framing + three GEMMs writing one [n, 3N, K] tensor.
    The right zero-padding of the input (encoder.py:103-109) is the zero fill of the framing."""

    @staticmethod
    def forward(ctx, x, hop, *wb):
        ws, bs = wb[0::2], wb[1::2]
        _check_cuda(x, *ws)
        n, T = x.shape
        N = ws[0].shape[0]
        Ls = [w.shape[-1] for w in ws]
        if any(L % 4 for L in Ls) or T < Ls[0]:
            raise RuntimeError("MultiEncoder: filter lengths must be multiples of 4 and <= signal length")
        K = (T - Ls[0]) // hop + 1
        F = frames_raw(x, max(Ls), K, hop)
        w_cat = new_act(n, len(ws) * N, K, x.device)
        W2 = [_w2d(w) for w in ws]
        for i, (W, b, L) in enumerate(zip(W2, bs, Ls)):
            conv1x1_raw(F[:, :L], W, False, N, bias=_vec(b), epi=1, Y=w_cat[:, i * N:(i + 1) * N])
        ctx.N, ctx.Ls = N, Ls
        ctx.wshapes = [w.shape for w in ws]
        ctx.save_for_backward(F, w_cat)
        return w_cat

    @staticmethod
    def backward(ctx, g_cat):
        F, w_cat = ctx.saved_tensors
        N, Ls = ctx.N, ctx.Ls
        g = _relu_bwd(w_cat, g_cat)
        rs = rowsum_raw(g).sum(0)
        outs = []
        for i, L in enumerate(Ls):
            dW = torch.zeros((N, L), dtype=torch.float32, device=g.device)
            conv1x1_dw_raw(g[:, i * N:(i + 1) * N], F[:, :L], dW)
            outs += [dW.view(ctx.wshapes[i]), rs[i * N:(i + 1) * N]]
        return (None, None, *outs)


class DecoderMasksFn(torch.autograd.Function):
    """S_i = w_i * relu(mask_i(e)) for i = 1..3 as ONE GEMM over concatenated mask weights
    (MultiDecoder.forward, wesep/modules/tasnet/decoder.py:96-102)."""

    @staticmethod
    def forward(ctx, e, w_cat, *wb):
        ws, bs = wb[0::2], wb[1::2]
        e, w_cat = as_act(e), as_act(w_cat)
        Wc = torch.cat([_w2d(w) for w in ws], 0).contiguous()
        bc = torch.cat([_vec(b) for b in bs], 0).contiguous()
        n, _, T = e.shape
        M = Wc.shape[0]
        if w_cat.shape[1] != M:
            raise RuntimeError("decoder masks: encoder output / mask width mismatch")
        m = new_act(n, M, T, e.device)
        S = conv1x1_raw(e, Wc, False, M, bias=bc, epi=3, R=w_cat, Y2=m)
        if DEBUG_STASH is not None:
            DEBUG_STASH.update(decoder_m=m)
        ctx.wshapes = [w.shape for w in ws]
        ctx.save_for_backward(e, w_cat, Wc, m)
        return S

    @staticmethod
    def backward(ctx, gS):
        e, w_cat, Wc, m = ctx.saved_tensors
        n, B, T = e.shape
        M = Wc.shape[0]
        gS = as_act(gS)
        gw = new_act(n, M, T, e.device)
        gm = new_act(n, M, T, e.device)
        if not (gS.stride(1) == w_cat.stride(1) == m.stride(1) == gw.stride(1)):
            raise RuntimeError("decoder masks backward: stride mismatch")
        _lib.call("wesep_b200_mask_bwd", _args("WesepMaskBwdArgs", n=n, C=M, T=T, ld=gw.stride(1), gS=gS, w=w_cat, m=m,
                                               gw=gw, gm=gm, acc_w=0), _stream())
        dW = torch.zeros_like(Wc)
        conv1x1_dw_raw(gm, e, dW)
        db = rowsum_raw(gm).sum(0)
        de = conv1x1_raw(gm, Wc, True, B)
        k = len(ctx.wshapes)
        Ni = M // k
        outs = []
        for i in range(k):
            outs += [dW[i * Ni:(i + 1) * Ni].reshape(ctx.wshapes[i]), db[i * Ni:(i + 1) * Ni]]
        return (de, gw, *outs)


class DecoderBasisFn(torch.autograd.Function):
    """est_i = ConvTranspose1d(N, 1, L_i, stride=hop)(S_i)[:, :xlen]  (decoder.py:104-108) as a basis
    GEMM F_i = D_i^T S_i followed by a gather-form overlap-add (no atomics)."""

    @staticmethod
    def forward(ctx, S_cat, hop, xlen, *db):
        Ds, cs = db[0::2], db[1::2]
        S_cat = as_act(S_cat)
        n, M, K = S_cat.shape
        k = len(Ds)
        N = M // k
        D2 = [_w2d(D) for D in Ds]          # (N, 1, L) -> [N, L]
        ests = []
        for i, (D, c) in enumerate(zip(D2, cs)):
            L = D.shape[1]
            if L % 4:
                raise RuntimeError("MultiDecoder: filter lengths must be multiples of 4")
            Fi = conv1x1_raw(S_cat[:, i * N:(i + 1) * N], D, True, L)
            y = torch.empty((n, xlen), dtype=torch.float32, device=S_cat.device)
            _lib.call("wesep_b200_overlap_add", _args("WesepOlaArgs", n=n, J=L, K=K, hop=hop, S=xlen, F=Fi, ldf=Fi.stride(1),
                                                      bias=_vec(c), y=y, ldy=y.stride(0)), _stream())
            ests.append(y)
        ctx.hop, ctx.N = hop, N
        ctx.dshapes = [D.shape for D in Ds]
        ctx.save_for_backward(S_cat, *D2)
        return tuple(ests)

    @staticmethod
    def backward(ctx, *gs):
        S_cat, *D2 = ctx.saved_tensors
        n, M, K = S_cat.shape
        N, hop = ctx.N, ctx.hop
        dS = new_act(n, M, K, S_cat.device)
        outs = []
        for i, (D, g) in enumerate(zip(D2, gs)):
            L = D.shape[1]
            if g is None:
                dS[:, i * N:(i + 1) * N].zero_()
                outs += [None, None]
                continue
            g = g.contiguous()
            dF = frames_raw(g, L, K, hop)
            conv1x1_raw(dF, D, False, N, Y=dS[:, i * N:(i + 1) * N])
            dD = torch.zeros_like(D)
            conv1x1_dw_raw(S_cat[:, i * N:(i + 1) * N], dF, dD)
            outs += [dD.view(ctx.dshapes[i]), g.sum().reshape(1)]
        return (dS, None, None, *outs)


# --------------------------------------------------------------------------- SI-SDR
class SisdrFn(torch.autograd.Function):
    """losses[i] = SISDRLoss()(est_i, tgt) for up to 4 estimates sharing one target
    (auraloss.time.SISDRLoss as used at wesep/utils/losses.py:24-25)."""

    @staticmethod
    def forward(ctx, tgt, *ests):
        k = len(ests)
        if not 1 <= k <= 4:
            raise RuntimeError("sisdr: 1..4 estimates")
        _check_cuda(tgt, *ests)
        tgt, ldt = _sig(tgt)
        n, L = tgt.shape
        es = [_sig(e) for e in ests]
        for e, _ in es:
            if e.shape != (n, L):
                raise RuntimeError("sisdr: estimate/target shape mismatch")
        dev = tgt.device
        sums = torch.empty((k, n, 5), dtype=torch.float64, device=dev)
        rows = torch.empty((k, n), dtype=torch.float32, device=dev)
        loss = torch.empty((k,), dtype=torch.float32, device=dev)
        a = _args("WesepSisdrFwdArgs", n_est=k, n=n, L=L, tgt=tgt, ld_tgt=ldt, sums=sums, sisdr_rows=rows, loss=loss)
        for i, (e, ld) in enumerate(es):
            a.est[i] = e.data_ptr()
            a.ld_est[i] = ld
        _lib.call("wesep_b200_sisdr_fwd", a, _stream())
        ctx.save_for_backward(tgt, sums, *[e for e, _ in es])
        ctx.mark_non_differentiable(rows)
        return loss, rows

    @staticmethod
    def backward(ctx, gloss, _grows):
        tgt, sums, *es = ctx.saved_tensors
        k = len(es)
        n, L = tgt.shape
        gloss = gloss.contiguous().float()
        gs = [torch.empty((n, L), dtype=torch.float32, device=tgt.device) for _ in range(k)]
        a = _args("WesepSisdrBwdArgs", n_est=k, n=n, L=L, tgt=tgt, ld_tgt=tgt.stride(0), sums=sums, gloss=gloss)
        for i in range(k):
            a.est[i] = es[i].data_ptr()
            a.ld_est[i] = es[i].stride(0)
            a.gest[i] = gs[i].data_ptr()
            a.ld_gest[i] = L
        _lib.call("wesep_b200_sisdr_bwd", a, _stream())
        return (None, *gs)


def sisdr_losses(ests, tgt):
    """Returns (losses [k], per-row SI-SDR [k, n] in dB)."""
    return SisdrFn.apply(tgt, *ests)


# --------------------------------------------------------------------------- speaker ResBlock / linear / CE
def _bn_finalize(ch_stats, count, weight, bias, running_mean, running_var, momentum, eps, training):
    C = weight.numel()
    dev = weight.device
    out = torch.empty((4, C), dtype=torch.float32, device=dev)   # scale, shift, mean, rstd
    _lib.call("wesep_b200_bn_finalize", _args("WesepBnFinalizeArgs", C=C, count=float(count), ch_stats=ch_stats,
                                              weight=_vec(weight), bias=_vec(bias), running_mean=running_mean,
                                              running_var=running_var, momentum=float(momentum), eps=float(eps),
                                              training=int(training), scale=out[0], shift=out[1], mean=out[2], rstd=out[3]),
              _stream())
    return out


class ResBlockFn(torch.autograd.Function):
    """ResBlock.forward (wesep/modules/tasnet/speaker.py:31-45): conv1 - BN - PReLU - conv2 - BN - (+ residual) -
    PReLU - MaxPool1d(3).  The two pointwise convs are tensor-core GEMMs whose epilogues accumulate the BatchNorm
    batch statistics; BN1-apply + PReLU is conv2's operand prologue; BN2-apply + residual + PReLU + max-pool is one
    kernel.  Backward: two passes per BatchNorm (sums, then apply)."""

    @staticmethod
    def forward(ctx, x, W1, W2, Wd, g1, b1, rm1, rv1, g2, b2, rm2, rv2, a1, a2, training, momentum, eps):
        x = as_act(x)
        n, Ci, T = x.shape
        W1c, W2c = _w2d(W1), _w2d(W2)
        Wdc = None if Wd is None else _w2d(Wd)
        Co = W1c.shape[0]
        dev = x.device
        count = n * T
        a1v, a2v = _vec(a1), _vec(a2)
        st = torch.zeros((2, Co, 2), dtype=torch.float64, device=dev)
        c1 = conv1x1_raw(x, W1c, False, Co, ch_stats=st[0] if training else None)
        bn1 = _bn_finalize(st[0], count, g1, b1, rm1, rv1, momentum, eps, training)
        c2 = conv1x1_raw(c1, W2c, False, Co, pro=3, alpha=a1v, ch_scale=bn1[0], ch_shift=bn1[1],
                         ch_stats=st[1] if training else None)
        bn2 = _bn_finalize(st[1], count, g2, b2, rm2, rv2, momentum, eps, training)
        res = x if Wdc is None else conv1x1_raw(x, Wdc, False, Co)
        Tp = T // 3
        y = new_act(n, Co, Tp, dev)
        _lib.call("wesep_b200_bn_act_pool_fwd", _args("WesepBnActPoolFwdArgs", n=n, C=Co, T=T, pool=3, ldx=c2.stride(1),
                                                      ldr=res.stride(1), ldy=y.stride(1), x=c2, res=res, scale=bn2[0],
                                                      shift=bn2[1], alpha=a2v, y=y), _stream())
        ctx.training = bool(training)
        ctx.shapes = (W1.shape, W2.shape, None if Wd is None else Wd.shape)
        ctx.save_for_backward(x, c1, c2, res if Wdc is not None else None, W1c, W2c, Wdc, bn1, bn2, a1v, a2v)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, c1, c2, res, W1c, W2c, Wdc, bn1, bn2, a1v, a2v = ctx.saved_tensors
        n, Ci, T = x.shape
        Co = W1c.shape[0]
        dev = x.device
        count = n * T
        if res is None:
            res = x
        gy = as_act(gy)
        sums = torch.zeros((2, Co, 2), dtype=torch.float64, device=dev)
        dal = torch.zeros(2, dtype=torch.float32, device=dev)
        # ---- second half: maxpool + PReLU_2 + residual + BN2
        gv2 = new_act(n, Co, T, dev)
        _lib.call("wesep_b200_bn_act_pool_bwd", _args("WesepBnActPoolBwdArgs", n=n, C=Co, T=T, pool=3, ldx=c2.stride(1),
                                                      ldr=res.stride(1), ldgy=gy.stride(1), ldgv=gv2.stride(1), x=c2,
                                                      res=res, scale=bn2[0], shift=bn2[1], alpha=a2v, mean=bn2[2],
                                                      rstd=bn2[3], gy=gy, gv=gv2, ch_sums=sums[1], dalpha=dal[1:]),
                  _stream())
        dc2 = new_act(n, Co, T, dev)
        _lib.call("wesep_b200_bn_bwd", _args("WesepBnBwdArgs", n=n, C=Co, T=T, training=int(ctx.training),
                                             ldgv=gv2.stride(1), ldx=c2.stride(1), lddx=dc2.stride(1), count=float(count),
                                             gv=gv2, x=c2, dx=dc2, scale=bn2[0], mean=bn2[2], rstd=bn2[3],
                                             ch_sums=sums[1]), _stream())
        dW2 = torch.zeros_like(W2c)
        conv1x1_dw_raw(dc2, c1, dW2, pro_b=3, alpha_b=a1v, ch_scale_b=bn1[0], ch_shift_b=bn1[1])
        gp1 = conv1x1_raw(dc2, W2c, True, Co)
        # ---- first half: PReLU_1 + BN1
        gv1 = new_act(n, Co, T, dev)
        _lib.call("wesep_b200_bn_act_pool_bwd", _args("WesepBnActPoolBwdArgs", n=n, C=Co, T=T, pool=1, ldx=c1.stride(1),
                                                      ldr=0, ldgy=gp1.stride(1), ldgv=gv1.stride(1), x=c1, res=None,
                                                      scale=bn1[0], shift=bn1[1], alpha=a1v, mean=bn1[2], rstd=bn1[3],
                                                      gy=gp1, gv=gv1, ch_sums=sums[0], dalpha=dal[0:]), _stream())
        dc1 = new_act(n, Co, T, dev)
        _lib.call("wesep_b200_bn_bwd", _args("WesepBnBwdArgs", n=n, C=Co, T=T, training=int(ctx.training),
                                             ldgv=gv1.stride(1), ldx=c1.stride(1), lddx=dc1.stride(1), count=float(count),
                                             gv=gv1, x=c1, dx=dc1, scale=bn1[0], mean=bn1[2], rstd=bn1[3],
                                             ch_sums=sums[0]), _stream())
        dW1 = torch.zeros_like(W1c)
        conv1x1_dw_raw(dc1, x, dW1)
        dWd = None
        if Wdc is None:
            dx = conv1x1_raw(dc1, W1c, True, Ci, epi=2, R=gv2)          # + identity residual
        else:
            dx = conv1x1_raw(dc1, W1c, True, Ci)
            conv1x1_raw(gv2, Wdc, True, Ci, epi=2, R=dx, Y=dx)          # + downsample-conv residual (in place)
            dWd = torch.zeros_like(Wdc)
            conv1x1_dw_raw(gv2, x, dWd)
        sf = sums.float()
        W1s, W2s, Wds = ctx.shapes
        return (dx, dW1.view(W1s), dW2.view(W2s), None if dWd is None else dWd.view(Wds),
                sf[0, :, 1].contiguous(), sf[0, :, 0].contiguous(), None, None,
                sf[1, :, 1].contiguous(), sf[1, :, 0].contiguous(), None, None, dal[0:1], dal[1:2], None, None, None)


class MeanTimeFn(torch.autograd.Function):
    """x.mean(dim=-1) over the valid frames of an act tensor (ResNet4SpExplus.forward, speaker.py:63)."""

    @staticmethod
    def forward(ctx, x):
        x = as_act(x)
        ctx.T = x.shape[2]
        return rowsum_raw(x) / x.shape[2]

    @staticmethod
    def backward(ctx, g):
        return (g / ctx.T).unsqueeze(-1).expand(-1, -1, ctx.T)


class LinearFn(torch.autograd.Function):
    """nn.Linear on [n, K] rows (pred_linear, wesep/models/convtasnet.py:194)."""

    @staticmethod
    def forward(ctx, x, W, b):
        _check_cuda(x, W)
        x = x.contiguous().float()
        W = W.contiguous()
        n, K = x.shape
        J = W.shape[0]
        y = torch.empty((n, J), dtype=torch.float32, device=x.device)
        _lib.call("wesep_b200_linear_fwd", _args("WesepLinearArgs", n=n, J=J, K=K, x=x, W=W, b=b, y=y), _stream())
        ctx.has_b = b is not None
        ctx.save_for_backward(x, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W = ctx.saved_tensors
        n, K = x.shape
        J = W.shape[0]
        gy = gy.contiguous().float()
        dW = torch.empty_like(W)
        db = torch.empty(J, dtype=torch.float32, device=x.device) if ctx.has_b else None
        dx = torch.empty_like(x)
        _lib.call("wesep_b200_linear_bwd", _args("WesepLinearArgs", n=n, J=J, K=K, x=x, W=W, gy=gy, dW=dW, db=db, dx=dx),
                  _stream())
        return dx, dW, db


class CrossEntropyFn(torch.autograd.Function):
    """nn.CrossEntropyLoss() (mean over rows) on [n, J] logits with int64 labels (wesep/utils/losses.py:11)."""

    @staticmethod
    def forward(ctx, logits, labels):
        _check_cuda(logits, labels)
        logits = logits.contiguous().float()
        labels = labels.contiguous().long()
        n, J = logits.shape
        loss = torch.empty(1, dtype=torch.float32, device=logits.device)
        dlogits = torch.empty_like(logits)
        _lib.call("wesep_b200_cross_entropy", _args("WesepCeArgs", n=n, J=J, logits=logits, labels=labels, loss=loss,
                                                    dlogits=dlogits), _stream())
        ctx.save_for_backward(dlogits)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dlogits,) = ctx.saved_tensors
        return g * dlogits, None


def cross_entropy(logits, labels):
    return CrossEntropyFn.apply(logits, labels)


# --------------------------------------------------------------------------- alternative speaker fusion
class Conv1x1RowBiasFn(torch.autograd.Function):
    """y[n] = W x[n] + row_bias[n] (+ bias): the concat-type speaker fusion with the embedding half of the Linear
    folded into a per-row bias (SpeakerFuseLayer 'concat', wesep/modules/common/speaker.py:88-94)."""

    @staticmethod
    def forward(ctx, x, W2d, row_bias):
        x = as_act(x)
        W2d = W2d.contiguous()
        rb = row_bias.contiguous().float()
        y = conv1x1_raw(x, W2d, False, W2d.shape[0], row_bias=rb)
        ctx.save_for_backward(x, W2d)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, W2d = ctx.saved_tensors
        g = as_act(gy)
        dW = torch.zeros_like(W2d)
        conv1x1_dw_raw(g, x, dW)
        drb = rowsum_raw(g)
        dx = conv1x1_raw(g, W2d, True, x.shape[1])
        return dx, dW, drb


class FusePreluGlnFn(torch.autograd.Function):
    """z = gLN(PReLU(ra[n,c] * x + rb[n,c])): speaker fusion (multiply / additive / FiLM; identity affine for
    concat) + the nn.PReLU and gLN that follow it in FuseSeparation (wesep/modules/tasnet/separation.py:116-126)."""

    @staticmethod
    def forward(ctx, x, ra, rb, alpha, gamma, beta, eps):
        x = as_act(x)
        n, C, T = x.shape
        dev = x.device
        ra_c = None if ra is None else ra.reshape(n, C).contiguous().float()
        rb_c = None if rb is None else rb.reshape(n, C).contiguous().float()
        al, gm, bt = _vec(alpha), _vec(gamma), _vec(beta)
        stats = torch.empty((n, 2), dtype=torch.float64, device=dev)
        y = new_act(n, C, T, dev)
        _lib.call("wesep_b200_fuse_prelu_gln_fwd", _args("WesepFuseArgs", n=n, C=C, T=T, ldx=x.stride(1), ldy=y.stride(1),
                                                         x=x, ra=ra_c, rb=rb_c, alpha=al, gamma=gm, beta=bt, stats=stats,
                                                         y=y, eps=float(eps)), _stream())
        ctx.eps = float(eps)
        ctx.gshape = (gamma.shape, beta.shape)
        ctx.flags = (ra is not None, rb is not None, None if ra is None else ra.shape, None if rb is None else rb.shape)
        ctx.save_for_backward(x, ra_c, rb_c, al, gm, bt, stats)
        return y

    @staticmethod
    def backward(ctx, gz):
        x, ra_c, rb_c, al, gm, bt, stats = ctx.saved_tensors
        n, C, T = x.shape
        dev = x.device
        has_a, has_b, sha, shb = ctx.flags
        gz = as_act(gz)
        dx = new_act(n, C, T, dev)
        rowsums = torch.empty((n, 2), dtype=torch.float64, device=dev)
        dra = torch.empty((n, C), dtype=torch.float32, device=dev) if has_a else None
        drb = torch.empty((n, C), dtype=torch.float32, device=dev) if has_b else None
        acc = torch.zeros(1 + 2 * C, dtype=torch.float32, device=dev)
        _lib.call("wesep_b200_fuse_prelu_gln_bwd", _args("WesepFuseArgs", n=n, C=C, T=T, ldx=x.stride(1), ldg=gz.stride(1),
                                                         lddx=dx.stride(1), x=x, ra=ra_c, rb=rb_c, alpha=al, gamma=gm,
                                                         beta=bt, stats=stats, gz=gz, rowsums=rowsums, dx=dx, dra=dra,
                                                         drb=drb, dalpha=acc[0:], dgamma=acc[1:], dbeta=acc[1 + C:],
                                                         eps=ctx.eps),
                  _stream())
        return (dx, None if dra is None else dra.view(sha), None if drb is None else drb.view(shb), acc[0:1],
                acc[1:1 + C].view(ctx.gshape[0]), acc[1 + C:].view(ctx.gshape[1]), None)


# --------------------------------------------------------------------------- pBSRNN building blocks
def swap_oi_raw(x, nb=1, res=None):
    """x: act [nb*Q, C, S] -> act [nb*S, C, Q] with out[b, s, c, q] = x[b, q, c, s] (+ res, laid out like the result)."""
    x = as_act(x)
    nQ, C, S = x.shape
    if nQ % nb:
        raise RuntimeError("swap_oi: leading dimension is not a multiple of the batch count")
    Q = nQ // nb
    out = new_act(nb * S, C, Q, x.device)
    if res is not None:
        res = as_act(res)
        if res.shape != out.shape or res.stride(1) != out.stride(1):
            raise RuntimeError("swap_oi: residual layout")
    _lib.call("wesep_b200_swap_outer_inner", _args("WesepTransposeArgs", nb=nb, Q=Q, C=C, S=S, ld_in=x.stride(1),
                                                   ld_out=out.stride(1), **{"in": x}, out=out, res=res), _stream())
    return out


class SwapOIFn(torch.autograd.Function):
    """[nb][Q][C][S] -> [nb][S][C][Q] (optionally + residual): ResRNN's reference <-> time-major layout change and
    BSNet's permute(0, 3, 2, 1) (wesep/models/bsrnn.py:38-46,70-83).  The adjoint is the same kernel."""

    @staticmethod
    def forward(ctx, x, nb, res):
        ctx.nb = int(nb)
        ctx.has_res = res is not None
        return swap_oi_raw(x, ctx.nb, res)

    @staticmethod
    def backward(ctx, gy):
        gy = as_act(gy)
        return swap_oi_raw(gy, ctx.nb), None, (gy if ctx.has_res else None)


def _cell_args(G, c_prev, c, h=None, dh=None, dc_in=None, dc_prev=None):
    _, Hd4, Q = G.shape
    return _args("WesepLstmCellArgs", Hd=Hd4 // 4, Q=Q, ld=G.stride(1), G=G, c_prev=c_prev, c=c, h=h, dh=dh, dc_in=dc_in,
                 dc_prev=dc_prev)


class LstmTmFn(torch.autograd.Function):
    """Bidirectional single-layer nn.LSTM (wesep/models/bsrnn.py:25-31) on a TIME-MAJOR act tensor xn [S, C, Q]
    (rows = time steps, columns = sequences) -> h [S, 2*Hd, Q] (forward direction in channels [0, Hd), backward in
    [Hd, 2*Hd)).  Input projection of both directions = one GEMM.  The recurrence of both directions and all S steps is
    ONE persistent cluster kernel (`wesep_b200_lstm_rec_fwd`: W_hh resident in distributed shared memory, tcgen05 step
    product, h exchanged over DSMEM) when Hd = 32 * (1..8); other sizes take the step-by-step path (one conv1x1 GEMM +
    one cell kernel per step).  The gate activations overwrite the pre-activations and are what the backward (BPTT,
    `wesep_b200_lstm_rec_bwd`) consumes — it overwrites them in turn with d(pre-activations), so a second backward over
    the same graph raises."""

    @staticmethod
    def forward(ctx, xn, w_ih_f, w_hh_f, b_ih_f, b_hh_f, w_ih_b, w_hh_b, b_ih_b, b_hh_b):
        xn = as_act(xn)
        S, C, Q = xn.shape
        Hd = w_hh_f.shape[1]
        dev = xn.device
        Wih = torch.cat([w_ih_f, w_ih_b], 0).contiguous()
        bias = torch.cat([b_ih_f + b_hh_f, b_ih_b + b_hh_b]).contiguous()
        whh = (w_hh_f.contiguous(), w_hh_b.contiguous())
        G = conv1x1_raw(xn, Wih, False, 8 * Hd, bias=bias)
        H = new_act(S, 2 * Hd, Q, dev)
        Cs = new_act(S, 2 * Hd, Q, dev)
        st = _stream()
        rec = lstm_rec_supported(Hd)
        if rec:
            _lib.call("wesep_b200_lstm_rec_fwd",
                      _args("WesepLstmRecArgs", S=S, Q=Q, Hd=Hd, ld=G.stride(1), bsG=G.stride(0), bsH=H.stride(0), G=G, H=H,
                            C=Cs, Whh_f=whh[0], Whh_r=whh[1], seqs_per_cluster=_rec_force()), st)
        for d in range(0 if rec else 2):
            prev = None
            for s in (range(S) if d == 0 else range(S - 1, -1, -1)):
                Gs = G[s:s + 1, 4 * Hd * d:4 * Hd * (d + 1)]
                if prev is not None:
                    conv1x1_raw(H[prev:prev + 1, Hd * d:Hd * (d + 1)], whh[d], False, 4 * Hd, epi=2, R=Gs, Y=Gs)
                _lib.call("wesep_b200_lstm_cell_fwd",
                          _cell_args(Gs, None if prev is None else Cs[prev:prev + 1, Hd * d:Hd * (d + 1)],
                                     Cs[s:s + 1, Hd * d:Hd * (d + 1)], h=H[s:s + 1, Hd * d:Hd * (d + 1)]), st)
                prev = s
        ctx.save_for_backward(xn, Wih, whh[0], whh[1], G, Cs, H)
        ctx.rec = rec
        ctx.consumed = False
        return H

    @staticmethod
    def backward(ctx, gH):
        if ctx.consumed:
            raise RuntimeError("LstmTmFn: the saved gate activations were overwritten by the first backward; "
                               "a second backward over the same graph is not supported")
        ctx.consumed = True
        xn, Wih, whh_f, whh_b, G, Cs, H = ctx.saved_tensors
        S, C, Q = xn.shape
        Hd = whh_f.shape[1]
        dev = xn.device
        st = _stream()
        whh = (whh_f, whh_b)
        if ctx.rec:
            gH = as_act(gH)
            _lib.call("wesep_b200_lstm_rec_bwd",
                      _args("WesepLstmRecArgs", S=S, Q=Q, Hd=Hd, ld=G.stride(1), bsG=G.stride(0), bsH=H.stride(0), G=G, H=H,
                            C=Cs, Whh_f=whh[0], Whh_r=whh[1], dH=gH, seqs_per_cluster=_rec_force()), st)
        else:
            dH = new_act(S, 2 * Hd, Q, dev)
            dH.copy_(gH)                               # accumulated in place below: never touch autograd's tensor
            dc = [new_act(1, Hd, Q, dev), new_act(1, Hd, Q, dev)]
        for d in range(0 if ctx.rec else 2):
            order = list(range(S)) if d == 0 else list(range(S - 1, -1, -1))
            dc_in = None
            for k in range(S - 1, -1, -1):             # reverse of the forward order of this direction
                s = order[k]
                prev = order[k - 1] if k > 0 else None
                Gs = G[s:s + 1, 4 * Hd * d:4 * Hd * (d + 1)]
                dc_out = dc[k & 1]
                _lib.call("wesep_b200_lstm_cell_bwd",
                          _cell_args(Gs, None if prev is None else Cs[prev:prev + 1, Hd * d:Hd * (d + 1)],
                                     Cs[s:s + 1, Hd * d:Hd * (d + 1)], dh=dH[s:s + 1, Hd * d:Hd * (d + 1)], dc_in=dc_in,
                                     dc_prev=dc_out), st)
                if prev is not None:                   # dL/dh_prev += W_hh^T . d(gates_s)
                    dHp = dH[prev:prev + 1, Hd * d:Hd * (d + 1)]
                    conv1x1_raw(Gs, whh[d], True, Hd, epi=2, R=dHp, Y=dHp)
                dc_in = dc_out
        # G now holds d(pre-activations) of both directions
        dWih = torch.zeros((8 * Hd, C), dtype=torch.float32, device=dev)
        conv1x1_dw_raw(G, xn, dWih)
        db = rowsum_raw(G).sum(0)
        dWhh = [torch.zeros((4 * Hd, Hd), dtype=torch.float32, device=dev) for _ in range(2)]
        if S > 1:
            conv1x1_dw_raw(G[1:, :4 * Hd], H[:-1, :Hd], dWhh[0])
            conv1x1_dw_raw(G[:-1, 4 * Hd:], H[1:, Hd:], dWhh[1])
        # W_ih^T . dG: 8 Hd = 2048 input channels exceed what one tcgen05 launch covers (1024), so one launch per direction,
        # the second accumulating onto the first
        dxn = conv1x1_raw(G[:, :4 * Hd], Wih[:4 * Hd], True, C)
        conv1x1_raw(G[:, 4 * Hd:], Wih[4 * Hd:], True, C, epi=2, R=dxn, Y=dxn)
        return (dxn, dWih[:4 * Hd], dWhh[0], db[:4 * Hd], db[:4 * Hd], dWih[4 * Hd:], dWhh[1], db[4 * Hd:], db[4 * Hd:])


def _rec_force():
    """WESEP_LSTM_REC_SEQS = 64 / 128 forces the sequences-per-cluster grouping of the recurrence kernels (tests)."""
    return int(os.environ.get("WESEP_LSTM_REC_SEQS", "0") or 0)


def lstm_rec_supported(Hd):
    """True when the persistent cluster recurrence covers this hidden size (WESEP_LSTM_REC=0 forces the step-by-step path:
    A/B tests only)."""
    if os.environ.get("WESEP_LSTM_REC") == "0":
        return False
    return bool(_lib.lib().wesep_b200_lstm_rec_supported(int(Hd)))


_ONE = {}


def _one(device):
    t = _ONE.get(device)
    if t is None:
        t = _ONE[device] = torch.ones(1, dtype=torch.float32, device=device)
    return t


GN_EPS = float(torch.finfo(torch.float32).eps)


class GroupNorm1Fn(torch.autograd.Function):
    """nn.GroupNorm(1, C, eps) of an act tensor / channel slice [n, C, T] (per row over (C, T)), one CTA per row
    (`wesep_b200_groupnorm1_fwd / _bwd`)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        if not is_act_slice(x):
            x = as_act(x)
        n, C, T = x.shape
        y = new_act(n, C, T, x.device)
        stats = torch.empty((n, 2), dtype=torch.float64, device=x.device)
        w, b = _vec(weight), _vec(bias)
        _lib.call("wesep_b200_groupnorm1_fwd",
                  _args("WesepGroupNorm1Args", n=n, C=C, T=T, ldx=x.stride(1), bsx=x.stride(0), ldy=y.stride(1), bsy=y.stride(0),
                        x=x, gamma=w, beta=b, eps=float(eps), y=y, stats=stats), _stream())
        ctx.eps = float(eps)
        ctx.save_for_backward(x, w, b, stats)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, b, stats = ctx.saved_tensors
        n, C, T = x.shape
        gy = gy if is_act_slice(gy) else as_act(gy)
        dx = new_act(n, C, T, x.device)
        acc = torch.zeros(2 * C, dtype=torch.float32, device=x.device)
        bsum = torch.empty((n, 2), dtype=torch.float64, device=x.device)
        _lib.call("wesep_b200_groupnorm1_bwd",
                  _args("WesepGroupNorm1Args", n=n, C=C, T=T, ldx=x.stride(1), bsx=x.stride(0), ldg=gy.stride(1), bsg=gy.stride(0),
                        lddx=dx.stride(1), bsdx=dx.stride(0), x=x, gamma=w, beta=b, eps=ctx.eps, stats=stats, gy=gy, dx=dx,
                        dgamma=acc[:C], dbeta=acc[C:], bsum=bsum), _stream())
        return dx, acc[:C], acc[C:], None


def group_norm1(x, weight, bias, eps=GN_EPS):
    """nn.GroupNorm(1, C, eps) of an act tensor [n, C, T] (per row over (C, T))."""
    return GroupNorm1Fn.apply(x, weight, bias, eps)


def res_rnn(x, norm_w, norm_b, lstm, proj_w, proj_b, layer_norm_eps=None):
    """ResRNN.forward (bsrnn.py:38-46) on an act tensor x [Q, C, S] (rows = sequences): GroupNorm -> time-major ->
    BLSTM -> Linear -> back to [Q, C, S] + x.  `lstm` = the 8 nn.LSTM parameters in (forward, reverse) order.
    With `layer_norm_eps` the norm is nn.LayerNorm(C, eps) at every (sequence, step) instead — the intra / inter paths of a
    TF-GridNet block with emb_ks == emb_hs == 1 (gridnet_block.py:139-146,166-172)."""
    xh = group_norm1(x, norm_w, norm_b) if layer_norm_eps is None else cln(x, norm_w, norm_b, layer_norm_eps)
    xn = SwapOIFn.apply(xh, 1, None)                                   # [S, C, Q]
    h = LstmTmFn.apply(xn, *lstm)                                       # [S, 2Hd, Q]
    p = Conv1x1Fn.apply(h, proj_w, proj_b, False, None)                 # [S, C, Q]
    return SwapOIFn.apply(p, 1, x)                                      # [Q, C, S] + residual


def res_rnn_unfold(x, norm_w, norm_b, lstm, ct_w, ct_b, ks, hs, eps):
    """The intra / inter path of a TF-GridNet block with emb_ks != emb_hs (gridnet_block.py:147-161,173-187) on an act tensor
    x [rows, C, S]: LayerNorm(C) -> F.unfold (ks taps, stride hs) -> BLSTM over the (S - ks) / hs + 1 windows ->
    nn.ConvTranspose1d(2 Hd, C, ks, stride hs) = transposed pointwise product + overlap-add -> + x."""
    rows, C, S = x.shape
    xh = cln(x, norm_w, norm_b, eps)
    col = Unfold1dFn.apply(xh, ks, hs)                                  # [rows, C*ks, L]
    xn = SwapOIFn.apply(col, 1, None)                                   # [L, C*ks, rows]
    h = LstmTmFn.apply(xn, *lstm)                                       # [L, 2Hd, rows]
    d = Conv1x1Fn.apply(h, ct_w.reshape(ct_w.shape[0], C * ks), None, True, None)     # [L, C*ks, rows]
    y = Fold1dFn.apply(SwapOIFn.apply(d, 1, None), C, S, ks, hs)        # [rows, C, S]
    y = RowAffineFn.apply(y, None, ct_b[None].expand(rows, C))
    return AddFn.apply(y, x)


def bsnet(x, nband, band_rnn, band_comm):
    """BSNet.forward (bsrnn.py:70-83) on an act tensor x [B, nband*N, T].  band_rnn / band_comm = argument tuples of
    res_rnn (norm_w, norm_b, lstm[8], proj_w, proj_b)."""
    x = as_act(x)
    B, NN, T = x.shape
    N = NN // nband
    ld = x.stride(1)
    xb = x.as_strided((B * nband, N, T), (N * ld, ld, 1))            # [B*nband, N, T]: same memory, bands as rows
    y = res_rnn(xb, *band_rnn)
    yc = SwapOIFn.apply(y, B, None)                                    # [B*T, N, nband]  (permute(0, 3, 2, 1))
    z = res_rnn(yc, *band_comm)
    out = SwapOIFn.apply(z, B, None)                                   # [B*nband, N, T]
    ld2 = out.stride(1)
    return out.as_strided((B, NN, T), (NN * ld2, ld2, 1))


class FixedGemmFn(torch.autograd.Function):
    """y = W x with a constant W (the windowed inverse-DFT basis): conv1x1 without a weight gradient."""

    @staticmethod
    def forward(ctx, x, W, w_trans):
        x = as_act(x)
        M = W.shape[1] if w_trans else W.shape[0]
        ctx.w_trans = bool(w_trans)
        ctx.Kd = x.shape[1]
        ctx.save_for_backward(W)
        return conv1x1_raw(x, W, w_trans, M)

    @staticmethod
    def backward(ctx, gy):
        (W,) = ctx.saved_tensors
        return conv1x1_raw(as_act(gy), W, not ctx.w_trans, ctx.Kd), None, None


def overlap_add_raw(F, hop, S):
    n, J, K = F.shape
    y = torch.empty((n, S), dtype=torch.float32, device=F.device)
    _lib.call("wesep_b200_overlap_add", _args("WesepOlaArgs", n=n, J=J, K=K, hop=hop, S=S, F=F, ldf=F.stride(1), bias=None,
                                              y=y, ldy=y.stride(0)), _stream())
    return y


class OverlapAddFn(torch.autograd.Function):
    """y[n, s] = sum_{j + k*hop = s} F[n, j, k]  (iSTFT synthesis); the adjoint is the framing gather."""

    @staticmethod
    def forward(ctx, F, hop, S):
        F = as_act(F)
        ctx.meta = (F.shape[1], F.shape[2], int(hop))
        return overlap_add_raw(F, int(hop), int(S))

    @staticmethod
    def backward(ctx, gy):
        J, K, hop = ctx.meta
        return frames_raw(gy.contiguous(), J, K, hop), None, None


class RowAffineFn(torch.autograd.Function):
    """y = ra[n, c] * x + rb[n, c] (either may be None)."""

    @staticmethod
    def forward(ctx, x, ra, rb):
        x = as_act(x)
        n, C, T = x.shape
        ra_c = None if ra is None else ra.reshape(n, C).contiguous().float()
        rb_c = None if rb is None else rb.reshape(n, C).contiguous().float()
        y = new_act(n, C, T, x.device)
        _lib.call("wesep_b200_rowaffine_fwd", _args("WesepRowAffineArgs", n=n, C=C, T=T, ld=x.stride(1), x=x, ra=ra_c, rb=rb_c,
                                                    y=y), _stream())
        ctx.shapes = (None if ra is None else ra.shape, None if rb is None else rb.shape)
        ctx.save_for_backward(x, ra_c)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, ra_c = ctx.saved_tensors
        n, C, T = x.shape
        sha, shb = ctx.shapes
        gy = as_act(gy)
        if gy.stride(1) != x.stride(1):
            raise RuntimeError("rowaffine backward: stride mismatch")
        dx = new_act(n, C, T, x.device)
        dra = torch.empty((n, C), dtype=torch.float32, device=x.device) if sha is not None else None
        drb = torch.empty((n, C), dtype=torch.float32, device=x.device) if shb is not None else None
        _lib.call("wesep_b200_rowaffine_bwd", _args("WesepRowAffineArgs", n=n, C=C, T=T, ld=x.stride(1), x=x, ra=ra_c, gy=gy,
                                                    dx=dx, dra=dra, drb=drb), _stream())
        return dx, None if dra is None else dra.view(sha), None if drb is None else drb.view(shb)


class TanhFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = as_act(x)
        n, C, T = x.shape
        y = new_act(n, C, T, x.device)
        _lib.call("wesep_b200_tanh_fwd", _args("WesepTanhArgs", rows=n * C, T=T, ld=x.stride(1), x=x, y=y), _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        n, C, T = y.shape
        gy = as_act(gy)
        dx = new_act(n, C, T, y.device)
        _lib.call("wesep_b200_tanh_bwd", _args("WesepTanhArgs", rows=n * C, T=T, ld=y.stride(1), y=y, gy=gy, dx=dx), _stream())
        return dx


class MaskApplyFn(torch.autograd.Function):
    """Mask-head tail of one band (bsrnn.py:365-379): o [n, 4*bw, T] -> estimate band [n, 2*bw, T] (bw re rows, bw im rows)
    = mixture band * (value * sigmoid(gate)) as a complex product.  `s` is a channel slice of the mixture spectrum."""

    @staticmethod
    def forward(ctx, o, s):
        o = as_act(o)
        n, C4, T = o.shape
        bw = C4 // 4
        if not is_act_slice(s) or s.shape != (n, 2 * bw, T):
            raise RuntimeError("mask_apply: mixture band layout")
        e = new_act(n, 2 * bw, T, o.device)
        _lib.call("wesep_b200_mask_apply_fwd", _args("WesepMaskApplyArgs", n=n, bw=bw, T=T, ldo=o.stride(1), lds=s.stride(1),
                                                     lde=e.stride(1), bso=o.stride(0), bss=s.stride(0), bse=e.stride(0), o=o,
                                                     s=s, e=e), _stream())
        ctx.save_for_backward(o, s)
        return e

    @staticmethod
    def backward(ctx, ge):
        o, s = ctx.saved_tensors
        n, C4, T = o.shape
        bw = C4 // 4
        ge = as_act(ge)
        go = new_act(n, C4, T, o.device)
        _lib.call("wesep_b200_mask_apply_bwd", _args("WesepMaskApplyArgs", n=n, bw=bw, T=T, ldo=o.stride(1), lds=s.stride(1),
                                                     lde=ge.stride(1), bso=o.stride(0), bss=s.stride(0), bse=ge.stride(0), o=o,
                                                     s=s, ge=ge, go=go), _stream())
        return go, None


# --------------------------------------------------------------------------- wespeaker ResNet building blocks
def conv1x1_bigk(x, W2d, bias=None):
    """Conv1x1Fn for any number of input channels: > 1024 channels (what one tcgen05 launch covers) are contracted in chunks,
    the later chunks accumulating onto the first (autograd sees ordinary Conv1x1Fn nodes + adds)."""
    Kd = x.shape[1]
    if Kd <= 1024:
        return Conv1x1Fn.apply(x, W2d, bias, False, None)
    y = None
    for k0 in range(0, Kd, 1024):
        k1 = min(Kd, k0 + 1024)
        part = Conv1x1Fn.apply(x[:, k0:k1], W2d[:, k0:k1], bias if k0 == 0 else None, False, None)
        y = part if y is None else AddFn.apply(y, part)
    return y


def conv1x1_bigk_relu(x, W2d, bias):
    """relu(conv1x1) for any number of input channels: the ReLU is fused into the GEMM epilogue when one launch covers the
    contraction, otherwise applied to the accumulated sum."""
    if x.shape[1] <= 1024:
        return Conv1x1Fn.apply(x, W2d, bias, False, "relu")
    return as_act(UnaryFn.apply(conv1x1_bigk(x, W2d, bias), 0))


class AddFn(torch.autograd.Function):
    """a + b for two act tensors (the gradient is shared, not copied)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = as_act(a), as_act(b)
        out = new_act(a.shape[0], a.shape[1], a.shape[2], a.device)
        if a.shape != b.shape or a.stride() != b.stride():
            raise RuntimeError("add: layout mismatch")
        _lib.call("wesep_b200_add", _args("WesepAddArgs", count=a.shape[0] * a.stride(0), a=a, b=b, out=out), _stream())
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def _strides2(stride):
    """int or (sh, sw) -> (sh, sw)."""
    if isinstance(stride, (tuple, list)):
        return int(stride[0]), int(stride[1])
    return int(stride), int(stride)


def _i2c_args(n, C, H, W, stride, ldx, ct, **kw):
    """`ct`: the col-side tensor (its row / batch strides)."""
    sh, sw = _strides2(stride)
    return _args("WesepIm2colArgs", n=n, C=C, H=H, W=W, stride=sh, stride_w=0 if sw == sh else sw, Ho=(H - 1) // sh + 1,
                 Wo=(W - 1) // sw + 1, ldx=ldx, ldc=ct.stride(1), bsc=ct.stride(0), **kw)


class Im2Col3x3Fn(torch.autograd.Function):
    """[n, C, H*W] -> [n, 9C, Ho*Wo]: the patches of a 3x3 / pad 1 / stride s convolution (rows ordered (c, kh, kw) = the
    Conv2d weight viewed [Cout, Cin*9]); the adjoint gathers."""

    @staticmethod
    def forward(ctx, x, H, W, stride):
        x = as_act(x)
        n, C, HW = x.shape
        if HW != H * W:
            raise RuntimeError("im2col: H * W does not match the tensor")
        sh, sw = _strides2(stride)
        Ho, Wo = (H - 1) // sh + 1, (W - 1) // sw + 1
        Kp = (9 * C + 15) // 16 * 16                    # GEMM-friendly channel count: zero rows beyond 9 C (first conv: 9 -> 16)
        col = new_act(n, Kp, Ho * Wo, x.device)
        if Kp != 9 * C:
            col[:, 9 * C:].zero_()
        _lib.call("wesep_b200_im2col3x3_fwd", _i2c_args(n, C, H, W, stride, x.stride(1), col, x=x, col=col), _stream())
        ctx.meta = (n, C, H, W, stride, x.stride(1))
        return col

    @staticmethod
    def backward(ctx, gcol):
        n, C, H, W, stride, ldx = ctx.meta
        gcol = as_act(gcol)                             # [n, Kp, Ho*Wo]: rows >= 9 C are ignored
        gx = new_act(n, C, H * W, gcol.device)
        _lib.call("wesep_b200_im2col3x3_bwd", _i2c_args(n, C, H, W, stride, gx.stride(1), gcol, gcol=gcol, gx=gx), _stream())
        return gx, None, None, None


class Subsample2dFn(torch.autograd.Function):
    """x[..., ::s, ::s] of a [n, C, H*W] map (input of the 1x1 stride-s shortcut convolution)."""

    @staticmethod
    def forward(ctx, x, H, W, stride):
        x = as_act(x)
        n, C, _ = x.shape
        Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
        y = new_act(n, C, Ho * Wo, x.device)
        _lib.call("wesep_b200_subsample2d_fwd", _i2c_args(n, C, H, W, stride, x.stride(1), y, x=x, col=y), _stream())
        ctx.meta = (n, C, H, W, stride)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, C, H, W, stride = ctx.meta
        gy = as_act(gy)
        gx = new_act(n, C, H * W, gy.device)
        _lib.call("wesep_b200_subsample2d_bwd", _i2c_args(n, C, H, W, stride, gx.stride(1), gy, gcol=gy, gx=gx), _stream())
        return gx, None, None, None


class BnActFn(torch.autograd.Function):
    """y = act(BatchNorm(x) (+ res)) on [n, C, T] with batch statistics over (n, T) (nn.BatchNorm2d on a [n, C, H, W] map),
    act = ReLU or identity; running buffers updated like nn.BatchNorm (momentum, unbiased variance)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, res, relu, training, momentum, eps):
        x = as_act(x)
        n, C, T = x.shape
        dev = x.device
        stats = torch.zeros((C, 2), dtype=torch.float64, device=dev)
        st = _stream()
        if training:
            _lib.call("wesep_b200_bn2_stats", _args("WesepBn2Args", n=n, C=C, T=T, ld=x.stride(1), x=x, stats=stats), st)
        fin = _bn_finalize(stats, n * T, weight, bias, running_mean, running_var, momentum, eps, training)
        if res is not None:
            res = as_act(res)
            if res.stride() != x.stride():
                raise RuntimeError("bn: residual layout")
        y = new_act(n, C, T, dev)
        _lib.call("wesep_b200_bn2_apply", _args("WesepBn2Args", n=n, C=C, T=T, relu=int(relu), ld=x.stride(1), x=x, res=res, y=y,
                                                scale=fin[0], shift=fin[1]), st)
        ctx.relu, ctx.has_res, ctx.training = bool(relu), res is not None, bool(training)
        ctx.save_for_backward(x, y if relu else None, fin, _vec(weight))
        return y

    @staticmethod
    def backward(ctx, gy):
        if not ctx.training:
            raise NotImplementedError("BatchNorm backward with running statistics (eval mode) is not built")
        x, y, fin, w = ctx.saved_tensors
        n, C, T = x.shape
        gy = as_act(gy)
        if gy.stride() != x.stride():
            raise RuntimeError("bn backward: gradient layout")
        gx = new_act(n, C, T, x.device)
        gres = new_act(n, C, T, x.device) if ctx.has_res else None
        bsum = torch.zeros((C, 2), dtype=torch.float64, device=x.device)
        _lib.call("wesep_b200_bn2_bwd", _args("WesepBn2Args", n=n, C=C, T=T, relu=int(ctx.relu), ld=x.stride(1), count=float(n * T),
                                              x=x, y=y, gy=gy, mean=fin[2], rstd=fin[3], gamma=w, bsum=bsum, gx=gx, gres=gres),
                  _stream())
        bf = bsum.float()
        return gx, bf[:, 1].contiguous(), bf[:, 0].contiguous(), None, None, gres, None, None, None, None


class TstpFn(torch.autograd.Function):
    """wespeaker TSTP pooling: [n, R, T] -> [n, 2R] = (mean | sqrt(unbiased var + 1e-7)) over time."""

    @staticmethod
    def forward(ctx, x, eps=0.0):
        x = as_act(x)
        n, R, T = x.shape
        out = torch.empty((n, 2 * R), dtype=torch.float32, device=x.device)
        _lib.call("wesep_b200_tstp_fwd", _args("WesepTstpArgs", n=n, R=R, T=T, ld=x.stride(1), x=x, out=out, eps=float(eps)), _stream())
        ctx.eps = float(eps)
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        n, R, T = x.shape
        gx = new_act(n, R, T, x.device)
        _lib.call("wesep_b200_tstp_bwd", _args("WesepTstpArgs", n=n, R=R, T=T, ld=x.stride(1), x=x, gout=g.contiguous().float(),
                                               gx=gx, eps=ctx.eps), _stream())
        return gx, None


# --------------------------------------------------------------------------- pDPCCN building blocks
class Col2Im3x3Fn(torch.autograd.Function):
    """The adjoint of Im2Col3x3Fn as a forward op: [n, >= 9C, Ho*Wo] patch rows -> [n, C, H*W] (overlap-add of the 3x3
    patches).  With the pointwise GEMM in front this is nn.ConvTranspose2d(kernel 3, padding 1, stride (sh, sw))."""

    @staticmethod
    def forward(ctx, dcol, C, H, W, stride):
        dcol = as_act(dcol)
        n, Kp, HWo = dcol.shape
        sh, sw = _strides2(stride)
        if Kp < 9 * C or HWo != ((H - 1) // sh + 1) * ((W - 1) // sw + 1):
            raise RuntimeError("col2im: shape mismatch")
        y = new_act(n, C, H * W, dcol.device)
        _lib.call("wesep_b200_im2col3x3_bwd", _i2c_args(n, C, H, W, stride, y.stride(1), dcol, gcol=dcol, gx=y), _stream())
        ctx.meta = (n, C, H, W, stride, Kp)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, C, H, W, stride, Kp = ctx.meta
        gy = as_act(gy)
        sh, sw = _strides2(stride)
        gcol = new_act(n, Kp, ((H - 1) // sh + 1) * ((W - 1) // sw + 1), gy.device)
        if Kp != 9 * C:
            gcol[:, 9 * C:].zero_()
        _lib.call("wesep_b200_im2col3x3_fwd", _i2c_args(n, C, H, W, stride, gy.stride(1), gcol, x=gy, col=gcol), _stream())
        return gcol, None, None, None, None


def conv3x3(x, H, W, weight, bias, stride=1):
    """nn.Conv2d(Ci, Co, 3, stride, padding 1) on an act map [n, Ci, H*W]: im2col + tcgen05 pointwise GEMM."""
    col = Im2Col3x3Fn.apply(x, H, W, stride)
    w2 = weight.reshape(weight.shape[0], -1)
    if col.shape[1] != w2.shape[1]:
        w2 = torch.nn.functional.pad(w2, (0, col.shape[1] - w2.shape[1]))
    return conv1x1_bigk(col, w2, bias)


def conv_transpose3x3(x, H, W_out, weight, bias, stride=1):
    """nn.ConvTranspose2d(Ci, Co, 3, stride, padding 1) on [n, Ci, H*Wi] -> [n, Co, H_out*W_out]: pointwise GEMM with the
    weight viewed [Ci, Co*9] (transposed product) + col2im.  `weight` is the module's [Ci, Co, 3, 3] tensor."""
    Ci, Co = weight.shape[0], weight.shape[1]
    sh, sw = _strides2(stride)
    if sh != 1:
        raise RuntimeError("conv_transpose3x3: only stride (1, s) is built")
    w2 = weight.reshape(Ci, Co * 9)
    M = (Co * 9 + 3) // 4 * 4
    if M != Co * 9:
        w2 = torch.nn.functional.pad(w2, (0, M - Co * 9))
    dcol = Conv1x1Fn.apply(x, w2, None, True, None)                      # [n, Co*9 (+pad), H*Wi]
    y = Col2Im3x3Fn.apply(dcol, Co, H, W_out, stride)
    if bias is not None:
        y = RowAffineFn.apply(y, None, bias[None].expand(y.shape[0], Co))
    return y


class CatActFn(torch.autograd.Function):
    """torch.cat(xs, 1) written straight into an act tensor (torch.cat would return a dense [n, C, L] tensor that every
    kernel wrapper then copies into the padded layout); the gradient pieces are channel slices of gy (no copies)."""

    @staticmethod
    def forward(ctx, *xs):
        n, L = xs[0].shape[0], xs[0].shape[2]
        out = new_act(n, sum(x.shape[1] for x in xs), L, xs[0].device)
        c0, cuts = 0, []
        for x in xs:
            if x.shape[0] != n or x.shape[2] != L:
                raise RuntimeError("cat_act: shape mismatch")
            out[:, c0:c0 + x.shape[1]].copy_(x)
            cuts.append((c0, c0 + x.shape[1]))
            c0 += x.shape[1]
        ctx.cuts = cuts
        return out

    @staticmethod
    def backward(ctx, gy):
        return tuple(gy[:, a:b] for a, b in ctx.cuts)


def cat_act(xs):
    return xs[0] if len(xs) == 1 else CatActFn.apply(*xs)


def _gemm_rows_chunked(X, W2d, M, bias):
    """Y = W2d X (+ bias) with the contraction split into <= 1008-row chunks (the tcgen05 kernels take <= 1024), the later
    chunks accumulating in place (epilogue `+= R`)."""
    Kd = X.shape[1]
    Y = None
    for k0 in range(0, Kd, 1008):
        k1 = min(Kd, k0 + 1008)
        Wc = W2d[:, k0:k1].contiguous() if (k0 or k1 != Kd) else W2d
        if Y is None:
            Y = conv1x1_raw(X[:, k0:k1], Wc, False, M, bias=bias)
        else:
            conv1x1_raw(X[:, k0:k1], Wc, False, M, epi=2, R=Y, Y=Y)
    return Y


class DenseBlockFn(torch.autograd.Function):
    """DenseBlock.forward (wesep/modules/dpccn/convs.py:99-106): five Conv2dBlocks (3x3 / pad 1 / stride 1 -> ELU ->
    InstanceNorm) where conv k reads cat([x, y1 .. y_{k-1}]).  One patch buffer per block: every map is gathered (im2col) ONCE into
    its 9 C rows and conv k multiplies the row prefix; the backward accumulates W_k^T g_k into one gradient patch buffer (GEMM
    epilogue +=) and scatters each map's rows back once (col2im) — no concatenations, a third of the gather traffic.
    Arguments: x act [n, C0, T*F], T, F, then (weight_k [Co, Ci, 3, 3], bias_k) for k = 1..5."""

    @staticmethod
    def forward(ctx, x, T, F, *wb):
        x = as_act(x)
        n, C0, L = x.shape
        Ws, Bs = wb[0::2], wb[1::2]
        Cg = Ws[0].shape[0]
        Ctot = C0 + 4 * Cg
        if any(Ws[k].shape[1] != C0 + k * Cg for k in range(5)) or any(Ws[k].shape[0] != Cg for k in range(4)):
            raise RuntimeError("dense block: weight shapes")
        if (9 * C0) % 16 or (9 * Cg) % 16:
            raise RuntimeError("dense block: channel counts must be multiples of 16")
        dev = x.device
        col = new_act(n, 9 * Ctot, L, dev)
        st = _stream()
        zs, mrs = [], []
        feat, c_lo = x, 0
        for k in range(5):
            Cf = feat.shape[1]
            rows = col[:, 9 * c_lo:9 * (c_lo + Cf)]
            _lib.call("wesep_b200_im2col3x3_fwd", _i2c_args(n, Cf, T, F, 1, feat.stride(1), rows, x=feat, col=rows), st)
            c_lo += Cf
            Co = Ws[k].shape[0]
            z = _gemm_rows_chunked(col[:, :9 * c_lo], Ws[k].reshape(Co, 9 * c_lo), Co, _vec(Bs[k]))
            y = new_act(n, Co, L, dev)
            stats = torch.empty((n * Co, 2), dtype=torch.float64, device=dev)
            mr = torch.empty((n * Co, 2), dtype=torch.float32, device=dev)
            _lib.call("wesep_b200_elu_in_fwd", _args("WesepEluInArgs", rows=n * Co, L=L, ld=z.stride(1), mode=0, eps=1e-5, x=z, y=y,
                                                     stats=stats, mr=mr), st)
            zs.append(z)
            mrs.append(mr)
            feat = y
        ctx.meta = (n, C0, Cg, T, F, L)
        ctx.save_for_backward(col, *zs, *mrs, *Ws)
        return feat

    @staticmethod
    def backward(ctx, g):
        n, C0, Cg, T, F, L = ctx.meta
        saved = ctx.saved_tensors
        col, zs, mrs, Ws = saved[0], saved[1:6], saved[6:11], saved[11:16]
        dev = col.device
        st = _stream()
        g = as_act(g)
        Ctot = C0 + 4 * Cg
        dcol = new_act(n, 9 * Ctot, L, dev)
        grads = [None] * 10
        for k in range(4, -1, -1):
            z, mr = zs[k], mrs[k]
            Co = z.shape[1]
            c_hi = C0 + k * Cg
            if g.stride(1) != z.stride(1):
                raise RuntimeError("dense block backward: stride mismatch")
            gz = new_act(n, Co, L, dev)
            stats = torch.empty((n * Co, 2), dtype=torch.float64, device=dev)
            _lib.call("wesep_b200_elu_in_bwd", _args("WesepEluInArgs", rows=n * Co, L=L, ld=z.stride(1), mode=0, eps=1e-5, x=z,
                                                     stats=stats, mr=mr, gy=g, gx=gz), st)
            W2 = Ws[k].reshape(Co, 9 * c_hi)
            dW = torch.zeros_like(W2)
            conv1x1_dw_raw(gz, col[:, :9 * c_hi], dW)
            grads[2 * k] = dW.view(Ws[k].shape)
            grads[2 * k + 1] = rowsum_raw(gz).sum(0)
            rows = dcol[:, :9 * c_hi]
            if k == 4:
                conv1x1_raw(gz, W2, True, 9 * c_hi, Y=rows)
            else:
                conv1x1_raw(gz, W2, True, 9 * c_hi, epi=2, R=rows, Y=rows)
            # gradient of the map this convolution's input list ends with (y_k, or x when k == 0): its rows are complete now
            Cf, c_lo = (Cg, c_hi - Cg) if k > 0 else (C0, 0)
            gmap = new_act(n, Cf, L, dev)
            src = dcol[:, 9 * c_lo:9 * (c_lo + Cf)]
            _lib.call("wesep_b200_im2col3x3_bwd", _i2c_args(n, Cf, T, F, 1, gmap.stride(1), src, gcol=src, gx=gmap), st)
            g = gmap
        return (g if ctx.needs_input_grad[0] else None, None, None, *grads)


class EluInFn(torch.autograd.Function):
    """mode 0: InstanceNorm(ELU(x)); mode 1: ELU(InstanceNorm(x)); per (n, c) plane, eps 1e-5, no affine."""

    @staticmethod
    def forward(ctx, x, mode):
        x = as_act(x)
        n, C, L = x.shape
        dev = x.device
        y = new_act(n, C, L, dev)
        stats = torch.empty((n * C, 2), dtype=torch.float64, device=dev)
        mr = torch.empty((n * C, 2), dtype=torch.float32, device=dev)
        _lib.call("wesep_b200_elu_in_fwd", _args("WesepEluInArgs", rows=n * C, L=L, ld=x.stride(1), mode=int(mode), eps=1e-5,
                                                 x=x, y=y, stats=stats, mr=mr), _stream())
        ctx.mode = int(mode)
        ctx.save_for_backward(x, mr)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, mr = ctx.saved_tensors
        n, C, L = x.shape
        gy = as_act(gy)
        if gy.stride(1) != x.stride(1):
            raise RuntimeError("elu_in backward: stride mismatch")
        gx = new_act(n, C, L, x.device)
        stats = torch.empty((n * C, 2), dtype=torch.float64, device=x.device)
        _lib.call("wesep_b200_elu_in_bwd", _args("WesepEluInArgs", rows=n * C, L=L, ld=x.stride(1), mode=ctx.mode, eps=1e-5,
                                                 x=x, stats=stats, mr=mr, gy=gy, gx=gx), _stream())
        return gx, None


class DwConv1dFn(torch.autograd.Function):
    """nn.Conv1d(C, C, 3, padding=dil, dilation=dil, groups=C) — weight [C, 1, 3]."""

    @staticmethod
    def forward(ctx, x, weight, bias, dil):
        x = as_act(x)
        n, C, L = x.shape
        w = weight.reshape(C, 3).contiguous()
        y = new_act(n, C, L, x.device)
        _lib.call("wesep_b200_dwconv1d_fwd", _args("WesepDwConv1dArgs", n=n, C=C, L=L, ld=x.stride(1), dil=int(dil), x=x, w=w,
                                                   b=None if bias is None else _vec(bias), y=y), _stream())
        ctx.dil, ctx.wshape, ctx.has_bias = int(dil), weight.shape, bias is not None
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        n, C, L = x.shape
        gy = as_act(gy)
        if gy.stride(1) != x.stride(1):
            raise RuntimeError("dwconv1d backward: stride mismatch")
        gx = new_act(n, C, L, x.device)
        gw = torch.empty((C, 3), dtype=torch.float32, device=x.device)
        gb = torch.empty((C,), dtype=torch.float32, device=x.device) if ctx.has_bias else None
        _lib.call("wesep_b200_dwconv1d_bwd", _args("WesepDwConv1dArgs", n=n, C=C, L=L, ld=x.stride(1), dil=ctx.dil, x=x, w=w,
                                                   gy=gy, gx=gx, gw=gw, gb=gb), _stream())
        return gx, gw.view(ctx.wshape), gb, None


class AvgPool2dFn(torch.autograd.Function):
    """nn.AvgPool2d(k) on an act map [n, C, H*W] -> [n, C, (H//k)*(W//k)]."""

    @staticmethod
    def forward(ctx, x, H, W, k):
        x = as_act(x)
        n, C, _ = x.shape
        Ho, Wo = H // k, W // k
        y = new_act(n, C, Ho * Wo, x.device)
        _lib.call("wesep_b200_avgpool2d_fwd", _args("WesepPool2dArgs", rows=n * C, H=H, W=W, k=k, Ho=Ho, Wo=Wo, ldx=x.stride(1),
                                                    ldy=y.stride(1), x=x, y=y), _stream())
        ctx.meta = (n, C, H, W, k, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, C, H, W, k, Ho, Wo = ctx.meta
        gy = as_act(gy)
        gx = new_act(n, C, H * W, gy.device)
        _lib.call("wesep_b200_avgpool2d_bwd", _args("WesepPool2dArgs", rows=n * C, H=H, W=W, k=k, Ho=Ho, Wo=Wo, ldx=gx.stride(1),
                                                    ldy=gy.stride(1), gy=gy, gx=gx), _stream())
        return gx, None, None, None


class Upsample2dFn(torch.autograd.Function):
    """nn.Upsample(size=(Ho, Wo), mode="bilinear") on [n, C, Hi*Wi]."""

    @staticmethod
    def forward(ctx, x, Hi, Wi, Ho, Wo):
        x = as_act(x)
        n, C, _ = x.shape
        y = new_act(n, C, Ho * Wo, x.device)
        _lib.call("wesep_b200_upsample2d_fwd", _args("WesepUpsample2dArgs", rows=n * C, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo,
                                                     ldi=x.stride(1), ldo=y.stride(1), x=x, y=y), _stream())
        ctx.meta = (n, C, Hi, Wi, Ho, Wo)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, C, Hi, Wi, Ho, Wo = ctx.meta
        gy = as_act(gy)
        gx = new_act(n, C, Hi * Wi, gy.device)
        _lib.call("wesep_b200_upsample2d_bwd", _args("WesepUpsample2dArgs", rows=n * C, Hi=Hi, Wi=Wi, Ho=Ho, Wo=Wo,
                                                     ldi=gx.stride(1), ldo=gy.stride(1), gy=gy, gx=gx), _stream())
        return gx, None, None, None, None


class ColScaleFn(torch.autograd.Function):
    """y[n, c, t*F + f] = x[n, c, t*F + f] * s[n, f] (4-D multiply fusion, speaker.py:117-121)."""

    @staticmethod
    def forward(ctx, x, s, T, F):
        x = as_act(x)
        n, C, L = x.shape
        if L != T * F or s.shape != (n, F):
            raise RuntimeError("colscale: shape mismatch")
        s = s.contiguous().float()
        y = new_act(n, C, L, x.device)
        _lib.call("wesep_b200_colscale_fwd", _args("WesepColScaleArgs", n=n, C=C, T=T, F=F, ld=x.stride(1), x=x, s=s, y=y),
                  _stream())
        ctx.meta = (T, F)
        ctx.save_for_backward(x, s)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, s = ctx.saved_tensors
        T, F = ctx.meta
        n, C, L = x.shape
        gy = as_act(gy)
        if gy.stride(1) != x.stride(1):
            raise RuntimeError("colscale backward: stride mismatch")
        gx = new_act(n, C, L, x.device)
        gs = torch.empty((n, F), dtype=torch.float32, device=x.device)
        _lib.call("wesep_b200_colscale_bwd", _args("WesepColScaleArgs", n=n, C=C, T=T, F=F, ld=x.stride(1), x=x, s=s, gy=gy,
                                                   gx=gx, gs=gs), _stream())
        return gx, gs, None, None


# --------------------------------------------------------------------------- TF-GridNet building blocks
class HeadLnFn(torch.autograd.Function):
    """PReLU (slope per head or single) -> LayerNorm over (E, F) of every (b, h, t), affine gamma / beta [H, E, F], on an act map
    [B, H*E, T*F] (gridnet_block.py:229-284)."""

    @staticmethod
    def forward(ctx, x, alpha, gamma, beta, H, T, F, eps):
        x = as_act(x)
        B, HE, L = x.shape
        E = HE // H
        if HE != H * E or L != T * F or gamma.numel() != H * E * F or beta.numel() != H * E * F or alpha.numel() not in (1, H):
            raise RuntimeError("head_ln: shape mismatch")
        al, ga, be = _vec(alpha).float(), _vec(gamma).float(), _vec(beta).float()
        y = new_act(B, HE, L, x.device)
        mr = torch.empty((B, H, T, 2), dtype=torch.float32, device=x.device)
        a = _args("WesepHeadLnArgs", B=B, H=H, E=E, T=T, F=F, ld=x.stride(1), alpha_per_head=int(alpha.numel() == H and H > 1),
                  eps=float(eps), x=x, alpha=al, gamma=ga, beta=be, y=y, mr=mr)
        _lib.call("wesep_b200_head_ln_fwd", a, _stream())
        ctx.meta = (H, E, T, F, float(eps), alpha.shape, gamma.shape, beta.shape)
        ctx.save_for_backward(x, al, ga, be, mr)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, al, ga, be, mr = ctx.saved_tensors
        H, E, T, F, eps, sha, shg, shb = ctx.meta
        B, HE, L = x.shape
        gy = as_act(gy)
        if gy.stride(1) != x.stride(1):
            raise RuntimeError("head_ln backward: stride mismatch")
        gx = new_act(B, HE, L, x.device)
        dg = torch.empty(H * E * F, dtype=torch.float32, device=x.device)
        db = torch.empty(H * E * F, dtype=torch.float32, device=x.device)
        da = torch.empty(al.numel(), dtype=torch.float32, device=x.device)
        a = _args("WesepHeadLnArgs", B=B, H=H, E=E, T=T, F=F, ld=x.stride(1), alpha_per_head=int(al.numel() == H and H > 1),
                  eps=eps, x=x, alpha=al, gamma=ga, beta=be, mr=mr, gy=gy, gx=gx, dgamma=dg, dbeta=db, dalpha=da)
        _lib.call("wesep_b200_head_ln_bwd", a, _stream())
        return gx, da.view(sha), dg.view(shg), db.view(shb), None, None, None, None


class SoftmaxFn(torch.autograd.Function):
    """softmax(scale * x, dim=-1) over the C valid columns of an act tensor [n, R, C]; padding columns come out zero."""

    @staticmethod
    def forward(ctx, x, scale):
        x = as_act(x)
        n, R, C = x.shape
        y = new_act(n, R, C, x.device)
        _lib.call("wesep_b200_softmax_fwd", _args("WesepSoftmaxArgs", rows=n * R, C=C, ld=x.stride(1), scale=float(scale), x=x,
                                                  y=y), _stream())
        ctx.scale = float(scale)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        y, = ctx.saved_tensors
        n, R, C = y.shape
        gy = as_act(gy)
        if gy.stride(1) != y.stride(1):
            raise RuntimeError("softmax backward: stride mismatch")
        gx = new_act(n, R, C, y.device)
        _lib.call("wesep_b200_softmax_bwd", _args("WesepSoftmaxArgs", rows=n * R, C=C, ld=y.stride(1), scale=ctx.scale, y=y,
                                                  gy=gy, gx=gx), _stream())
        return gx, None


def row_std(x):
    """(torch.std(x, dim=1), its reciprocal) for a [n, L] signal batch (unbiased; data only, no gradient)."""
    x, ld = _sig(x)
    n, L = x.shape
    sd = torch.empty(n, dtype=torch.float32, device=x.device)
    inv = torch.empty(n, dtype=torch.float32, device=x.device)
    _lib.call("wesep_b200_rowstd", _args("WesepRowStdArgs", n=n, L=L, ld=ld, x=x, std=sd, inv_std=inv), _stream())
    return sd, inv


def attention_rows(Qm, Kt, Vm):
    """softmax(Qm Kt / sqrt(d)) Vm for ONE (batch, head) pair (gridnet_block.py:213-215) on the pointwise GEMMs:
    Qm [T, d] dense, Kt [d, T], Vm [T, dv] -> [T, dv].  The score and the probability matrices are act tensors whose padded row
    length (a multiple of 32) is used as the contraction length of the second product (padding columns are zero)."""
    T, d = Qm.shape
    Tp, dp = (T + 3) // 4 * 4, (d + 3) // 4 * 4                     # GEMM contraction / output-row counts: multiples of 4
    pad = torch.nn.functional.pad
    Kx = as_act(pad(Kt, (0, 0, 0, dp - d))[None])                   # [1, dp, T]
    S = Conv1x1Fn.apply(Kx, pad(Qm, (0, dp - d, 0, Tp - T)).contiguous(), None, False, None)   # [1, Tp(t), T(s)] = Qm @ Kt
    P = SoftmaxFn.apply(S, 1.0 / (d ** 0.5))                        # (the Tp - T extra rows are uniform and dropped below)
    ld = P.stride(1)
    if P.stride(0) != Tp * ld:
        raise RuntimeError("attention: probability matrix layout")
    Pw = _ActAsMatrixFn.apply(P)
    dv = Vm.shape[1]
    Vx = new_act(1, ld, dv, Vm.device, zero=True)                   # contraction rows padded with zeros
    Vx = _RowsIntoFn.apply(Vx, Vm)
    return Conv1x1Fn.apply(Vx, Pw, None, False, None)[0, :T]        # [T, dv]


class _ActAsMatrixFn(torch.autograd.Function):
    """act [1, R, C] -> the dense [R, ld] matrix over the same memory (padding columns included); the gradient of the
    padding columns is dropped."""

    @staticmethod
    def forward(ctx, P):
        _, R, C = P.shape
        ld = P.stride(1)
        ctx.C = C
        return P.detach().as_strided((R, ld), (ld, 1))

    @staticmethod
    def backward(ctx, g):
        return g[:, :ctx.C].unsqueeze(0)


class _RowsIntoFn(torch.autograd.Function):
    """buf[0, :T] = rows (buf zero-filled act [1, Kp, dv]); gradient = the same slice."""

    @staticmethod
    def forward(ctx, buf, rows):
        buf[0, :rows.shape[0]].copy_(rows)
        ctx.T = rows.shape[0]
        ctx.mark_dirty(buf)
        return buf

    @staticmethod
    def backward(ctx, g):
        return None, g[0, :ctx.T]


# --------------------------------------------------------------------------- wespeaker ECAPA-TDNN building blocks
class Im2Col1dFn(torch.autograd.Function):
    """[n, C, T] -> [n, C*K (padded to a multiple of 4), T]: the taps of nn.Conv1d(kernel K odd, dilation d, padding d (K-1)/2);
    rows ordered (c, k) = the Conv1d weight viewed [Cout, Cin*K]."""

    @staticmethod
    def forward(ctx, x, K, dil):
        x = as_act(x)                                   # dense [n, C, ld] (a channel slice of a wider tensor is copied)
        n, C, T = x.shape
        Kp = (C * K + 3) // 4 * 4
        col = new_act(n, Kp, T, x.device)
        if Kp != C * K:
            col[:, C * K:].zero_()
        _lib.call("wesep_b200_im2col1d_fwd", _args("WesepIm2col1dArgs", n=n, C=C, T=T, K=int(K), dil=int(dil), ldx=x.stride(1),
                                                   ldc=col.stride(1), bsc=col.stride(0), x=x, col=col), _stream())
        ctx.meta = (n, C, T, int(K), int(dil))
        return col

    @staticmethod
    def backward(ctx, gcol):
        n, C, T, K, dil = ctx.meta
        gcol = as_act(gcol)
        gx = new_act(n, C, T, gcol.device)
        _lib.call("wesep_b200_im2col1d_bwd", _args("WesepIm2col1dArgs", n=n, C=C, T=T, K=K, dil=dil, ldx=gx.stride(1),
                                                   ldc=gcol.stride(1), bsc=gcol.stride(0), gcol=gcol, gx=gx), _stream())
        return gx, None, None


class Unfold1dFn(torch.autograd.Function):
    """F.unfold(x[..., None], (K, 1), stride=(hs, 1)) on an act tensor [n, C, T] -> [n, C*K, (T - K) // hs + 1] (rows (c, k))."""

    @staticmethod
    def forward(ctx, x, K, hs):
        x = as_act(x)
        n, C, T = x.shape
        L = (T - K) // hs + 1
        col = new_act(n, C * K, L, x.device)
        _lib.call("wesep_b200_im2col1d_fwd", _args("WesepIm2col1dArgs", n=n, C=C, T=T, K=int(K), dil=1, ldx=x.stride(1),
                                                   ldc=col.stride(1), bsc=col.stride(0), x=x, col=col, unfold=1, stride=int(hs)), _stream())
        ctx.meta = (n, C, T, int(K), int(hs))
        return col

    @staticmethod
    def backward(ctx, gcol):
        n, C, T, K, hs = ctx.meta
        return _fold1d(as_act(gcol), n, C, T, K, hs), None, None


def _fold1d(col, n, C, T, K, hs):
    y = new_act(n, C, T, col.device)
    _lib.call("wesep_b200_im2col1d_bwd", _args("WesepIm2col1dArgs", n=n, C=C, T=T, K=K, dil=1, ldx=y.stride(1), ldc=col.stride(1),
                                               bsc=col.stride(0), gcol=col, gx=y, unfold=1, stride=hs), _stream())
    return y


class Fold1dFn(torch.autograd.Function):
    """The adjoint of Unfold1dFn as a forward op: [n, C*K, L] -> [n, C, T] overlap-add (the data movement of
    nn.ConvTranspose1d(kernel K, stride hs) after its pointwise product)."""

    @staticmethod
    def forward(ctx, col, C, T, K, hs):
        col = as_act(col)
        n = col.shape[0]
        if col.shape[1] != C * K or col.shape[2] != (T - K) // hs + 1:
            raise RuntimeError("fold1d: shape mismatch")
        ctx.meta = (n, C, T, int(K), int(hs))
        return _fold1d(col, n, C, T, int(K), int(hs))

    @staticmethod
    def backward(ctx, gy):
        n, C, T, K, hs = ctx.meta
        gy = as_act(gy)
        gcol = new_act(n, C * K, (T - K) // hs + 1, gy.device)
        _lib.call("wesep_b200_im2col1d_fwd", _args("WesepIm2col1dArgs", n=n, C=C, T=T, K=K, dil=1, ldx=gy.stride(1),
                                                   ldc=gcol.stride(1), bsc=gcol.stride(0), x=gy, col=gcol, unfold=1, stride=hs), _stream())
        return gcol, None, None, None, None


def conv1d_k(x, weight, bias, dil=1, act=None):
    """nn.Conv1d(Ci, Co, K, dilation=dil, padding=dil*(K-1)//2) (+ ReLU in the GEMM epilogue) on an act tensor."""
    Co, Ci, K = weight.shape
    if K == 1:
        return Conv1x1Fn.apply(x, weight[:, :, 0], bias, False, act)
    col = Im2Col1dFn.apply(as_act(x), K, dil)
    w2 = weight.reshape(Co, Ci * K)
    if col.shape[1] != w2.shape[1]:
        w2 = torch.nn.functional.pad(w2, (0, col.shape[1] - w2.shape[1]))
    if col.shape[1] > 1024:
        if act is not None:
            raise RuntimeError("conv1d_k: fused activation with more than 1024 gathered channels")
        return conv1x1_bigk(col, w2, bias)
    return Conv1x1Fn.apply(col, w2, bias, False, act)


class UnaryFn(torch.autograd.Function):
    """Elementwise ReLU (mode 0) / sigmoid (mode 1) on a contiguous tensor."""

    @staticmethod
    def forward(ctx, x, mode):
        _check_cuda(x)
        x = x.contiguous().float()
        y = torch.empty_like(x)
        _lib.call("wesep_b200_unary_fwd", _args("WesepUnaryArgs", count=x.numel(), mode=int(mode), x=x, y=y), _stream())
        ctx.mode = int(mode)
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, gy):
        (y,) = ctx.saved_tensors
        gy = gy.contiguous().float()
        gx = torch.empty_like(y)
        _lib.call("wesep_b200_unary_bwd", _args("WesepUnaryArgs", count=y.numel(), mode=ctx.mode, y=y, gy=gy, gx=gx), _stream())
        return gx, None


class AstpFn(torch.autograd.Function):
    """Attentive statistics: x, alpha (softmax over time) [n, C, T] -> [n, 2C] = (sum alpha x | sqrt(clamp(sum alpha x^2 - mean^2)))."""

    @staticmethod
    def forward(ctx, x, alpha):
        x, alpha = as_act(x), as_act(alpha)
        n, C, T = x.shape
        if alpha.shape != x.shape or alpha.stride(1) != x.stride(1):
            raise RuntimeError("astp: layout mismatch")
        out = torch.empty((n, 2 * C), dtype=torch.float32, device=x.device)
        _lib.call("wesep_b200_astp_fwd", _args("WesepAstpArgs", n=n, C=C, T=T, ld=x.stride(1), x=x, alpha=alpha, out=out), _stream())
        ctx.save_for_backward(x, alpha)
        return out

    @staticmethod
    def backward(ctx, g):
        x, alpha = ctx.saved_tensors
        n, C, T = x.shape
        gx, ga = new_act(n, C, T, x.device), new_act(n, C, T, x.device)
        _lib.call("wesep_b200_astp_bwd", _args("WesepAstpArgs", n=n, C=C, T=T, ld=x.stride(1), x=x, alpha=alpha,
                                               gout=g.contiguous().float(), gx=gx, galpha=ga), _stream())
        return gx, ga


class ColVecMulFn(torch.autograd.Function):
    """y[r, t] = x[r, t] * v[t] for a [rows, L] signal batch and a constant vector (the iSTFT envelope 1 / sum_k w^2)."""

    @staticmethod
    def forward(ctx, x, v):
        _check_cuda(x, v)
        if x.dim() != 2 or x.stride(1) != 1 or x.dtype != torch.float32:
            x = x.contiguous().float()
        rows, L = x.shape
        v = v.contiguous().float()
        if v.numel() != L:
            raise RuntimeError("colvec_mul: vector length")
        y = torch.empty((rows, L), dtype=torch.float32, device=x.device)
        _lib.call("wesep_b200_colvec_mul", _args("WesepColVecArgs", rows=rows, L=L, x=x, ldx=x.stride(0), v=v, y=y, ldy=L), _stream())
        ctx.save_for_backward(v)
        return y

    @staticmethod
    def backward(ctx, g):
        (v,) = ctx.saved_tensors
        if g.stride(1) != 1 or g.dtype != torch.float32:
            g = g.contiguous().float()
        rows, L = g.shape
        gx = torch.empty((rows, L), dtype=torch.float32, device=g.device)
        _lib.call("wesep_b200_colvec_mul", _args("WesepColVecArgs", rows=rows, L=L, x=g, ldx=g.stride(0), v=v, y=gx, ldy=L), _stream())
        return gx, None
