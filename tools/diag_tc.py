"""Diagnostic for the tcgen05 GEMM: structured inputs first (to decode layout errors), then random parity."""
import os, sys, json, math
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from wesep_b200 import _lib, ops

DEV = "cuda"


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def run(n, Kd, M, T, w_trans=False, mode="rand", epi=0, pro=0):
    g = torch.Generator().manual_seed(n * 1000 + Kd + M + T)
    x = ops.new_act(n, Kd, T, DEV)
    if mode == "ident":
        base = torch.arange(Kd, dtype=torch.float32)[None, :, None] + torch.arange(T, dtype=torch.float32)[None, None, :] / 1024.0
        x.copy_((base + 1000.0 * torch.arange(n, dtype=torch.float32)[:, None, None]).to(DEV))
        W = torch.zeros(M, Kd)
        for o in range(M):
            W[o, o % Kd] = 1.0
    else:
        x.copy_(torch.randn(n, Kd, T, generator=g).to(DEV))
        W = torch.randn(M, Kd, generator=g) / math.sqrt(Kd)
    W = W.to(DEV)
    Wp = W.t().contiguous() if w_trans else W
    bias = torch.randn(M, generator=g).to(DEV) if mode != "ident" else None
    kw = {}
    R = None
    if epi == 2:
        R = ops.new_act(n, M, T, DEV)
        R.copy_(torch.randn(n, M, T, generator=g).to(DEV))
        kw["R"] = R
    stats = None
    if epi == 0 and mode != "ident":
        stats = torch.zeros(n, 2, dtype=torch.float64, device=DEV)
        kw["out_stats"] = stats
        kw["out_alpha"] = torch.tensor([0.25], device=DEV)
    if pro == 2:
        kw.update(alpha=torch.tensor([0.2], device=DEV), ch_scale=(1 + 0.1 * torch.randn(Kd, generator=g)).to(DEV),
                  ch_shift=(0.1 * torch.randn(Kd, generator=g)).to(DEV), row_stats=torch.tensor([[0.3 * Kd * T, 1.5 * Kd * T]] * n, dtype=torch.float64, device=DEV),
                  stat_count=Kd * T, stat_eps=1e-5)
    res = {}
    outs = {}
    for backend in (0, 1):
        _lib.set_gemm_backend(backend)
        if stats is not None:
            stats.zero_()
        y = ops.conv1x1_raw(x, Wp, w_trans, M, bias=bias, epi=epi, pro=pro, **kw)
        torch.cuda.synchronize()
        outs[backend] = y.clone()
        if stats is not None:
            res[f"stats{backend}"] = stats.cpu().tolist()
    _lib.set_gemm_backend(0)
    xin = x.double()
    if pro == 2:
        mu = 0.3
        r = 1 / math.sqrt(1.5 - 0.09 + 1e-5)
        xin = torch.where(xin > 0, xin, 0.2 * xin)
        xin = kw["ch_scale"].double()[None, :, None] * (xin - mu) * r + kw["ch_shift"].double()[None, :, None]
    ref = torch.einsum("mk,nkt->nmt", W.double(), xin)
    if bias is not None:
        ref = ref + bias.double()[None, :, None]
    if R is not None:
        ref = ref + R.double()
    res["err_mma"] = rel(outs[0], ref)
    res["err_tc"] = rel(outs[1], ref)
    if res["err_tc"] > 1e-4:
        y, r_ = outs[1], ref
        e = (y.double() - r_).abs()
        res["frac_bad"] = float((e > 1e-3 * r_.abs().mean()).float().mean())
        res["y_sample"] = y[0, :4, :8].cpu().tolist()
        res["y_absmax"] = float(y.abs().max())
        res["y_minus_bias_absmax"] = float((y - (bias[None, :, None] if bias is not None else 0)).abs().max())
        res["ref_sample"] = r_[0, :4, :8].cpu().tolist()
        # error by (o block of 32, t block of 32)
        nb_o, nb_t = min(M // 32, 8), min((T + 31) // 32, 12)
        em = [[float(e[0, 32 * i:32 * i + 32, 32 * j:32 * j + 32].mean()) for j in range(nb_t)] for i in range(nb_o)]
        res["err_blocks_o32_t32"] = em
        res["nan"] = int(torch.isnan(y).sum())
    return res


def run_dw(n, M, N, T, pro=0, per_row=False, mode="rand"):
    g = torch.Generator().manual_seed(7)
    A = ops.new_act(n, M, T, DEV); B = ops.new_act(n, N, T, DEV)
    if mode == "ident":
        # A[o][t] = 1 if t == o else 0 (t < M); B[c][t] = c + t/1024  ->  C[o][c] = B[c][o]
        a = torch.zeros(n, M, T); idx = torch.arange(min(M, T)); a[:, idx, idx] = 1.0
        A.copy_(a.to(DEV))
        b = torch.arange(N, dtype=torch.float32)[None, :, None] + torch.arange(T, dtype=torch.float32)[None, None, :] / 1024.0
        B.copy_(b.expand(n, N, T).to(DEV))
    else:
        A.copy_(torch.randn(n, M, T, generator=g).to(DEV)); B.copy_(torch.randn(n, N, T, generator=g).to(DEV))
    alpha = torch.tensor([0.3], device=DEV)
    res = {}
    outs = {}
    for backend in (0, 1):
        _lib.set_gemm_backend(backend)
        C = torch.zeros((n, M, N) if per_row else (M, N), device=DEV)
        ops.conv1x1_dw_raw(A, B, C, per_row=per_row, pro_b=pro, alpha_b=alpha if pro else None)
        torch.cuda.synchronize()
        outs[backend] = C
    _lib.set_gemm_backend(1)
    Bd = B.double()
    if pro:
        Bd = torch.where(Bd > 0, Bd, 0.3 * Bd)
    ref = torch.einsum("nmt,nkt->nmk", A.double(), Bd)
    if not per_row:
        ref = ref.sum(0)
    res["err_mma"] = rel(outs[0], ref); res["err_tc"] = rel(outs[1], ref)
    if res["err_tc"] > 1e-4:
        y = outs[1] if not per_row else outs[1][0]
        r_ = ref if not per_row else ref[0]
        res["y_sample"] = y[:4, :8].cpu().tolist(); res["ref_sample"] = r_[:4, :8].cpu().tolist()
        e = (y.double() - r_).abs()
        res["err_blocks_o32_c32"] = [[float(e[32 * i:32 * i + 32, 32 * j:32 * j + 32].mean()) for j in range(8)] for i in range(4)]
        res["nan"] = int(torch.isnan(y).sum())
    return res


def run_flags():
    """does the tensor core truncate tf32 inputs? compare flag 0 (hi stored) vs flag 1 (raw tile as hi)"""
    out = {}
    for fl in (0, 1):
        _lib.lib().wesep_b200_set_tc_flags(fl)
        out[f"flags{fl}"] = run(2, 256, 512, 6399)["err_tc"]
    _lib.lib().wesep_b200_set_tc_flags(0)
    return out


if __name__ == "__main__":
    cases = [dict(n=1, Kd=128, M=128, T=256, mode="ident"), dict(n=1, Kd=128, M=128, T=256, mode="ident", w_trans=True),
             dict(n=2, Kd=128, M=256, T=700, mode="ident"),
             dict(n=1, Kd=128, M=128, T=256), dict(n=2, Kd=256, M=512, T=6399), dict(n=2, Kd=256, M=512, T=6399, w_trans=True),
             dict(n=3, Kd=512, M=256, T=1000, epi=2), dict(n=2, Kd=512, M=256, T=6399, epi=2, pro=2),
             dict(n=32, Kd=256, M=512, T=6399)]
    out = {}
    for c in [dict(n=1, M=128, N=256, T=256, mode="ident"), dict(n=1, M=128, N=256, T=256), dict(n=2, M=256, N=512, T=6399, pro=1, per_row=True),
              dict(n=32, M=512, N=256, T=6399)]:
        try:
            r = run_dw(**c)
        except Exception as ex:  # noqa
            r = {"exception": repr(ex)}
        out["dw" + json.dumps(c)] = r
        print("dw", json.dumps(c), json.dumps(r)[:1200], flush=True)
    try:
        out["flags"] = run_flags()
    except Exception as ex:  # noqa
        out["flags"] = {"exception": repr(ex)}
    print("flags", out["flags"], flush=True)
    cases = cases[3:6]
    for c in cases:
        try:
            r = run(**c)
        except Exception as ex:  # noqa
            r = {"exception": repr(ex)}
        out[json.dumps(c)] = r
        print(json.dumps(c), json.dumps(r)[:1500], flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "diag_tc.json"), "w"), indent=1)
