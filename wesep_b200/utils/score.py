"""Evaluation scoring on the device — reference wesep/utils/score.py:7-36 (cal_SISNR / cal_SISNRi)
and the peak rule of wesep/bin/infer.py:124-129, batched: one C-ABI call per batch of rows
instead of a numpy round trip per utterance.  PESQ / STOI (score.py:39-62) are third-party CPU
packages (pesq, pystoi) and are out of scope of this path.
"""
import torch

from wesep_b200 import _lib, ops


def _rows(x):
    if x.dim() == 1:
        x = x[None]
    if x.dim() != 2:
        raise RuntimeError("score: expected [rows, samples]")
    ops._check_cuda(x)
    x = x.float()
    return x if x.stride(1) == 1 else x.contiguous()


def score_batch(est, ref, mix, lengths=None, peak_norm=True):
    """est / ref / mix: [n, T*] CUDA waves (T may differ per tensor; the common prefix is scored,
    infer.py:147-152).  Returns (est_out [n, T_est], sisnr [n], sisnri [n], normed [1] int32),
    all on the device; est_out is a scaled copy when the peak rule fires (est itself is untouched)."""
    est, ref, mix = _rows(est), _rows(ref), _rows(mix)
    n = est.shape[0]
    if ref.shape[0] != n or mix.shape[0] != n:
        raise RuntimeError("score: row count mismatch")
    Le = est.shape[1]
    end = min(Le, ref.shape[1], mix.shape[1])
    dev = est.device
    if lengths is None:
        lens = torch.full((n,), end, dtype=torch.int32, device=dev) if end != Le else None
    else:
        lens = torch.as_tensor(lengths, dtype=torch.int32, device=dev).clamp(max=end)
    out = est.clone() if peak_norm else est
    # rows shorter than est are read only below len[r] <= end, so their own strides are enough
    ws = torch.empty(_lib.lib().wesep_b200_score_ws_bytes(n) // 8, dtype=torch.float64, device=dev)
    sisnr = torch.empty(n, dtype=torch.float32, device=dev)
    sisnri = torch.empty(n, dtype=torch.float32, device=dev)
    normed = torch.zeros(1, dtype=torch.int32, device=dev)
    if ref.shape[1] < Le:
        ref = torch.nn.functional.pad(ref, (0, Le - ref.shape[1]))
    if mix.shape[1] < Le:
        mix = torch.nn.functional.pad(mix, (0, Le - mix.shape[1]))
    a = ops._args("WesepScoreArgs", n=n, L=Le, len=lens, est=out, ld_est=out.stride(0), ref=ref, ld_ref=ref.stride(0),
                  mix=mix, ld_mix=mix.stride(0), peak_norm=int(bool(peak_norm)), ws=ws, sisnr=sisnr, sisnri=sisnri,
                  normed=normed)
    _lib.call("wesep_b200_score", a, ops._stream())
    return out, sisnr, sisnri, normed


def cal_SISNRi(est, ref, mix, eps=1e-8):
    """score.py:24-36 for one pair or a batch of rows (CUDA tensors); returns (sisnr, sisnr - sisnr_mix)."""
    if eps != 1e-8:
        raise RuntimeError("cal_SISNRi: the kernel fixes eps = 1e-8 (the reference default)")
    if not (est.shape[-1] == ref.shape[-1] == mix.shape[-1]):
        raise AssertionError("cal_SISNRi: lengths differ")     # score.py:32 assert
    _, s, d, _ = score_batch(est, ref, mix, peak_norm=False)
    return (s, d) if est.dim() == 2 else (s[0], d[0])


def cal_SISNR(est, ref, eps=1e-8):
    """score.py:7-21 (CUDA tensors)."""
    return cal_SISNRi(est, ref, ref, eps)[0]
