"""In-model enrollment features for `spk_feat: False`, `feat_type: "consistent"` — reference wesep/models/bsrnn.py:231-241,
343-351 (the same block in dpccn.py:88-99,233-240 and tfgridnet.py:144-155,253-260): PreEmphasis
(wesep/modules/common/speaker.py:10-23) -> torchaudio.transforms.MelSpectrogram(sample_rate, n_fft = win, win_length = win,
hop_length = stride, f_min = 20, window_fn = hamming_window, n_mels) + 1e-8 -> log -> minus the mean over frames ->
[B, frames, n_mels], all under no_grad.  The modules carry the same buffers under the same names as the reference's
(`preEmphasis.flipped_filter`, `spk_encoder.spectrogram.window`, `spk_encoder.mel_scale.fb`); the arithmetic runs on
libwesep_b200 (pre-emphasis, framing + windowed-DFT GEMM, power, mel GEMM, log + mean kernels)."""
import math

import torch
import torch.nn as nn

from wesep_b200 import _lib, ops


def melscale_fbanks_htk(n_freqs, f_min, f_max, n_mels, sample_rate):
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk") restated (fp32 torch ops): [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


class PreEmphasis(nn.Module):
    """speaker.py:10-23 (buffer kept for state_dict compatibility; the filter runs in wesep_b200_preemphasis)."""

    def __init__(self, coef=0.97):
        super().__init__()
        self.coef = coef
        self.register_buffer("flipped_filter", torch.FloatTensor([-self.coef, 1.0]).unsqueeze(0).unsqueeze(0))


class _Spectrogram(nn.Module):
    def __init__(self, n_fft):
        super().__init__()
        self.register_buffer("window", torch.hamming_window(n_fft))


class _MelScale(nn.Module):
    def __init__(self, n_freqs, n_mels, sample_rate, f_min):
        super().__init__()
        self.register_buffer("fb", melscale_fbanks_htk(n_freqs, f_min, float(sample_rate // 2), n_mels, sample_rate))


class MelSpectrogram(nn.Module):
    """torchaudio.transforms.MelSpectrogram with the arguments the reference passes (parameter container)."""

    def __init__(self, sample_rate=16000, n_fft=512, hop_length=128, f_min=20.0, n_mels=80):
        super().__init__()
        self.sample_rate, self.n_fft, self.hop_length, self.n_mels = sample_rate, n_fft, hop_length, n_mels
        self.spectrogram = _Spectrogram(n_fft)
        self.mel_scale = _MelScale(n_fft // 2 + 1, n_mels, sample_rate, f_min)


def _tables(enc, device):
    key = str(device)
    cache = enc.__dict__.setdefault("_table_cache", {})
    if key not in cache:
        win, F = enc.n_fft, enc.n_fft // 2 + 1
        w = enc.spectrogram.window.detach().cpu().double()
        k = torch.arange(win, dtype=torch.float64)
        f = torch.arange(F, dtype=torch.float64)
        ang = 2.0 * math.pi * f[:, None] * k[None, :] / win
        R = (2 * F + 3) // 4 * 4
        fwd = torch.zeros(R, win, dtype=torch.float64)
        fwd[:F] = torch.cos(ang) * w
        fwd[F:2 * F] = -torch.sin(ang) * w
        Fp = (F + 3) // 4 * 4
        fbT = torch.zeros(enc.n_mels, Fp)
        fbT[:, :F] = enc.mel_scale.fb.detach().cpu().t()
        cache[key] = (fwd.float().to(device).contiguous(), R, fbT.to(device).contiguous(), Fp)
    return cache[key]


@torch.no_grad()
def consistent_features(wave, pre, enc):
    """wave [B, L] CUDA -> [B, 1 + L // hop, n_mels]: the block under `with torch.no_grad()` at bsrnn.py:345-351."""
    ops._check_cuda(wave)
    x = wave.float()
    x = x if x.stride(1) == 1 else x.contiguous()
    B, L = x.shape
    dev = x.device
    win, hop, M = enc.n_fft, enc.hop_length, enc.n_mels
    fwd, R, fbT, Fp = _tables(enc, dev)
    F = win // 2 + 1
    st = ops._stream()
    y = torch.empty((B, L), dtype=torch.float32, device=dev)
    _lib.call("wesep_b200_preemphasis", ops._args("WesepPreEmphArgs", n=B, L=L, x=x, ldx=x.stride(0), coef=float(pre.coef), y=y,
                                                  ldy=L), st)
    T = 1 + L // hop
    pad = win // 2
    yp = torch.cat([y[:, 1:pad + 1].flip(1), y, y[:, L - pad - 1:L - 1].flip(1)], 1).contiguous()     # center=True, reflect
    spec = ops.conv1x1_raw(ops.frames_raw(yp, win, T, hop), fwd, False, R)                           # [B, R, T]
    pw = ops.new_act(B, Fp, T, dev, zero=Fp != F)
    _lib.call("wesep_b200_power_spec", ops._args("WesepPowerSpecArgs", n=B, F=F, T=T, spec=spec, ld=spec.stride(1), bs=spec.stride(0),
                                                 pw=pw, ldp=pw.stride(1), bsp=pw.stride(0)), st)
    mel = ops.conv1x1_raw(pw, fbT, False, M)                                                          # [B, n_mels, T]
    out = torch.empty((B, T, M), dtype=torch.float32, device=dev)
    _lib.call("wesep_b200_log_cmn", ops._args("WesepLogCmnArgs", n=B, M=M, T=T, mel=mel, ld=mel.stride(1), eps=1e-8, out=out), st)
    return out
