"""ORACLE (test infrastructure) — CPU restatement of the pBSRNN forward.

Plain torch ops in the dtype of the inputs (fp32 or fp64), parameters taken from a ``state_dict`` with the
reference's key names (wesep/models/bsrnn.py).  The recurrence, the STFT / iSTFT and the GroupNorm are
written out explicitly (no nn.LSTM / torch.stft / F.group_norm) so that this file states the arithmetic the
CUDA path has to reproduce; ``tests/test_oracle_bsrnn.py`` pins it against the real reference module and
against torch.stft / torch.istft.  Backward = torch autograd of this forward.

Scope: ``joint_training=False`` (the separator consumes a given 256-d speaker embedding; the wespeaker
ResNet34 of bsrnn.yaml:58-64 is an external package), ``use_spk_transform=False``,
``spk_fuse_type`` in {multiply, additive, concat}, ``multi_fuse`` False or True.
"""
import math

import torch

EPS_GN = float(torch.finfo(torch.float32).eps)   # nn.GroupNorm(1, C, eps) in bsrnn.py:24,256,274


def band_widths(sr=16000, win=512):
    """wesep/models/bsrnn.py:228-242 -> [3]*15 + [6]*10 + [16]*5 + [64] + [8] for 16 kHz / 512."""
    enc_dim = win // 2 + 1
    bw100 = int(math.floor(100 / (sr / 2.0) * enc_dim))
    bw200 = int(math.floor(200 / (sr / 2.0) * enc_dim))
    bw500 = int(math.floor(500 / (sr / 2.0) * enc_dim))
    bw2k = int(math.floor(2000 / (sr / 2.0) * enc_dim))
    bands = [bw100] * 15 + [bw200] * 10 + [bw500] * 5 + [bw2k]
    bands.append(enc_dim - sum(bands))
    return bands


def hann(win, dtype):
    """The reference builds the periodic Hann window 0.5 - 0.5 cos(2 pi k / win) in fp32 and THEN casts it to the
    signal dtype (`torch.hann_window(self.win)...type(wav_input.type())`, bsrnn.py:313,386), so an fp64 run still
    uses fp32-rounded window values; mirrored here so fp64 comparisons with the reference are exact."""
    return torch.hann_window(win, dtype=torch.float32).to(dtype)


def stft(x, win=512, hop=128, window=None):
    """torch.stft(x, n_fft=win, hop_length=hop, window=hann, center=True (reflect), onesided, not normalised,
    return_complex=True), bsrnn.py:309-316.  x [B, L] -> (re, im) each [B, win/2+1, 1 + L//hop]."""
    B, L = x.shape
    pad = win // 2
    left = x[:, 1:pad + 1].flip(1)
    right = x[:, L - pad - 1:L - 1].flip(1)
    xp = torch.cat([left, x, right], 1)                       # reflect padding
    T = 1 + L // hop
    idx = (torch.arange(T) * hop)[:, None] + torch.arange(win)[None, :]
    fr = xp[:, idx] * (hann(win, x.dtype) if window is None else window)   # [B, T, win]
    k = torch.arange(win, dtype=torch.float64)
    f = torch.arange(win // 2 + 1, dtype=torch.float64)
    ang = 2.0 * math.pi * f[:, None] * k[None, :] / win      # [F, win]
    cr, ci = torch.cos(ang).to(x.dtype), (-torch.sin(ang)).to(x.dtype)
    re = torch.einsum("fk,btk->bft", cr, fr)
    im = torch.einsum("fk,btk->bft", ci, fr)
    return re, im


def istft(re, im, win=512, hop=128, length=None, window=None):
    """torch.istft(spec, n_fft=win, hop_length=hop, window=hann, center=True, length=length), bsrnn.py:382-389:
    frame = irfft(X_t) * w; y = overlap_add(frame) / overlap_add(w^2); drop win/2 samples at the start."""
    B, F, T = re.shape
    k = torch.arange(win, dtype=torch.float64)
    f = torch.arange(F, dtype=torch.float64)
    ang = 2.0 * math.pi * f[:, None] * k[None, :] / win      # [F, win]
    wgt = torch.full((F,), 2.0, dtype=torch.float64)
    wgt[0] = 1.0
    wgt[-1] = 1.0                                             # onesided: DC and Nyquist once, the rest twice
    cr = (wgt[:, None] * torch.cos(ang) / win).to(re.dtype)
    ci = (-wgt[:, None] * torch.sin(ang) / win).to(re.dtype)
    ci[0] = 0.0                                               # irfft ignores the imaginary part of DC / Nyquist
    ci[-1] = 0.0
    w = hann(win, re.dtype) if window is None else window
    fr = (torch.einsum("fk,bft->btk", cr, re) + torch.einsum("fk,bft->btk", ci, im)) * w   # [B, T, win]
    n_out = win + hop * (T - 1)
    y = torch.zeros(B, n_out, dtype=re.dtype)
    env = torch.zeros(n_out, dtype=re.dtype)
    for t in range(T):
        y[:, t * hop:t * hop + win] += fr[:, t]
        env[t * hop:t * hop + win] += w * w
    start = win // 2
    end = n_out - win // 2 if length is None else start + length
    return y[:, start:end] / env[start:end]


def group_norm1(x, w, b, eps=EPS_GN):
    """nn.GroupNorm(1, C, eps): per row over (C, T), biased variance; affine per channel.  x [N, C, T]."""
    mu = x.mean((1, 2), keepdim=True)
    var = ((x - mu) ** 2).mean((1, 2), keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w[None, :, None] + b[None, :, None]


def lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM(batch_first=True): x [N, S, I] -> h [N, S, Hd].  Gate order i, f, g, o
    (rows of w_ih / w_hh in blocks of Hd); c_t = f*c + i*g, h_t = o*tanh(c_t); zero initial state."""
    N, S, _ = x.shape
    Hd = w_hh.shape[1]
    gx = x @ w_ih.t() + (b_ih + b_hh)                         # [N, S, 4Hd]
    h = x.new_zeros(N, Hd)
    c = x.new_zeros(N, Hd)
    out = [None] * S
    for t in (range(S - 1, -1, -1) if reverse else range(S)):
        g = gx[:, t] + h @ w_hh.t()
        i, f, gg, o = g[:, :Hd], g[:, Hd:2 * Hd], g[:, 2 * Hd:3 * Hd], g[:, 3 * Hd:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out, 1)


def res_rnn(x, sd, pre):
    """ResRNN.forward, bsrnn.py:38-46.  x [N, C, S] -> x + proj(blstm(norm(x)^T))^T."""
    y = group_norm1(x, sd[pre + "norm.weight"], sd[pre + "norm.bias"]).transpose(1, 2)      # [N, S, C]
    hf = lstm_dir(y, sd[pre + "rnn.weight_ih_l0"], sd[pre + "rnn.weight_hh_l0"], sd[pre + "rnn.bias_ih_l0"],
                  sd[pre + "rnn.bias_hh_l0"], False)
    hb = lstm_dir(y, sd[pre + "rnn.weight_ih_l0_reverse"], sd[pre + "rnn.weight_hh_l0_reverse"],
                  sd[pre + "rnn.bias_ih_l0_reverse"], sd[pre + "rnn.bias_hh_l0_reverse"], True)
    h = torch.cat([hf, hb], 2)                                                              # [N, S, 2Hd]
    p = h @ sd[pre + "proj.weight"].t() + sd[pre + "proj.bias"]                            # [N, S, C]
    return x + p.transpose(1, 2)


def bsnet(x, sd, pre, nband):
    """BSNet.forward, bsrnn.py:70-83.  x [B, nband*N, T]."""
    B, NN, T = x.shape
    N = NN // nband
    y = res_rnn(x.reshape(B * nband, N, T), sd, pre + "band_rnn.").reshape(B, nband, N, T)
    y = y.permute(0, 3, 2, 1).reshape(B * T, N, nband)
    y = res_rnn(y, sd, pre + "band_comm.").reshape(B, T, N, nband).permute(0, 3, 2, 1)
    return y.reshape(B, NN, T)


def speaker_fuse_4d(x, emb, sd, pre, fuse_type):
    """SpeakerFuseLayer.forward, 4-D branch (wesep/modules/common/speaker.py:88-121).  x [B, nband, N, T],
    emb [B, E]: every variant reduces to a per-row [N] vector (or an [N, N] map for concat) applied at every
    (band, frame)."""
    W, b = sd[pre + "fc.linear.weight"], sd[pre + "fc.linear.bias"]
    if fuse_type == "multiply":
        return x * (emb @ W.t() + b)[:, None, :, None]
    if fuse_type == "additive":
        return x + (emb @ W.t() + b)[:, None, :, None]
    if fuse_type == "concat":                                 # Linear over cat([x, e]) along the feature axis
        N = x.shape[2]
        return torch.einsum("on,bknt->bkot", W[:, :N], x) + (emb @ W[:, N:].t() + b)[:, None, :, None]
    raise ValueError("fuse type not restated: " + fuse_type)


def separator(x, emb, sd, nband, num_repeat, fuse_type, multi_fuse):
    """FuseSeparation.forward, bsrnn.py:125-148.  x [B, nband, N, T] -> same."""
    B, _, N, T = x.shape
    pre = "separator.separation."
    if multi_fuse:
        for r in range(num_repeat):
            x = speaker_fuse_4d(x, emb, sd, f"{pre}{2 * r}.", fuse_type)
            x = bsnet(x.reshape(B, nband * N, T), sd, f"{pre}{2 * r + 1}.", nband).reshape(B, nband, N, T)
        return x
    x = speaker_fuse_4d(x, emb, sd, pre + "0.", fuse_type).reshape(B, nband * N, T)
    for r in range(num_repeat):
        x = bsnet(x, sd, f"{pre}{r + 1}.", nband)
    return x.reshape(B, nband, N, T)


def bsrnn_forward(sd, mix, emb, sr=16000, win=512, stride=128, num_repeat=6, spk_fuse_type="multiply",
                  multi_fuse=False):
    """BSRNN.forward with joint_training=False, use_spk_transform=False (bsrnn.py:300-394).
    mix [B, L], emb [B, E] -> estimate [B, L]."""
    B, L = mix.shape
    bands = band_widths(sr, win)
    nband = len(bands)
    re, im = stft(mix, win, stride)                           # [B, F, T]
    feats, lo = [], 0
    for i, bw in enumerate(bands):                            # band split: GroupNorm(1, 2bw) + Conv1d(2bw, N, 1)
        sub = torch.cat([re[:, lo:lo + bw], im[:, lo:lo + bw]], 1)                         # [B, 2bw, T]
        y = group_norm1(sub, sd[f"BN.{i}.0.weight"], sd[f"BN.{i}.0.bias"])
        feats.append(torch.einsum("oc,bct->bot", sd[f"BN.{i}.1.weight"][:, :, 0], y) + sd[f"BN.{i}.1.bias"][None, :, None])
        lo += bw
    x = torch.stack(feats, 1)                                 # [B, nband, N, T]
    x = separator(x, emb, sd, nband, num_repeat, spk_fuse_type, multi_fuse)
    est_re, est_im, lo = [], [], 0
    for i, bw in enumerate(bands):                            # mask head, bsrnn.py:365-381
        y = group_norm1(x[:, i], sd[f"mask.{i}.0.weight"], sd[f"mask.{i}.0.bias"])
        y = torch.tanh(torch.einsum("oc,bct->bot", sd[f"mask.{i}.1.weight"][:, :, 0], y) + sd[f"mask.{i}.1.bias"][None, :, None])
        y = torch.tanh(torch.einsum("oc,bct->bot", sd[f"mask.{i}.3.weight"][:, :, 0], y) + sd[f"mask.{i}.3.bias"][None, :, None])
        y = torch.einsum("oc,bct->bot", sd[f"mask.{i}.5.weight"][:, :, 0], y) + sd[f"mask.{i}.5.bias"][None, :, None]
        y = y.reshape(B, 2, 2, bw, -1)
        m = y[:, 0] * torch.sigmoid(y[:, 1])                  # [B, 2(re/im), bw, T]
        mr, mi = m[:, 0], m[:, 1]
        sr_, si_ = re[:, lo:lo + bw], im[:, lo:lo + bw]
        est_re.append(sr_ * mr - si_ * mi)
        est_im.append(sr_ * mi + si_ * mr)
        lo += bw
    return istft(torch.cat(est_re, 1), torch.cat(est_im, 1), win, stride, length=L)
