"""ORACLE (test infrastructure) — CPU restatement of the TF-GridNet forward (SURVEY.md §8 row a24;
wesep/models/tfgridnet.py:197-302, wesep/modules/tfgridnet/gridnet_block.py:118-284).

Plain torch ops in the dtype of the inputs, parameters from a ``state_dict`` with the reference's key names; the
recurrences, the STFT / iSTFT, the layer norms and the attention are written out (oracle/bsrnn.py's explicit LSTM and
DFT sums).  Backward = torch autograd of this forward.  Pinned: tests/golden/tfgridnet_*.npz hold outputs / loss /
gradient summaries of the REAL reference module (tests/golden/make_golden_tfgridnet.py); tests/test_oracle_tfgridnet.py
compares this file with them.

Scope: ``joint_training=False`` (a given speaker embedding), ``emb_ks == emb_hs == 1`` (tfgridnet.yaml:52-53) or
``emb_ks != emb_hs`` (the class default 4 / 1) or ``emb_ks == emb_hs > 1``, one source, one microphone, ``multiply`` fusion.
"""
import math

import torch
import torch.nn.functional as F

from oracle import bsrnn as ob


def layer_norm_c(x, w, b, eps):
    """nn.LayerNorm(C, eps) over the last dimension."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def blstm(x, sd, pre):
    """nn.LSTM(batch_first=True, bidirectional=True): x [N, S, I] -> [N, S, 2 Hd]."""
    hf = ob.lstm_dir(x, sd[pre + "weight_ih_l0"], sd[pre + "weight_hh_l0"], sd[pre + "bias_ih_l0"], sd[pre + "bias_hh_l0"], False)
    hb = ob.lstm_dir(x, sd[pre + "weight_ih_l0_reverse"], sd[pre + "weight_hh_l0_reverse"], sd[pre + "bias_ih_l0_reverse"],
                     sd[pre + "bias_hh_l0_reverse"], True)
    return torch.cat([hf, hb], 2)


def prelu(x, a, dim):
    shape = [1] * x.dim()
    shape[dim] = a.numel()
    return torch.where(x >= 0, x, a.view(shape) * x)


def all_head_prelu_ln(x, sd, pre, H, E, eps):
    """gridnet_block.py:271-284: x [B, H*E, T, F] -> [B, H, E, T, F]."""
    B, _, T, Fq = x.shape
    x = x.view(B, H, E, T, Fq)
    x = prelu(x, sd[pre + "act.weight"], 1)
    mu = x.mean(dim=(2, 4), keepdim=True)
    std = torch.sqrt(((x - mu) ** 2).mean(dim=(2, 4), keepdim=True) + eps)
    return (x - mu) / std * sd[pre + "gamma"] + sd[pre + "beta"]


def ln_4dcf(x, sd, pre, eps):
    """gridnet_block.py:241-252: x [B, C, T, F], statistics over (C, F)."""
    mu = x.mean(dim=(1, 3), keepdim=True)
    std = torch.sqrt(((x - mu) ** 2).mean(dim=(1, 3), keepdim=True) + eps)
    return (x - mu) / std * sd[pre + "gamma"] + sd[pre + "beta"]


def _rnn_path(v, sd, pre, name, ks, hs, eps):
    """One of the two recurrent paths of a block on v [B, A, S, C] (sequences along S): LayerNorm(C) -> (unfold) -> BLSTM ->
    Linear or ConvTranspose1d -> + v (gridnet_block.py:135-161 / 163-187)."""
    B, A, S, C = v.shape
    y = layer_norm_c(v, sd[pre + name + "_norm.weight"], sd[pre + name + "_norm.bias"], eps).reshape(B * A, S, C)
    if ks == hs:                                                       # ks consecutive positions per step, features (k, c)
        y = y.reshape(B * A, S // ks, ks * C)
        y = blstm(y, sd, pre + name + "_rnn.") @ sd[pre + name + "_linear.weight"].t() + sd[pre + name + "_linear.bias"]
        return y.reshape(B, A, S, C) + v
    y = F.unfold(y.transpose(1, 2)[..., None], (ks, 1), stride=(hs, 1)).transpose(1, 2)       # [BA, L, C*ks]
    y = blstm(y, sd, pre + name + "_rnn.").transpose(1, 2)                                    # [BA, 2H, L]
    y = F.conv_transpose1d(y, sd[pre + name + "_linear.weight"], sd[pre + name + "_linear.bias"], stride=hs)   # [BA, C, S]
    return y.view(B, A, C, S).transpose(-2, -1) + v


def gridnet_block(x, sd, pre, n_head, eps, ks=1, hs=1):
    """gridnet_block.py:118-227.  x [B, C, T, Q]."""
    B, C, old_T, old_Q = x.shape
    olp = ks - hs
    T = math.ceil((old_T + 2 * olp - ks) / hs) * hs + ks
    Q = math.ceil((old_Q + 2 * olp - ks) / hs) * hs + ks
    x = x.permute(0, 2, 3, 1)                                            # [B, T, Q, C]
    x = F.pad(x, (0, 0, olp, Q - old_Q - olp, olp, T - old_T - olp))
    y = _rnn_path(x, sd, pre, "intra", ks, hs, eps)                      # along Q
    y = y.transpose(1, 2)                                                # [B, Q, T, C]
    z = _rnn_path(y, sd, pre, "inter", ks, hs, eps)                      # along T
    batch = z.permute(0, 3, 2, 1)[..., olp:olp + old_T, olp:olp + old_Q]  # [B, C, T, Q]
    T, Q = old_T, old_Q
    E = sd[pre + "attn_conv_Q.weight"].shape[0] // n_head
    Ev = C // n_head
    Qh = all_head_prelu_ln(F.conv2d(batch, sd[pre + "attn_conv_Q.weight"], sd[pre + "attn_conv_Q.bias"]), sd, pre + "attn_norm_Q.",
                           n_head, E, eps)
    Kh = all_head_prelu_ln(F.conv2d(batch, sd[pre + "attn_conv_K.weight"], sd[pre + "attn_conv_K.bias"]), sd, pre + "attn_norm_K.",
                           n_head, E, eps)
    Vh = all_head_prelu_ln(F.conv2d(batch, sd[pre + "attn_conv_V.weight"], sd[pre + "attn_conv_V.bias"]), sd, pre + "attn_norm_V.",
                           n_head, Ev, eps)
    Qm = Qh.reshape(B * n_head, E, T, Q).transpose(1, 2).flatten(2)                       # [B', T, E*Q]
    Km = Kh.reshape(B * n_head, E, T, Q).transpose(2, 3).reshape(B * n_head, E * Q, T)    # [B', E*Q, T]
    Vm = Vh.reshape(B * n_head, Ev, T, Q).transpose(1, 2)                                 # [B', T, Ev, Q]
    shape = Vm.shape
    att = torch.softmax(Qm @ Km / (Qm.shape[-1] ** 0.5), dim=2) @ Vm.flatten(2)           # [B', T, Ev*Q]
    att = att.reshape(shape).transpose(1, 2).reshape(B, C, T, Q)
    p = F.conv2d(att, sd[pre + "attn_concat_proj.0.weight"], sd[pre + "attn_concat_proj.0.bias"])
    p = prelu(p, sd[pre + "attn_concat_proj.1.weight"], 1)
    p = ln_4dcf(p, sd, pre + "attn_concat_proj.2.", eps)
    return p + batch


def tfgridnet_forward(sd, mix, emb, n_fft=128, stride=64, n_layers=6, n_head=4, eps=1e-5, emb_ks=1, emb_hs=1):
    """tfgridnet.py:197-302 with joint_training=False: returns the estimate [B, L]."""
    B, L = mix.shape
    std = torch.std(mix, dim=1, keepdim=True)
    x = mix / std
    window = torch.hann_window(n_fft, dtype=x.dtype)                      # built in the input dtype, tfgridnet.py:224-227
    re, im = ob.stft(x, n_fft, stride, window=window)                     # [B, F, T]
    batch = torch.stack([re.transpose(1, 2), im.transpose(1, 2)], 1)      # [B, 2, T, F]
    batch = F.conv2d(batch, sd["conv.0.weight"], sd["conv.0.bias"], padding=(1, 1))
    mu = batch.mean(dim=(1, 2, 3), keepdim=True)
    var = ((batch - mu) ** 2).mean(dim=(1, 2, 3), keepdim=True)
    batch = (batch - mu) / torch.sqrt(var + eps) * sd["conv.1.weight"].view(1, -1, 1, 1) + sd["conv.1.bias"].view(1, -1, 1, 1)
    gain = F.linear(emb, sd["spk_fuse.fc.linear.weight"], sd["spk_fuse.fc.linear.bias"])   # [B, F]
    for i in range(n_layers):
        batch = batch * gain[:, None, None, :]                            # speaker.py:117-121 (4-D multiply)
        batch = gridnet_block(batch, sd, f"blocks.{i}.", n_head, eps, emb_ks, emb_hs)
    out = F.conv_transpose2d(batch, sd["deconv.weight"], sd["deconv.bias"], padding=(1, 1))   # [B, 2, T, F]
    est = ob.istft(out[:, 0].transpose(1, 2), out[:, 1].transpose(1, 2), n_fft, stride, length=L, window=window)
    return est * std


def make_state_dict(n_layers=6, emb_dim=128, hidden=192, n_head=4, approx_qk_dim=512, n_fft=128, spk_emb_dim=256,
                    dtype=torch.float32, emb_ks=1, emb_hs=1):
    """Keys / shapes / ORDER of TFGridNet(joint_training=False, emb_ks=1, emb_hs=1).state_dict()
    (tfgridnet.py:169-195, gridnet_block.py:45-111)."""
    Fq = n_fft // 2 + 1
    E = math.ceil(approx_qk_dim * 1.0 / Fq)
    sd = {}

    def t(*shape):
        return torch.empty(*shape, dtype=dtype)

    sd["spk_fuse.fc.linear.weight"] = t(Fq, spk_emb_dim)
    sd["spk_fuse.fc.linear.bias"] = t(Fq)
    sd["conv.0.weight"] = t(emb_dim, 2, 3, 3)
    sd["conv.0.bias"] = t(emb_dim)
    sd["conv.1.weight"] = t(emb_dim)
    sd["conv.1.bias"] = t(emb_dim)
    for i in range(n_layers):
        pre = f"blocks.{i}."
        for path in ("intra", "inter"):
            sd[pre + path + "_norm.weight"] = t(emb_dim)
            sd[pre + path + "_norm.bias"] = t(emb_dim)
            for suf in ("", "_reverse"):
                sd[pre + path + "_rnn.weight_ih_l0" + suf] = t(4 * hidden, emb_dim * emb_ks)
                sd[pre + path + "_rnn.weight_hh_l0" + suf] = t(4 * hidden, hidden)
                sd[pre + path + "_rnn.bias_ih_l0" + suf] = t(4 * hidden)
                sd[pre + path + "_rnn.bias_hh_l0" + suf] = t(4 * hidden)
            sd[pre + path + "_linear.weight"] = t(emb_dim * emb_ks, 2 * hidden) if emb_ks == emb_hs else t(2 * hidden, emb_dim, emb_ks)
            sd[pre + path + "_linear.bias"] = t(emb_dim * emb_ks if emb_ks == emb_hs else emb_dim)
        for name, co, e in (("Q", n_head * E, E), ("K", n_head * E, E), ("V", emb_dim, emb_dim // n_head)):
            sd[pre + f"attn_conv_{name}.weight"] = t(co, emb_dim, 1, 1)
            sd[pre + f"attn_conv_{name}.bias"] = t(co)
            sd[pre + f"attn_norm_{name}.gamma"] = t(1, n_head, e, 1, Fq)
            sd[pre + f"attn_norm_{name}.beta"] = t(1, n_head, e, 1, Fq)
            sd[pre + f"attn_norm_{name}.act.weight"] = t(n_head)
        sd[pre + "attn_concat_proj.0.weight"] = t(emb_dim, emb_dim, 1, 1)
        sd[pre + "attn_concat_proj.0.bias"] = t(emb_dim)
        sd[pre + "attn_concat_proj.1.weight"] = t(1)
        sd[pre + "attn_concat_proj.2.gamma"] = t(1, emb_dim, 1, Fq)
        sd[pre + "attn_concat_proj.2.beta"] = t(1, emb_dim, 1, Fq)
    sd["deconv.weight"] = t(emb_dim, 2, 3, 3)
    sd["deconv.bias"] = t(2)
    return sd
