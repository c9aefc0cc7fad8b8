"""ORACLE (test infrastructure) — CPU restatement of the pDPCCN forward (SURVEY.md §8 row a23;
wesep/models/dpccn.py:206-290, wesep/modules/dpccn/convs.py:28-152).

Plain torch functional ops in the dtype of the inputs (fp32 or fp64), parameters taken from a ``state_dict`` with the
reference's key names.  STFT / iSTFT are the explicit DFT sums of oracle/bsrnn.py (pinned there against torch.stft);
InstanceNorm, ELU, the pooling tail and the speaker gain are written out.  Backward = torch autograd of this forward.
Pinned: tests/golden/dpccn_*.npz hold outputs / loss / gradient summaries of the REAL reference module
(tests/golden/make_golden_dpccn.py) and tests/test_oracle_dpccn.py compares this file with them.

Scope: ``joint_training=False`` (a given 256-d speaker embedding; the wespeaker encoder is an external package),
``use_spk_transform=False``, ``spk_fuse_type="multiply"``, non-causal TCN.
"""
import torch
import torch.nn.functional as F

from oracle import bsrnn as ob


def inorm(x, eps=1e-5):
    """nn.InstanceNorm{1,2}d (no affine, no running stats): per (n, c) plane, biased variance."""
    dims = tuple(range(2, x.dim()))
    mu = x.mean(dims, keepdim=True)
    var = ((x - mu) ** 2).mean(dims, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def elu(x):
    return torch.where(x > 0, x, torch.expm1(x))


def conv2d_block(x, sd, pre, stride):
    """convs.py:28-47."""
    return inorm(elu(F.conv2d(x, sd[pre + "conv2d.weight"], sd[pre + "conv2d.bias"], stride=stride, padding=(1, 1))))


def convtrans2d_block(x, sd, pre, stride):
    """convs.py:50-69."""
    y = F.conv_transpose2d(x, sd[pre + "convtrans2d.weight"], sd[pre + "convtrans2d.bias"], stride=stride, padding=(1, 1))
    return inorm(elu(y))


def dense_block(x, sd, pre):
    """convs.py:99-106."""
    feats = [x]
    for i in range(1, 6):
        feats.append(conv2d_block(torch.cat(feats, 1), sd, pre + f"conv{i}.", (1, 1)))
    return feats[-1]


def tcn_block(x, sd, pre, dilation):
    """convs.py:142-152 (non-causal)."""
    y = elu(inorm(x))
    y = F.conv1d(y, sd[pre + "dconv1.weight"], sd[pre + "dconv1.bias"], padding=dilation, dilation=dilation, groups=x.shape[1])
    y = elu(inorm(y))
    y = F.conv1d(y, sd[pre + "dconv2.weight"], sd[pre + "dconv2.bias"])
    return x + y


def dpccn_forward(sd, mix, emb, win=512, stride=128, tcn_blocks=10, tcn_layers=2, pool_size=(4, 8, 16, 32)):
    """dpccn.py:206-290 with joint_training=False: returns the estimate [B, L]."""
    B, L = mix.shape
    re, im = ob.stft(mix, win, stride)                                   # [B, F, T]
    spec = torch.stack([re, im], 1).transpose(2, 3)                      # [B, 2, T, F]
    out = F.conv2d(spec, sd["conv2d.weight"], sd["conv2d.bias"], stride=(1, 1), padding=(1, 1))
    out = dense_block(out, sd, "encoder.0.")
    # 4-D multiply fusion (speaker.py:117-121): out.transpose(2, 3) * fc(embed) broadcast over channels and frames
    gain = F.linear(emb, sd["spk_fuse.fc.linear.weight"], sd["spk_fuse.fc.linear.bias"])     # [B, F]
    out = out * gain[:, None, None, :]
    outs = [out]
    for i in range(1, 5):
        out = conv2d_block(out, sd, f"encoder.{i}.0.", (1, 2))
        out = dense_block(out, sd, f"encoder.{i}.1.")
        outs.append(out)
    for i in range(5, 8):
        out = conv2d_block(out, sd, f"encoder.{i}.", (1, 2))
        outs.append(out)
    Bn, N, T, Fq = out.shape
    out = out.reshape(Bn, N, T * Fq)
    for li in range(tcn_layers):
        for b in range(tcn_blocks):
            out = tcn_block(out, sd, f"tcn_layers.{li}.{b}.", 2 ** b)
    out = out.reshape(Bn, N, T, Fq)
    outs = outs[::-1]
    for idx in range(8):
        x = torch.cat([outs[idx], out], 1)
        if idx < 3:
            out = convtrans2d_block(x, sd, f"decoder.{idx}.", (1, 2))
        elif idx < 7:
            out = dense_block(x, sd, f"decoder.{idx}.0.")
            out = convtrans2d_block(out, sd, f"decoder.{idx}.1.", (1, 2))
        else:
            out = dense_block(x, sd, "decoder.7.")
    Bn, N, T, Fq = out.shape
    pools = [out]
    for j, sz in enumerate(pool_size):
        p = F.avg_pool2d(out, sz)
        p = F.conv2d(p, sd[f"avg_pool.{j}.1.weight"], sd[f"avg_pool.{j}.1.bias"])
        pools.append(F.interpolate(p, size=(T, Fq), mode="bilinear", align_corners=False))
    out = F.conv2d(torch.cat(pools, 1), sd["avg_proj.weight"], sd["avg_proj.bias"])
    out = F.conv_transpose2d(out, sd["deconv2d.weight"], sd["deconv2d.bias"], stride=(1, 1), padding=(1, 1))
    est = out.transpose(2, 3)                                            # [B, 2, F, T]
    return ob.istft(est[:, 0], est[:, 1], win, stride, length=L)


def make_state_dict(tcn_blocks=10, tcn_layers=2, spk_emb_dim=256, feature_dim=257, tcn_dims=384, pool_size=(4, 8, 16, 32),
                    dtype=torch.float32):
    """Keys / shapes / ORDER of DPCCN(joint_training=False).state_dict() from the constructor rules of
    wesep/models/dpccn.py:58-204 and wesep/modules/dpccn/convs.py (uninitialised tensors; fill with synth)."""
    sd = {}

    def conv(pre, co, ci, k=(3, 3)):
        sd[pre + "weight"] = torch.empty(co, ci, *k, dtype=dtype)
        sd[pre + "bias"] = torch.empty(co, dtype=dtype)

    def convt(pre, ci, co):
        sd[pre + "weight"] = torch.empty(ci, co, 3, 3, dtype=dtype)
        sd[pre + "bias"] = torch.empty(co, dtype=dtype)

    def dense(pre, c, co, mode):
        n = 1 if mode == "enc" else 2
        for i in range(1, 5):
            conv(pre + f"conv{i}.conv2d.", c, c * (n + i - 1))
        conv(pre + "conv5.conv2d.", co, c * (n + 4))

    conv("conv2d.", 16, 2)
    dense("encoder.0.", 16, 16, "enc")
    for i in range(4):
        conv(f"encoder.{i + 1}.0.conv2d.", 32, 16 if i == 0 else 32)
        dense(f"encoder.{i + 1}.1.", 32, 32, "enc")
    conv("encoder.5.conv2d.", 64, 32)
    conv("encoder.6.conv2d.", 128, 64)
    conv("encoder.7.conv2d.", 384, 128)
    sd["spk_fuse.fc.linear.weight"] = torch.empty(feature_dim, spk_emb_dim, dtype=dtype)
    sd["spk_fuse.fc.linear.bias"] = torch.empty(feature_dim, dtype=dtype)
    for li in range(tcn_layers):
        for b in range(tcn_blocks):
            pre = f"tcn_layers.{li}.{b}."
            sd[pre + "dconv1.weight"] = torch.empty(tcn_dims, 1, 3, dtype=dtype)
            sd[pre + "dconv1.bias"] = torch.empty(tcn_dims, dtype=dtype)
            sd[pre + "dconv2.weight"] = torch.empty(tcn_dims, tcn_dims, 1, dtype=dtype)
            sd[pre + "dconv2.bias"] = torch.empty(tcn_dims, dtype=dtype)
    convt("decoder.0.convtrans2d.", 768, 128)
    convt("decoder.1.convtrans2d.", 256, 64)
    convt("decoder.2.convtrans2d.", 128, 32)
    for i in range(4):
        dense(f"decoder.{3 + i}.0.", 32, 64, "dec")
        convt(f"decoder.{3 + i}.1.convtrans2d.", 64, 32 if i != 3 else 16)
    dense("decoder.7.", 16, 32, "dec")
    for j in range(len(pool_size)):
        conv(f"avg_pool.{j}.1.", 8, 32, (1, 1))
    conv("avg_proj.", 32, 64, (1, 1))
    convt("deconv2d.", 32, 2)
    return sd
