"""ORACLE (test infrastructure) — functional restatement of the reference Spex+ /
ConvTasNet forward (wesep/models/convtasnet.py:162-219) in plain torch ops.

All functions take the reference ``state_dict`` (same keys / shapes)
as a flat ``dict[str, Tensor]`` plus a key prefix, so no reference class is needed on the
GPU box.  dtype / device follow the tensors passed in (fp32 or fp64).  Backward = torch
autograd through these ops, which is what the reference itself runs.

Supported configuration = what the recipes construct (spexplus.yaml:36-56): Multi
encoder/decoder, gLN, non-causal, skip_con False, multi_fuse True, joint_training True,
spk_feat False; spk_fuse_type in {concatConv, concat, additive, multiply, FiLM}.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- norms
def cln(x, w, b, eps=1e-5):
    """ChannelWiseLayerNorm — wesep/modules/common/norm.py:51-66 (LayerNorm over C per frame)."""
    xt = torch.transpose(x, 1, 2)
    xt = F.layer_norm(xt, (x.shape[1],), w, b, eps)
    return torch.transpose(xt, 1, 2)


def gln(x, w, b, eps=1e-5):
    """GlobalChannelLayerNorm — wesep/modules/common/norm.py:29-48 (biased var over (C,T))."""
    mean = torch.mean(x, (1, 2), keepdim=True)
    var = torch.mean((x - mean) ** 2, (1, 2), keepdim=True)
    return w * (x - mean) / torch.sqrt(var + eps) + b


def prelu(x, a):
    return F.prelu(x, a)


# --------------------------------------------------------------------------- encoder
def multi_encoder(sd, pre, x, L1=20, stride=10, L2=80, L3=160, only_w=False, relu=F.relu):
    """MultiEncoder.forward — wesep/modules/tasnet/encoder.py:95-114.  (`relu` overridable by tests to pin the
    branch at |x| ~ 0, like `prelu` in conv1d_block.)"""
    if x.dim() == 2:
        x = x.unsqueeze(1)  # Conv1D.forward, convs.py:19
    w1 = relu(F.conv1d(x, sd[pre + "encoder_1d_short.weight"], sd[pre + "encoder_1d_short.bias"], stride=stride))
    T = w1.shape[-1]
    xlen1 = x.shape[-1]
    xlen2 = (T - 1) * stride + L2
    xlen3 = (T - 1) * stride + L3
    w2 = relu(F.conv1d(F.pad(x, (0, xlen2 - xlen1)), sd[pre + "encoder_1d_middle.weight"],
                       sd[pre + "encoder_1d_middle.bias"], stride=stride))
    w3 = relu(F.conv1d(F.pad(x, (0, xlen3 - xlen1)), sd[pre + "encoder_1d_long.weight"],
                       sd[pre + "encoder_1d_long.bias"], stride=stride))
    if only_w:
        return None, w1, w2, w3
    e = cln(torch.cat([w1, w2, w3], 1), sd[pre + "ln.weight"], sd[pre + "ln.bias"])
    e = F.conv1d(e, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    return e, w1, w2, w3


# --------------------------------------------------------------------------- TCN blocks
def conv1d_block(sd, pre, x, dilation, P=3, prelu=prelu):
    """Conv1DBlock.forward (skip_con False, non-causal) — wesep/modules/tasnet/convs.py:84-104.
    `prelu` may be overridden by tests to pin the branch taken at |x| ~ 0 (PReLU's derivative is
    discontinuous there, so fp32 and fp64 runs legitimately disagree on isolated elements)."""
    H = sd[pre + "dwconv.weight"].shape[0]
    c = F.conv1d(x, sd[pre + "conv1x1.weight"], sd[pre + "conv1x1.bias"])
    c = prelu(c, sd[pre + "PReLU_1.weight"])
    c = gln(c, sd[pre + "norm_1.weight"], sd[pre + "norm_1.bias"])
    pad = (dilation * (P - 1)) // 2
    c = F.conv1d(c, sd[pre + "dwconv.weight"], sd[pre + "dwconv.bias"], padding=pad, dilation=dilation, groups=H)
    c = prelu(c, sd[pre + "PReLU_2.weight"])
    c = gln(c, sd[pre + "norm_2.weight"], sd[pre + "norm_2.bias"])
    c = F.conv1d(c, sd[pre + "Output.weight"], sd[pre + "Output.bias"])
    return x + c


def conv1d_block4fuse(sd, pre, x, aux, dilation=1, P=3, prelu=prelu):
    """Conv1DBlock4Fuse.forward — wesep/modules/tasnet/convs.py:148-160 (aux [n,E,1])."""
    H = sd[pre + "dconv.weight"].shape[0]
    T = x.shape[-1]
    y = torch.cat([x, aux.repeat(1, 1, T)], 1)
    y = F.conv1d(y, sd[pre + "conv1x1.weight"], sd[pre + "conv1x1.bias"])
    y = gln(prelu(y, sd[pre + "prelu1.weight"]), sd[pre + "lnorm1.weight"], sd[pre + "lnorm1.bias"])
    pad = (dilation * (P - 1)) // 2
    y = F.conv1d(y, sd[pre + "dconv.weight"], sd[pre + "dconv.bias"], padding=pad, dilation=dilation, groups=H)
    y = gln(prelu(y, sd[pre + "prelu2.weight"]), sd[pre + "lnorm2.weight"], sd[pre + "lnorm2.bias"])
    y = F.conv1d(y, sd[pre + "sconv.weight"], sd[pre + "sconv.bias"])
    return x + y


def speaker_fuse_layer(sd, pre, x, embed, fuse_type):
    """SpeakerFuseLayer.forward, 3-D branches — wesep/modules/common/speaker.py:81-125;
    FiLM — wesep/modules/common/norm.py:118-139."""
    if fuse_type == "concat":
        embed_t = embed.expand(-1, -1, x.size(2))
        y = torch.transpose(torch.cat([x, embed_t], 1), 1, 2)
        return torch.transpose(F.linear(y, sd[pre + "fc.linear.weight"], sd[pre + "fc.linear.bias"]), 1, 2)
    if fuse_type in ("additive", "multiply"):
        embed_t = torch.transpose(embed.expand(-1, -1, x.size(2)), 1, 2)
        v = torch.transpose(F.linear(embed_t, sd[pre + "fc.linear.weight"], sd[pre + "fc.linear.bias"]), 1, 2)
        return x + v if fuse_type == "additive" else x * v
    if fuse_type == "FiLM":
        e = embed.squeeze(-1)
        gamma = F.linear(e, sd[pre + "fc.gamma_fcs.0.weight"], sd[pre + "fc.gamma_fcs.0.bias"])
        beta = F.linear(e, sd[pre + "fc.beta_fcs.0.weight"], sd[pre + "fc.beta_fcs.0.bias"])
        return (1 + gamma.unsqueeze(-1)) * x + beta.unsqueeze(-1)
    raise ValueError("Fuse type not defined.")


def fuse_separation(sd, pre, x, spk, cfg):
    """FuseSeparation.forward (multi_fuse True) — wesep/modules/tasnet/separation.py:166-186;
    construction order :93-135; inner Separation(1, X, start_dilation) :34-37,53-56."""
    R, X, P = cfg["R"], cfg["X"], cfg["P"]
    ftype = cfg.get("spk_fuse_type", "concatConv")
    for r in range(R):
        if ftype == "concatConv":
            x = conv1d_block4fuse(sd, f"{pre}separation.{2 * r}.", x, spk, 1, P)
            for j, xx in enumerate(range(1, X)):
                x = conv1d_block(sd, f"{pre}separation.{2 * r + 1}.separation.{j}.", x, 2 ** xx, P)
        else:
            x = speaker_fuse_layer(sd, f"{pre}separation.{4 * r}.", x, spk, ftype)
            # NB reference quirk: indices 4r+1 (PReLU) and 4r+2 (norm) are called with x only
            x = prelu(x, sd[f"{pre}separation.{4 * r + 1}.weight"])
            x = gln(x, sd[f"{pre}separation.{4 * r + 2}.weight"], sd[f"{pre}separation.{4 * r + 2}.bias"])
            for j, xx in enumerate(range(0, X)):
                x = conv1d_block(sd, f"{pre}separation.{4 * r + 3}.separation.{j}.", x, 2 ** xx, P)
    return x


# --------------------------------------------------------------------------- speaker encoder
def resblock(sd, pre, x, training, buffers_out=None, momentum=0.1, eps=1e-5):
    """ResBlock.forward — wesep/modules/tasnet/speaker.py:31-45 (BatchNorm1d batch stats in train)."""
    def bn(v, name):
        rm, rv = sd[pre + name + ".running_mean"], sd[pre + name + ".running_var"]
        if training:
            rm, rv = rm.clone(), rv.clone()
        out = F.batch_norm(v, rm, rv, sd[pre + name + ".weight"], sd[pre + name + ".bias"],
                           training=training, momentum=momentum, eps=eps)
        if training and buffers_out is not None:
            buffers_out[pre + name + ".running_mean"] = rm
            buffers_out[pre + name + ".running_var"] = rv
        return out

    residual = x
    y = F.conv1d(x, sd[pre + "conv1.weight"])
    y = bn(y, "batch_norm1")
    y = prelu(y, sd[pre + "prelu1.weight"])
    y = F.conv1d(y, sd[pre + "conv2.weight"])
    y = bn(y, "batch_norm2")
    if (pre + "conv_downsample.weight") in sd:
        residual = F.conv1d(residual, sd[pre + "conv_downsample.weight"])
    y = y + residual
    y = prelu(y, sd[pre + "prelu2.weight"])
    return F.max_pool1d(y, 3)


def resnet4spexplus(sd, pre, x, training, buffers_out=None):
    """ResNet4SpExplus.forward — wesep/modules/tasnet/speaker.py:61-64."""
    p = pre + "aux_enc3."
    y = cln(x, sd[p + "0.weight"], sd[p + "0.bias"])
    y = F.conv1d(y, sd[p + "1.weight"], sd[p + "1.bias"])
    for i in (2, 3, 4):
        y = resblock(sd, f"{p}{i}.", y, training, buffers_out)
    y = F.conv1d(y, sd[p + "5.weight"], sd[p + "5.bias"])
    return y.mean(dim=-1)


# --------------------------------------------------------------------------- decoder
def multi_decoder(sd, pre, e, w1, w2, w3, stride=10, relu_on=None):
    """MultiDecoder.forward (actLayer = ReLU) — wesep/modules/tasnet/decoder.py:92-114.
    `relu_on` (tests only): three boolean masks that PIN the ReLU branch of every element (two fp32 implementations
    legitimately disagree where a pre-activation is ~1e-7; a flipped branch is a discrete change of the gradient)."""
    ests = []
    xlen = None
    for i, w in enumerate((w1, w2, w3), start=1):
        a = F.conv1d(e, sd[f"{pre}mask{i}.weight"], sd[f"{pre}mask{i}.bias"])
        m = F.relu(a) if relu_on is None else a * relu_on[i - 1].to(a.dtype)
        s = w * m
        est = F.conv_transpose1d(s, sd[f"{pre}decoder_1d_{i}.weight"], sd[f"{pre}decoder_1d_{i}.bias"], stride=stride)
        est = est.squeeze(1)  # torch.squeeze of the size-1 channel dim; n==1 handled at decoder.py:109-112
        if xlen is None:
            xlen = est.shape[-1]
        ests.append(est[:, :xlen])
    return ests


# --------------------------------------------------------------------------- model
DEFAULT_CFG = dict(N=256, L=20, B=256, H=512, P=3, X=8, R=4, spk_emb_dim=256, spk_fuse_type="concatConv",
                   multi_task=True, spksInTrain=251)


def convtasnet_forward(sd, cfg, x, enroll, training=True, buffers_out=None):
    """ConvTasNet.forward — wesep/models/convtasnet.py:162-219 (recipe configuration).

    Returns [est1, est2, est3, (speaker logits if multi_task)].
    """
    if x.dim() >= 3:
        raise RuntimeError("ConvTasNet accept 1/2D tensor as input, but got {:d}".format(x.dim()))
    if x.dim() == 1:
        x = x.unsqueeze(0)
    L = cfg["L"]
    e, w1, w2, w3 = multi_encoder(sd, "encoder.", x, L1=L, stride=L // 2)               # :171
    _, a1, a2, a3 = multi_encoder(sd, "encoder.", enroll, L1=L, stride=L // 2, only_w=True)  # :184 (ln/proj discarded)
    emb = resnet4spexplus(sd, "spk_model.", torch.cat([a1, a2, a3], 1), training, buffers_out)  # :185-190
    logits = None
    if cfg.get("multi_task", True):
        logits = F.linear(emb, sd["pred_linear.weight"], sd["pred_linear.bias"])       # :194
    spk = emb.unsqueeze(-1)                                                             # :196 (Identity transform)
    e = fuse_separation(sd, "separation.", e, spk, cfg)                                 # :197
    s = multi_decoder(sd, "decoder.", e, w1, w2, w3, stride=L // 2)                     # :201
    if logits is not None:
        s.append(logits)                                                                # :213-217
    return s


# --------------------------------------------------------------------------- state_dict contract
def state_dict_spec(cfg):
    """Ordered [(key, shape)] of the reference ConvTasNet.state_dict() for `cfg`
    (construction order of convtasnet.py:65-152).  Lets the GPU box
    rebuild seeded weights without the reference classes; checked against the real
    reference in tests/test_oracle_vs_reference.py."""
    N, L, B, H, P, X, R = (cfg[k] for k in ("N", "L", "B", "H", "P", "X", "R"))
    E = cfg.get("spk_emb_dim", 256)
    ftype = cfg.get("spk_fuse_type", "concatConv")
    spec = []

    def add(k, *shape):
        spec.append((k, tuple(shape)))

    for nm, k in (("short", L), ("middle", 80), ("long", 160)):
        add(f"encoder.encoder_1d_{nm}.weight", N, 1, k)
        add(f"encoder.encoder_1d_{nm}.bias", N)
    add("encoder.ln.weight", 3 * N)
    add("encoder.ln.bias", 3 * N)
    add("encoder.proj.weight", B, 3 * N, 1)
    add("encoder.proj.bias", B)
    p = "spk_model.aux_enc3."
    add(p + "0.weight", 3 * N)
    add(p + "0.bias", 3 * N)
    add(p + "1.weight", 256, 3 * 256, 1)
    add(p + "1.bias", 256)
    for i, (ci, co) in zip((2, 3, 4), ((256, 256), (256, 512), (512, 512))):
        q = f"{p}{i}."
        add(q + "conv1.weight", co, ci, 1)
        add(q + "conv2.weight", co, co, 1)
        for bn in ("batch_norm1", "batch_norm2"):
            add(f"{q}{bn}.weight", co)
            add(f"{q}{bn}.bias", co)
            add(f"{q}{bn}.running_mean", co)
            add(f"{q}{bn}.running_var", co)
            add(f"{q}{bn}.num_batches_tracked")
        add(q + "prelu1.weight", 1)
        add(q + "prelu2.weight", 1)
        if ci != co:
            add(q + "conv_downsample.weight", co, ci, 1)
    add(p + "5.weight", E, 512, 1)
    add(p + "5.bias", E)
    if cfg.get("multi_task", True):
        add("pred_linear.weight", cfg.get("spksInTrain", 251), E)
        add("pred_linear.bias", cfg.get("spksInTrain", 251))

    def plain_block(q):
        add(q + "conv1x1.weight", H, B, 1)
        add(q + "conv1x1.bias", H)
        add(q + "PReLU_1.weight", 1)
        add(q + "norm_1.weight", H, 1)
        add(q + "norm_1.bias", H, 1)
        add(q + "dwconv.weight", H, 1, P)
        add(q + "dwconv.bias", H)
        add(q + "PReLU_2.weight", 1)
        add(q + "norm_2.weight", H, 1)
        add(q + "norm_2.bias", H, 1)
        add(q + "Output.weight", B, H, 1)
        add(q + "Output.bias", B)

    for r in range(R):
        if ftype == "concatConv":
            q = f"separation.separation.{2 * r}."
            add(q + "conv1x1.weight", H, B + E, 1)
            add(q + "conv1x1.bias", H)
            add(q + "prelu1.weight", 1)
            add(q + "lnorm1.weight", H, 1)
            add(q + "lnorm1.bias", H, 1)
            add(q + "dconv.weight", H, 1, P)
            add(q + "dconv.bias", H)
            add(q + "prelu2.weight", 1)
            add(q + "lnorm2.weight", H, 1)
            add(q + "lnorm2.bias", H, 1)
            add(q + "sconv.weight", B, H, 1)
            add(q + "sconv.bias", B)
            for j in range(X - 1):
                plain_block(f"separation.separation.{2 * r + 1}.separation.{j}.")
        else:
            q = f"separation.separation.{4 * r}."
            if ftype == "concat":
                add(q + "fc.linear.weight", B, B + E)
                add(q + "fc.linear.bias", B)
            elif ftype in ("additive", "multiply"):
                add(q + "fc.linear.weight", B, E)
                add(q + "fc.linear.bias", B)
            elif ftype == "FiLM":
                add(q + "fc.gamma_fcs.0.weight", B, E)
                add(q + "fc.gamma_fcs.0.bias", B)
                add(q + "fc.beta_fcs.0.weight", B, E)
                add(q + "fc.beta_fcs.0.bias", B)
            add(f"separation.separation.{4 * r + 1}.weight", 1)
            add(f"separation.separation.{4 * r + 2}.weight", B, 1)
            add(f"separation.separation.{4 * r + 2}.bias", B, 1)
            for j in range(X):
                plain_block(f"separation.separation.{4 * r + 3}.separation.{j}.")
    for i in (1, 2, 3):
        add(f"decoder.mask{i}.weight", N, B, 1)
        add(f"decoder.mask{i}.bias", N)
    for i, k in zip((1, 2, 3), (L, 80, 160)):
        add(f"decoder.decoder_1d_{i}.weight", N, 1, k)
        add(f"decoder.decoder_1d_{i}.bias", 1)
    return spec


def make_state_dict(cfg, dtype=torch.float32, device="cpu"):
    """Zero-initialised state_dict with the reference's keys/shapes/order (fill with
    wesep_b200.synth.fill_state_dict_)."""
    sd = {}
    for k, shape in state_dict_spec(cfg):
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros(shape, dtype=torch.int64, device=device)
        else:
            sd[k] = torch.zeros(shape, dtype=dtype, device=device)
    return sd
