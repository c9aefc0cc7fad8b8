"""ORACLE (test infrastructure) — CPU restatement of wespeaker's ECAPA-TDNN.

PARITY UNPINNED: wespeaker is an external package that is neither in /root/reference nor installable here, and the reference
holds no vector for it; this file restates the published architecture (wespeaker/models/ecapa_tdnn.py: Conv1dReluBn = conv ->
ReLU -> BatchNorm, Res2Conv1dReluBn with scale 8, SE_Connect, SE_Res2Block with the residual, channel concatenation of the three
blocks, 1x1 conv + ReLU; pooling_layers.ASTP with global context; BatchNorm1d; Linear) in plain torch functional ops; the
parameter count of ECAPA_TDNN_GLOB_c512 (6.19 M) matches the published figure.  The CUDA module is tested against this file.
"""
import torch
import torch.nn.functional as F


def bn(x, sd, pre, training, eps=1e-5):
    """nn.BatchNorm1d on [n, C, T] or [n, C]: batch statistics in training (running buffers are not updated here)."""
    dims = (0, 2) if x.dim() == 3 else (0,)
    if training:
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
    else:
        mean, var = sd[pre + "running_mean"], sd[pre + "running_var"]
    shape = (1, -1, 1) if x.dim() == 3 else (1, -1)
    return (x - mean.view(shape)) / torch.sqrt(var.view(shape) + eps) * sd[pre + "weight"].view(shape) + sd[pre + "bias"].view(shape)


def conv_relu_bn(x, sd, pre, dilation, training):
    w = sd[pre + "conv.weight"]
    y = F.conv1d(x, w, sd[pre + "conv.bias"], padding=dilation * (w.shape[2] - 1) // 2, dilation=dilation)
    return bn(F.relu(y), sd, pre + "bn.", training)


def res2(x, sd, pre, dilation, scale, training):
    width = x.shape[1] // scale
    spx = torch.split(x, width, 1)
    out, sp = [], None
    for i in range(scale - 1):
        sp = spx[i] if i == 0 else sp + spx[i]
        w = sd[pre + f"convs.{i}.weight"]
        sp = F.conv1d(sp, w, sd[pre + f"convs.{i}.bias"], padding=dilation * (w.shape[2] - 1) // 2, dilation=dilation)
        sp = bn(F.relu(sp), sd, pre + f"bns.{i}.", training)
        out.append(sp)
    out.append(spx[scale - 1])
    return torch.cat(out, 1)


def se_res2block(x, sd, pre, dilation, training, scale=8):
    y = conv_relu_bn(x, sd, pre + "se_res2block.0.", 1, training)
    y = res2(y, sd, pre + "se_res2block.1.", dilation, scale, training)
    y = conv_relu_bn(y, sd, pre + "se_res2block.2.", 1, training)
    s = y.mean(2)
    s = F.relu(F.linear(s, sd[pre + "se_res2block.3.linear1.weight"], sd[pre + "se_res2block.3.linear1.bias"]))
    s = torch.sigmoid(F.linear(s, sd[pre + "se_res2block.3.linear2.weight"], sd[pre + "se_res2block.3.linear2.bias"]))
    return x + y * s.unsqueeze(2)


def astp(x, sd, pre, global_context):
    if global_context:
        m = x.mean(-1, keepdim=True).expand_as(x)
        s = torch.sqrt(x.var(-1, keepdim=True) + 1e-10).expand_as(x)
        x_in = torch.cat((x, m, s), 1)
    else:
        x_in = x
    a = torch.tanh(F.conv1d(x_in, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"]))
    a = torch.softmax(F.conv1d(a, sd[pre + "linear2.weight"], sd[pre + "linear2.bias"]), dim=2)
    mean = torch.sum(a * x, 2)
    var = torch.sum(a * x ** 2, 2) - mean ** 2
    return torch.cat([mean, torch.sqrt(var.clamp(min=1e-10))], 1)


def ecapa_forward(sd, feats, global_context=True, training=True):
    """feats [n, frames, feat_dim] -> embedding [n, embed_dim]."""
    x = feats.permute(0, 2, 1)
    o1 = conv_relu_bn(x, sd, "layer1.", 1, training)
    o2 = se_res2block(o1, sd, "layer2.", 2, training)
    o3 = se_res2block(o2, sd, "layer3.", 3, training)
    o4 = se_res2block(o3, sd, "layer4.", 4, training)
    out = F.relu(F.conv1d(torch.cat([o2, o3, o4], 1), sd["conv.weight"], sd["conv.bias"]))
    st = astp(out, sd, "pool.", global_context)
    st = bn(st, sd, "bn.", training)
    return F.linear(st, sd["linear.weight"], sd["linear.bias"])
