"""wespeaker ECAPA-TDNN speaker encoder (ECAPA_TDNN_c512 / _GLOB_c512 / _c1024 / _GLOB_c1024, ASTP pooling) — the alternative
`spk_model` of the recipes (examples/librimix/tse/v2/confs/dpccn.yaml:59-64, bsrnn_feats.yaml: `spk_model: ECAPA_TDNN_GLOB_c512`,
`spk_args: {embed_dim: 192, feat_dim: 80, pooling_func: ASTP}`), SURVEY.md 8f-2.

wespeaker is an EXTERNAL package absent from the reference tree and from this image, so this module is restated from its
published architecture (wespeaker/models/ecapa_tdnn.py: Conv1dReluBn, Res2Conv1dReluBn, SE_Connect, SE_Res2Block;
pooling_layers.ASTP): attribute names, parameter shapes and `state_dict()` keys follow it so that its pretrained checkpoints
load key for key.  The reference holds no test or vector for it: parity is anchored on oracle/ecapa.py only — "parity unpinned".
Forward runs on libwesep_b200: dilated Conv1d = im2col1d + tcgen05 pointwise GEMM (ReLU in the epilogue), BatchNorm1d with batch
statistics, the SE gate, the global-context ASTP (context statistics folded into a per-row bias of the first attention
convolution).  The nn.* members are parameter containers, never called."""
import torch
import torch.nn as nn

from wesep_b200 import ops


def _bn(x, bn):
    if bn.training:
        bn.num_batches_tracked.add_(1)
    return ops.BnActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, None, False, bn.training, bn.momentum, bn.eps)


class Conv1dReluBn(nn.Module):
    """conv -> ReLU -> BatchNorm (this order)."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=True):
        super().__init__()
        if stride != 1 or padding != dilation * (kernel_size - 1) // 2:
            raise NotImplementedError("ECAPA: 'same' Conv1d layers are built")
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, stride, padding, dilation, bias=bias)
        self.bn = nn.BatchNorm1d(out_channels)

    def run(self, x):
        y = ops.conv1d_k(x, self.conv.weight, self.conv.bias, self.conv.dilation[0], act="relu")
        return _bn(y, self.bn)


class Res2Conv1dReluBn(nn.Module):
    """Res2Net branch: `scale` channel groups, group i > 0 adds the previous branch's output before its conv -> ReLU -> BN."""

    def __init__(self, channels, kernel_size=1, stride=1, padding=0, dilation=1, bias=True, scale=4):
        super().__init__()
        assert channels % scale == 0
        self.scale, self.width = scale, channels // scale
        self.nums = scale if scale == 1 else scale - 1
        self.convs = nn.ModuleList([nn.Conv1d(self.width, self.width, kernel_size, stride, padding, dilation, bias=bias)
                                    for _ in range(self.nums)])
        self.bns = nn.ModuleList([nn.BatchNorm1d(self.width) for _ in range(self.nums)])

    def run(self, x):
        w = self.width
        out, sp = [], None
        for i, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            part = x[:, i * w:(i + 1) * w]
            sp = part if i == 0 else ops.AddFn.apply(sp, part)
            sp = _bn(ops.conv1d_k(sp, conv.weight, conv.bias, conv.dilation[0], act="relu"), bn)
            out.append(sp)
        if self.scale != 1:
            out.append(ops.as_act(x[:, self.nums * w:]))
        return ops.cat_act(out)


class SE_Connect(nn.Module):
    def __init__(self, channels, se_bottleneck_dim=128):
        super().__init__()
        self.linear1 = nn.Linear(channels, se_bottleneck_dim)
        self.linear2 = nn.Linear(se_bottleneck_dim, channels)

    def run(self, x):
        s = ops.MeanTimeFn.apply(x)                                                  # [n, C]
        s = ops.UnaryFn.apply(ops.LinearFn.apply(s, self.linear1.weight, self.linear1.bias), 0)
        s = ops.UnaryFn.apply(ops.LinearFn.apply(s, self.linear2.weight, self.linear2.bias), 1)
        return ops.RowAffineFn.apply(x, s, None)


class SE_Res2Block(nn.Module):
    def __init__(self, channels, kernel_size, stride, padding, dilation, scale):
        super().__init__()
        self.se_res2block = nn.Sequential(
            Conv1dReluBn(channels, channels, kernel_size=1, stride=1, padding=0),
            Res2Conv1dReluBn(channels, kernel_size, stride, padding, dilation, scale=scale),
            Conv1dReluBn(channels, channels, kernel_size=1, stride=1, padding=0),
            SE_Connect(channels))

    def run(self, x):
        y = x
        for m in self.se_res2block:
            y = m.run(y)
        return ops.AddFn.apply(x, y)


class ASTP(nn.Module):
    """Attentive statistics pooling with (optionally) global context: parameters `linear1` / `linear2` (Conv1d k = 1)."""

    def __init__(self, in_dim, bottleneck_dim=128, global_context_att=False, **kwargs):
        super().__init__()
        self.in_dim, self.global_context_att = in_dim, global_context_att
        self.linear1 = nn.Conv1d(in_dim * 3 if global_context_att else in_dim, bottleneck_dim, kernel_size=1)
        self.linear2 = nn.Conv1d(bottleneck_dim, in_dim, kernel_size=1)

    def get_out_dim(self):
        return self.in_dim * 2

    def run(self, x):
        C = self.in_dim
        W1 = self.linear1.weight[:, :, 0]
        if self.global_context_att:
            # cat([x, mean.expand, std.expand]) through a 1x1 conv = W_x x + a per-row vector (W_m mean + W_s std + b)
            ctxv = ops.TstpFn.apply(x, 1e-10)                                          # [n, 2C]: mean | sqrt(unbiased var + 1e-10)
            rb = ops.LinearFn.apply(ctxv, W1[:, C:], self.linear1.bias)
            a = ops.Conv1x1RowBiasFn.apply(x, W1[:, :C], rb)
        else:
            a = ops.Conv1x1Fn.apply(x, W1, self.linear1.bias, False, None)
        a = ops.TanhFn.apply(a)
        a = ops.Conv1x1Fn.apply(a, self.linear2.weight[:, :, 0], self.linear2.bias, False, None)
        alpha = ops.SoftmaxFn.apply(a, 1.0)                                            # over time
        return ops.AstpFn.apply(x, alpha)


class ECAPA_TDNN(nn.Module):

    def __init__(self, channels=512, feat_dim=80, embed_dim=192, pooling_func="ASTP", global_context_att=False, emb_bn=False,
                 **kwargs):
        super().__init__()
        if pooling_func != "ASTP":
            raise NotImplementedError("ECAPA: pooling_func ASTP (the recipes' choice) is built")
        self.feat_dim, self.embed_dim = feat_dim, embed_dim
        self.layer1 = Conv1dReluBn(feat_dim, channels, kernel_size=5, padding=2)
        self.layer2 = SE_Res2Block(channels, kernel_size=3, stride=1, padding=2, dilation=2, scale=8)
        self.layer3 = SE_Res2Block(channels, kernel_size=3, stride=1, padding=3, dilation=3, scale=8)
        self.layer4 = SE_Res2Block(channels, kernel_size=3, stride=1, padding=4, dilation=4, scale=8)
        cat_channels = channels * 3
        out_channels = 512 * 3
        self.conv = nn.Conv1d(cat_channels, out_channels, kernel_size=1)
        self.pool = ASTP(in_dim=out_channels, global_context_att=global_context_att)
        self.pool_out_dim = self.pool.get_out_dim()
        self.bn = nn.BatchNorm1d(self.pool_out_dim)
        self.linear = nn.Linear(self.pool_out_dim, embed_dim)
        self.emb_bn = emb_bn
        self.bn2 = nn.BatchNorm1d(embed_dim) if emb_bn else nn.Identity()

    def _get_frame_level_feat(self, x):
        n, T, Fd = x.shape
        h = ops.new_act(n, Fd, T, x.device)
        h.copy_(x.float().permute(0, 2, 1))                                          # (B, T, F) -> (B, F, T)
        out1 = self.layer1.run(h)
        out2 = self.layer2.run(out1)
        out3 = self.layer3.run(out2)
        out4 = self.layer4.run(out3)
        out = ops.cat_act([out2, out3, out4])
        return ops.conv1x1_bigk_relu(out, self.conv.weight[:, :, 0], self.conv.bias)

    def forward(self, x):
        """x: fbank [n, frames, feat_dim] -> embedding [n, embed_dim]."""
        if x.dim() != 3 or x.shape[2] != self.feat_dim:
            raise RuntimeError("ECAPA_TDNN expects [batch, frames, %d] features" % self.feat_dim)
        if not x.is_cuda:
            raise RuntimeError("wesep_b200 kernels need CUDA tensors (no CPU fallback)")
        out = self._get_frame_level_feat(x)
        stats = self.pool.run(out)                                                    # [n, 2 * 1536]
        stats = _bn(stats.unsqueeze(-1), self.bn)[:, :, 0]
        emb = ops.LinearFn.apply(stats, self.linear.weight, self.linear.bias)
        if self.emb_bn:
            emb = _bn(emb.unsqueeze(-1), self.bn2)[:, :, 0]
        return emb


def ECAPA_TDNN_c1024(feat_dim, embed_dim, pooling_func="ASTP", emb_bn=False, **kw):
    return ECAPA_TDNN(channels=1024, feat_dim=feat_dim, embed_dim=embed_dim, pooling_func=pooling_func, emb_bn=emb_bn)


def ECAPA_TDNN_GLOB_c1024(feat_dim, embed_dim, pooling_func="ASTP", emb_bn=False, **kw):
    return ECAPA_TDNN(channels=1024, feat_dim=feat_dim, embed_dim=embed_dim, pooling_func=pooling_func, global_context_att=True,
                      emb_bn=emb_bn)


def ECAPA_TDNN_c512(feat_dim, embed_dim, pooling_func="ASTP", emb_bn=False, **kw):
    return ECAPA_TDNN(channels=512, feat_dim=feat_dim, embed_dim=embed_dim, pooling_func=pooling_func, emb_bn=emb_bn)


def ECAPA_TDNN_GLOB_c512(feat_dim, embed_dim, pooling_func="ASTP", emb_bn=False, **kw):
    return ECAPA_TDNN(channels=512, feat_dim=feat_dim, embed_dim=embed_dim, pooling_func=pooling_func, global_context_att=True,
                      emb_bn=emb_bn)
