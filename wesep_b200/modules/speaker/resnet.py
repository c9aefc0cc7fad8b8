"""wespeaker ResNet speaker encoder (ResNet18 / ResNet34, TSTP pooling) — the `spk_model` pBSRNN trains jointly
(reference: `get_speaker_model(spk_model)(**spk_args)`, wesep/models/bsrnn.py:217; examples/librimix/tse/v2/confs/bsrnn.yaml:56-64).

wespeaker is an EXTERNAL package absent from the reference tree and from this image, so this module is
restated from its published architecture (wespeaker/models/resnet.py, pooling_layers.TSTP): attribute names, parameter
shapes and `state_dict()` keys follow it so that its pretrained checkpoints (`spk_model_init`) load key for key.  The
reference holds no test or vector for it: parity is anchored on oracle/resnet.py (plain torch) only — "parity unpinned".
Forward runs on libwesep_b200 (im2col + tcgen05 GEMM, fused BatchNorm / residual / ReLU, TSTP); the nn.Conv2d /
nn.BatchNorm2d / nn.Linear objects are parameter containers (same default initialisation as wespeaker's), never called."""
import torch
import torch.nn as nn

from wesep_b200 import ops


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, in_planes, planes, stride=1):
        super().__init__()
        self.stride = stride
        self.conv1 = nn.Conv2d(in_planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.shortcut = nn.Sequential()
        if stride != 1 or in_planes != self.expansion * planes:
            self.shortcut = nn.Sequential(nn.Conv2d(in_planes, self.expansion * planes, kernel_size=1, stride=stride, bias=False),
                                          nn.BatchNorm2d(self.expansion * planes))


class TSTP(nn.Module):
    """Temporal statistics pooling; no parameters (kept as a module so the attribute path `pool` exists)."""

    def __init__(self, in_dim=0, **kwargs):
        super().__init__()
        self.in_dim = in_dim

    def get_out_dim(self):
        return self.in_dim * 2


def _conv3x3(x, H, W, conv):
    s = conv.stride[0]
    col = ops.Im2Col3x3Fn.apply(x, H, W, s)
    w2d = conv.weight.reshape(conv.weight.shape[0], -1)
    if col.shape[1] != w2d.shape[1]:                   # im2col pads the gathered channels to a multiple of 16 (zero rows)
        w2d = torch.nn.functional.pad(w2d, (0, col.shape[1] - w2d.shape[1]))
    y = ops.conv1x1_bigk(col, w2d)
    return y, (H - 1) // s + 1, (W - 1) // s + 1


def _bn(x, bn, res=None, relu=True):
    if bn.training:
        bn.num_batches_tracked.add_(1)
    return ops.BnActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, res, relu, bn.training, bn.momentum, bn.eps)


class ResNet(nn.Module):

    def __init__(self, block, num_blocks, m_channels=32, feat_dim=40, embed_dim=128, pooling_func="TSTP", two_emb_layer=True):
        super().__init__()
        if block is not BasicBlock:
            raise NotImplementedError("only the BasicBlock ResNets (ResNet18 / ResNet34) are built")
        if pooling_func != "TSTP":
            raise NotImplementedError("pooling_func: only TSTP is built (bsrnn.yaml:62)")
        if two_emb_layer:
            raise NotImplementedError("two_emb_layer=True is not on the recipe path (bsrnn.yaml:63)")
        self.in_planes = m_channels
        self.feat_dim, self.embed_dim = feat_dim, embed_dim
        self.stats_dim = int(feat_dim / 8) * m_channels * 8
        self.two_emb_layer = two_emb_layer
        self.conv1 = nn.Conv2d(1, m_channels, kernel_size=3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(m_channels)
        self.layer1 = self._make_layer(block, m_channels, num_blocks[0], stride=1)
        self.layer2 = self._make_layer(block, m_channels * 2, num_blocks[1], stride=2)
        self.layer3 = self._make_layer(block, m_channels * 4, num_blocks[2], stride=2)
        self.layer4 = self._make_layer(block, m_channels * 8, num_blocks[3], stride=2)
        self.pool = TSTP(in_dim=self.stats_dim * block.expansion)
        self.pool_out_dim = self.pool.get_out_dim()
        self.seg_1 = nn.Linear(self.pool_out_dim, embed_dim)
        self.seg_bn_1 = nn.Identity()
        self.seg_2 = nn.Identity()

    def _make_layer(self, block, planes, num_blocks, stride):
        layers = []
        for s in [stride] + [1] * (num_blocks - 1):
            layers.append(block(self.in_planes, planes, s))
            self.in_planes = planes * block.expansion
        return nn.Sequential(*layers)

    def forward(self, x):
        """x: fbank [n, frames, feat_dim] -> (tensor(0.0), embedding [n, embed_dim]) as wespeaker returns it."""
        if x.dim() != 3 or x.shape[2] != self.feat_dim:
            raise RuntimeError("ResNet expects [batch, frames, %d] features" % self.feat_dim)
        if not x.is_cuda:
            raise RuntimeError("wesep_b200 kernels need CUDA tensors (no CPU fallback)")
        n, T, F = x.shape
        H, W = F, T
        h = ops.new_act(n, 1, H * W, x.device)
        h.copy_(x.float().permute(0, 2, 1).reshape(n, 1, H * W))     # (B, T, F) -> (B, 1, F, T), time contiguous
        h, H, W = _conv3x3(h, H, W, self.conv1)
        h = _bn(h, self.bn1)
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                y, Ho, Wo = _conv3x3(h, H, W, blk.conv1)
                y = _bn(y, blk.bn1)
                y, _, _ = _conv3x3(y, Ho, Wo, blk.conv2)
                if len(blk.shortcut) > 0:
                    sc = ops.Subsample2dFn.apply(h, H, W, blk.stride) if blk.stride != 1 else h
                    sc = ops.conv1x1_bigk(sc, blk.shortcut[0].weight.reshape(blk.shortcut[0].weight.shape[0], -1))
                    sc = _bn(sc, blk.shortcut[1], relu=False)
                else:
                    sc = h
                h = _bn(y, blk.bn2, res=sc, relu=True)               # relu(bn2(conv2) + shortcut)
                H, W = Ho, Wo
        C = h.shape[1]
        ld = h.stride(1)
        if ld != H * W:                                              # [n, C, H*W] -> rows (c, h), W valid columns each
            hc = ops.new_act(n, C * H, W, x.device)
            hc.copy_(h.reshape(n, C * H, W))
        else:
            hc = ops.as_act(h.reshape(n, C * H, W))
        stats = ops.TstpFn.apply(hc)                                 # [n, 2 * C * H]
        emb = ops.LinearFn.apply(stats, self.seg_1.weight, self.seg_1.bias)
        return torch.zeros((), device=x.device), emb


def ResNet18(feat_dim, embed_dim, pooling_func="TSTP", two_emb_layer=True):
    return ResNet(BasicBlock, [2, 2, 2, 2], feat_dim=feat_dim, embed_dim=embed_dim, pooling_func=pooling_func, two_emb_layer=two_emb_layer)


def ResNet34(feat_dim, embed_dim, pooling_func="TSTP", two_emb_layer=True):
    return ResNet(BasicBlock, [3, 4, 6, 3], feat_dim=feat_dim, embed_dim=embed_dim, pooling_func=pooling_func, two_emb_layer=two_emb_layer)


def get_speaker_model(model_name):
    """wespeaker.models.speaker_model.get_speaker_model for the encoders built here."""
    if model_name in ("ResNet18", "ResNet34"):
        return globals()[model_name]
    if model_name in ("ECAPA_TDNN_c512", "ECAPA_TDNN_GLOB_c512", "ECAPA_TDNN_c1024", "ECAPA_TDNN_GLOB_c1024"):
        from wesep_b200.modules.speaker import ecapa
        return getattr(ecapa, model_name)
    raise NotImplementedError("speaker model %s is not built in wesep_b200 (ResNet18 / ResNet34 / ECAPA_TDNN_* only)" % model_name)
