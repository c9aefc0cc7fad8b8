"""SpeakerTransform — reference wesep/modules/common/speaker.py:26-49 (tiny 1x1 convs on the
[n, 256, 1] embedding; not a hot op, kept on torch)."""
from typing import Optional

import torch
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.common.norm import FiLM


class SpeakerTransform(nn.Module):

    def __init__(self, embed_dim=256, num_layers=3, hid_dim=128):
        super().__init__()
        layers = [nn.Conv1d(embed_dim, hid_dim, 1)]
        for _ in range(num_layers - 2):
            layers.append(nn.Conv1d(hid_dim, hid_dim, 1))
            layers.append(nn.Tanh())
        layers.append(nn.Conv1d(hid_dim, embed_dim, 1))
        self.transforms = nn.Sequential(*layers)

    def forward(self, x):
        if len(x.size()) == 2:
            return self.transforms(x.unsqueeze(-1)).squeeze(-1)
        return self.transforms(x)


class LinearLayer(nn.Module):
    """reference wesep/modules/common/speaker.py:52-60"""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features, bias)

    def forward(self, x, dummy: Optional[torch.Tensor] = None):
        return ops.LinearFn.apply(x, self.linear.weight, self.linear.bias)


class SpeakerFuseLayer(nn.Module):
    """reference wesep/modules/common/speaker.py:63-125 (3-D branches).  The reference expands the embedding over
    every frame and runs the Linear per frame; all four types reduce exactly to a per-row channel vector, computed
    once here (`prepare`) and applied by the fused fusion + PReLU + gLN kernel inside FuseSeparation."""

    def __init__(self, embed_dim=256, feat_dim=512, fuse_type="concat"):
        super().__init__()
        assert fuse_type in ["concat", "additive", "multiply", "FiLM", "None"]
        self.fuse_type = fuse_type
        self.feat_dim = feat_dim
        if fuse_type == "concat":
            self.fc = LinearLayer(embed_dim + feat_dim, feat_dim)
        elif fuse_type == "additive":
            self.fc = LinearLayer(embed_dim, feat_dim)
        elif fuse_type == "multiply":
            self.fc = LinearLayer(embed_dim, feat_dim)
        elif fuse_type == "FiLM":
            self.fc = FiLM(feat_dim, embed_dim)
        else:
            raise ValueError("Fuse type not defined.")

    def prepare(self, x, embed):
        """x [n, feat, T], embed [n, E, 1] -> (y0, ra, rb): fused value = ra * y0 + rb (None = identity)."""
        if x.dim() != 3:
            raise NotImplementedError("SpeakerFuseLayer: only the 3-D (ConvTasNet) branch is accelerated")
        e = embed.reshape(embed.shape[0], -1)
        if self.fuse_type == "concat":
            W = self.fc.linear.weight
            F_ = self.feat_dim
            rb = ops.LinearFn.apply(e, W[:, F_:], self.fc.linear.bias)          # embedding half -> per-row bias
            return ops.Conv1x1RowBiasFn.apply(x, W[:, :F_], rb), None, None
        if self.fuse_type == "additive":
            return x, None, self.fc(e)
        if self.fuse_type == "multiply":
            return x, self.fc(e), None
        ra, rb = self.fc.row_affine(e)
        return x, ra, rb

    def forward(self, x, embed):
        raise NotImplementedError("stand-alone SpeakerFuseLayer: it runs fused with the following PReLU + gLN "
                                  "inside FuseSeparation")
