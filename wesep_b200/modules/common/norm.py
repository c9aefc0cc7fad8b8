"""Normalisation layers with the reference's names / parameters (wesep/modules/common/norm.py)."""
import torch
import torch.nn as nn

from wesep_b200 import ops


class GlobalChannelLayerNorm(nn.Module):
    """gLN — reference wesep/modules/common/norm.py:7-48.  Parameter holder: inside the TCN blocks
    the normalisation is fused into the neighbouring kernels (statistics in the producing GEMM /
    stencil epilogue, affine apply in the consuming kernel's prologue)."""

    def __init__(self, dim, eps=1e-05, elementwise_affine=True):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        if self.elementwise_affine:
            self.weight = nn.Parameter(torch.ones(self.dim, 1))
            self.bias = nn.Parameter(torch.zeros(self.dim, 1))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)

    def forward(self, x):
        if x.dim() != 3:
            raise RuntimeError("{} accept 3D tensor as input".format(self.__class__.__name__))
        raise NotImplementedError("stand-alone gLN is not part of the accelerated path: it only runs fused "
                                  "inside Conv1DBlock / Conv1DBlock4Fuse")


class ChannelWiseLayerNorm(nn.LayerNorm):
    """cLN — reference wesep/modules/common/norm.py:51-66 (LayerNorm over channels per frame)."""

    def forward(self, x):
        if x.dim() != 3:
            raise RuntimeError("{} accept 3D tensor as input".format(self.__class__.__name__))
        return ops.cln(x, self.weight, self.bias, self.eps)


class FiLM(nn.Module):
    """Feature-wise Linear Modulation — reference wesep/modules/common/norm.py:84-139 (gamma/beta Linears,
    zero-initialised).  Parameter holder: the modulation (1+gamma)*x + beta is applied by the fused fusion kernel."""

    def __init__(self, feat_size, embed_size, num_film_layers=1, layer_norm=False):
        super().__init__()
        if num_film_layers != 1 or layer_norm:
            raise NotImplementedError("FiLM: only num_film_layers=1, layer_norm=False (the reference's use) is accelerated")
        self.feat_size = feat_size
        self.embed_size = embed_size
        self.num_film_layers = num_film_layers
        self.layer_norm = None
        self.gamma_fcs = nn.ModuleList([nn.Linear(embed_size, feat_size)])
        self.beta_fcs = nn.ModuleList([nn.Linear(embed_size, feat_size)])
        self.init_weights()

    def init_weights(self):
        for i in range(self.num_film_layers):
            nn.init.zeros_(self.gamma_fcs[i].weight)
            nn.init.zeros_(self.gamma_fcs[i].bias)
            nn.init.zeros_(self.beta_fcs[i].weight)
            nn.init.zeros_(self.beta_fcs[i].bias)

    def row_affine(self, embed):
        """embed [n, E] -> (1 + gamma [n, feat], beta [n, feat])"""
        g = ops.LinearFn.apply(embed, self.gamma_fcs[0].weight, self.gamma_fcs[0].bias)
        b = ops.LinearFn.apply(embed, self.beta_fcs[0].weight, self.beta_fcs[0].bias)
        return 1 + g, b


def select_norm(norm, dim):
    """reference wesep/modules/common/norm.py:69-81"""
    if norm not in ["cLN", "gLN", "BN"]:
        raise RuntimeError("Unsupported normalize layer: {}".format(norm))
    if norm == "cLN":
        return ChannelWiseLayerNorm(dim, elementwise_affine=True)
    elif norm == "BN":
        return nn.BatchNorm1d(dim)
    else:
        return GlobalChannelLayerNorm(dim, elementwise_affine=True)
