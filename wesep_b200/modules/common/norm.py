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
