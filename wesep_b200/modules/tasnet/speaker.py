"""ResNet4SpExplus / ResBlock — reference wesep/modules/tasnet/speaker.py:7-64, executed on the library's
kernels: pointwise convs as tensor-core GEMMs (BatchNorm batch statistics accumulated in their epilogues),
BN-apply + PReLU as the next GEMM's operand prologue, BN-apply + residual + PReLU + MaxPool1d(3) fused."""
import torch
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.common.norm import ChannelWiseLayerNorm
from wesep_b200.modules.tasnet.convs import Conv1D


class _PW(nn.Conv1d):
    """bias-free pointwise conv (nn.Conv1d parameters, GEMM execution)."""

    def forward(self, x):
        return ops.conv1x1(x, self.weight.reshape(self.out_channels, self.in_channels), self.bias)


class ResBlock(nn.Module):

    def __init__(self, in_dims, out_dims):
        super().__init__()
        self.conv1 = _PW(in_dims, out_dims, kernel_size=1, bias=False)
        self.conv2 = _PW(out_dims, out_dims, kernel_size=1, bias=False)
        self.batch_norm1 = nn.BatchNorm1d(out_dims)
        self.batch_norm2 = nn.BatchNorm1d(out_dims)
        self.prelu1 = nn.PReLU()
        self.prelu2 = nn.PReLU()
        self.mp = nn.MaxPool1d(3)
        if in_dims != out_dims:
            self.downsample = True
            self.conv_downsample = _PW(in_dims, out_dims, kernel_size=1, bias=False)
        else:
            self.downsample = False

    def forward(self, x):
        b1, b2 = self.batch_norm1, self.batch_norm2
        if self.training:
            with torch.no_grad():
                b1.num_batches_tracked += 1
                b2.num_batches_tracked += 1
        return ops.ResBlockFn.apply(x, self.conv1.weight, self.conv2.weight,
                                    self.conv_downsample.weight if self.downsample else None,
                                    b1.weight, b1.bias, b1.running_mean, b1.running_var,
                                    b2.weight, b2.bias, b2.running_mean, b2.running_var,
                                    self.prelu1.weight, self.prelu2.weight, self.training, b1.momentum, b1.eps)


class ResNet4SpExplus(nn.Module):

    def __init__(self, in_channel=256, C_embedding=256):
        super().__init__()
        self.aux_enc3 = nn.Sequential(
            ChannelWiseLayerNorm(3 * in_channel),
            Conv1D(3 * 256, 256, 1),
            ResBlock(256, 256),
            ResBlock(256, 512),
            ResBlock(512, 512),
            Conv1D(512, C_embedding, 1),
        )

    def forward(self, x):
        aux = self.aux_enc3(x)
        return ops.MeanTimeFn.apply(aux)
