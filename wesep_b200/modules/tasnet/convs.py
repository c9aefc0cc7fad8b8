"""TCN building blocks with the reference's class names, ctor arguments and parameter names
(wesep/modules/tasnet/convs.py), executing as fused sm_100a kernels."""
import torch
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.common import select_norm
from wesep_b200.modules.common.norm import GlobalChannelLayerNorm


class Conv1D(nn.Conv1d):
    """reference convs.py:9-22.  Pointwise (kernel_size 1) convs run as tensor-core GEMMs; the
    strided encoder filters are driven by MultiEncoder (framing + GEMM)."""

    def forward(self, x, squeeze=False):
        if x.dim() not in [2, 3]:
            raise RuntimeError("{} accept 2/3D tensor as input".format(self.__class__.__name__))
        x = x if x.dim() == 3 else torch.unsqueeze(x, 1)
        if self.kernel_size[0] == 1 and self.stride[0] == 1 and self.groups == 1 and self.padding[0] == 0:
            y = ops.conv1x1(x, self.weight.reshape(self.out_channels, self.in_channels), self.bias)
        elif self.in_channels == 1:
            raise NotImplementedError("stand-alone strided Conv1D: driven by MultiEncoder (framing + GEMM)")
        else:
            raise NotImplementedError("Conv1D configuration outside the accelerated path")
        if squeeze:
            y = torch.squeeze(y)
        return y


class ConvTrans1D(nn.ConvTranspose1d):
    """reference convs.py:25-40 — parameter holder; executed by MultiDecoder (basis GEMM + overlap-add)."""

    def forward(self, x, squeeze=False):
        raise NotImplementedError("stand-alone ConvTrans1D: use MultiDecoder")


def _require_supported(norm_layer, causal):
    if causal:
        raise NotImplementedError("causal TCN blocks are outside the accelerated path (recipes use causal: false)")
    if not isinstance(norm_layer, GlobalChannelLayerNorm):
        raise NotImplementedError("only norm='gLN' TCN blocks are accelerated (recipes use gLN)")


class Conv1DBlock(nn.Module):
    """reference convs.py:43-104 (skip_con False): 1x1 conv - PReLU - gLN - depthwise dilated conv -
    PReLU - gLN - 1x1 conv + residual, as ONE fused op (three kernels forward)."""

    def __init__(self, in_channels=256, out_channels=512, kernel_size=3, dilation=1, norm="gln", causal=False,
                 skip_con=True):
        super().__init__()
        self.conv1x1 = Conv1D(in_channels, out_channels, 1)
        self.PReLU_1 = nn.PReLU()
        self.norm_1 = select_norm(norm, out_channels)
        self.pad = ((dilation * (kernel_size - 1)) // 2 if not causal else (dilation * (kernel_size - 1)))
        self.dwconv = Conv1D(out_channels, out_channels, kernel_size, groups=out_channels, padding=self.pad,
                             dilation=dilation)
        self.PReLU_2 = nn.PReLU()
        self.norm_2 = select_norm(norm, out_channels)
        if skip_con:
            self.Sc_conv = nn.Conv1d(out_channels, in_channels, 1, bias=True)
        self.Output = nn.Conv1d(out_channels, in_channels, 1, bias=True)
        self.causal = causal
        self.skip_con = skip_con
        self.dilation = dilation
        self.kernel_size = kernel_size

    def forward(self, x):
        _require_supported(self.norm_1, self.causal)
        if self.skip_con:
            raise NotImplementedError("skip_con=True is outside the accelerated path (recipes use skip_con: False)")
        if self.kernel_size != 3:
            raise NotImplementedError("depthwise kernel size must be 3 (P=3)")
        return ops.tcn_block(x, None, self.conv1x1.weight, self.conv1x1.bias, self.PReLU_1.weight, self.norm_1.weight,
                             self.norm_1.bias, self.dwconv.weight, self.dwconv.bias, self.PReLU_2.weight,
                             self.norm_2.weight, self.norm_2.bias, self.Output.weight, self.Output.bias, self.dilation)


class Conv1DBlock4Fuse(nn.Module):
    """reference convs.py:107-160.  The concat([x, aux.repeat(T)]) -> 1x1 conv is computed exactly as
    W[:, :B] x + (W[:, B:] aux + b): the speaker half becomes a per-row bias (no [n, B+E, T] tensor)."""

    def __init__(self, in_channels=256, spk_embed_dim=100, conv_channels=512, kernel_size=3, dilation=1, norm="cLN",
                 causal=False):
        super().__init__()
        self.conv1x1 = Conv1D(in_channels + spk_embed_dim, conv_channels, 1)
        self.prelu1 = nn.PReLU()
        self.lnorm1 = select_norm(norm, conv_channels)
        dconv_pad = ((dilation * (kernel_size - 1)) // 2 if not causal else (dilation * (kernel_size - 1)))
        self.dconv = nn.Conv1d(conv_channels, conv_channels, kernel_size, groups=conv_channels, padding=dconv_pad,
                               dilation=dilation, bias=True)
        self.prelu2 = nn.PReLU()
        self.lnorm2 = select_norm(norm, conv_channels)
        self.sconv = nn.Conv1d(conv_channels, in_channels, 1, bias=True)
        self.causal = causal
        self.dconv_pad = dconv_pad
        self.dilation = dilation
        self.kernel_size = kernel_size

    def forward(self, x, aux):
        _require_supported(self.lnorm1, self.causal)
        if self.kernel_size != 3:
            raise NotImplementedError("depthwise kernel size must be 3 (P=3)")
        return ops.tcn_block(x, aux, self.conv1x1.weight, self.conv1x1.bias, self.prelu1.weight, self.lnorm1.weight,
                             self.lnorm1.bias, self.dconv.weight, self.dconv.bias, self.prelu2.weight,
                             self.lnorm2.weight, self.lnorm2.bias, self.sconv.weight, self.sconv.bias, self.dilation)
