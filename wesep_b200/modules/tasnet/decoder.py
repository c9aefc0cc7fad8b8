"""MultiDecoder — reference wesep/modules/tasnet/decoder.py:60-114."""
import torch
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.tasnet.convs import Conv1D, ConvTrans1D


class MultiDecoder(nn.Module):

    def __init__(self, in_channels, middle_channels, out_channels, kernel_size, stride):
        super().__init__()
        B, N, L = in_channels, middle_channels, kernel_size
        self.mask1 = Conv1D(B, N, 1)
        self.mask2 = Conv1D(B, N, 1)
        self.mask3 = Conv1D(B, N, 1)
        self.decoder_1d_1 = ConvTrans1D(N, out_channels, kernel_size=L, stride=stride, bias=True)
        self.decoder_1d_2 = ConvTrans1D(N, out_channels, kernel_size=80, stride=stride, bias=True)
        self.decoder_1d_3 = ConvTrans1D(N, out_channels, kernel_size=160, stride=stride, bias=True)
        self.stride = stride
        self.L1 = L
        if out_channels != 1:
            raise NotImplementedError("MultiDecoder: out_channels must be 1")

    def forward(self, x, w1, w2, w3, actLayer):
        """x [n,B,K]; w1..3 [n,N,K] (channel slices of the encoder filterbank output, or separate tensors)."""
        if not isinstance(actLayer, nn.ReLU):
            raise NotImplementedError("MultiDecoder: only activate='relu' masks are accelerated (recipe setting)")
        base = getattr(w1, "_base", None)
        if (base is not None and base is getattr(w2, "_base", None) and base is getattr(w3, "_base", None)
                and base.dim() == 3 and base.shape[1] == 3 * w1.shape[1] and w1.data_ptr() == base.data_ptr()):
            w_cat = base[:, :, :w1.shape[2]] if base.shape[2] != w1.shape[2] else base
        else:
            w_cat = torch.cat([w1, w2, w3], 1)
        return self.forward_cat(x, w_cat)

    def forward_cat(self, x, w_cat):
        K = x.shape[-1]
        xlen = (K - 1) * self.stride + self.L1                       # len(est1); est2/3 trimmed to it (:105-108)
        S = ops.DecoderMasksFn.apply(x, w_cat, self.mask1.weight, self.mask1.bias, self.mask2.weight, self.mask2.bias,
                                     self.mask3.weight, self.mask3.bias)
        d = (self.decoder_1d_1, self.decoder_1d_2, self.decoder_1d_3)
        ests = ops.DecoderBasisFn.apply(S, self.stride, xlen, d[0].weight, d[0].bias, d[1].weight, d[1].bias,
                                        d[2].weight, d[2].bias)
        return list(ests)
