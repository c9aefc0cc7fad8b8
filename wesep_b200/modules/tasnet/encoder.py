"""MultiEncoder — reference wesep/modules/tasnet/encoder.py:63-114."""
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.common import select_norm
from wesep_b200.modules.tasnet.convs import Conv1D


class MultiEncoder(nn.Module):

    def __init__(self, in_channels, middle_channels, out_channels, kernel_size, stride):
        super().__init__()
        self.L1 = kernel_size
        self.L2 = 80
        self.L3 = 160
        self.stride = stride
        self.encoder_1d_short = Conv1D(in_channels, middle_channels, self.L1, stride=stride, padding=0)
        self.encoder_1d_middle = Conv1D(in_channels, middle_channels, self.L2, stride=stride, padding=0)
        self.encoder_1d_long = Conv1D(in_channels, middle_channels, self.L3, stride=stride, padding=0)
        self.ln = select_norm("cLN", 3 * middle_channels)
        self.proj = Conv1D(3 * middle_channels, out_channels, 1)
        self.middle_channels = middle_channels

    def filterbank(self, x):
        """cat([w1, w2, w3], 1) = [n, 3N, K]: the three ReLU'd strided convs (encoder.py:99-109)."""
        if x.dim() == 3:
            x = x.squeeze(1)
        if x.dim() != 2:
            raise RuntimeError("MultiEncoder accepts [n, T] (or [n, 1, T]) input")
        e = (self.encoder_1d_short, self.encoder_1d_middle, self.encoder_1d_long)
        return ops.MultiEncoderConvFn.apply(x.float(), self.stride, e[0].weight, e[0].bias, e[1].weight, e[1].bias,
                                            e[2].weight, e[2].bias)

    def forward(self, x):
        """returns (e [n,B,K], w1, w2, w3 [n,N,K]) like the reference; w_i are channel slices of one tensor."""
        w = self.filterbank(x)
        N = self.middle_channels
        y = self.ln(w)
        y = self.proj(y)
        return y, w[:, :N], w[:, N:2 * N], w[:, 2 * N:]
