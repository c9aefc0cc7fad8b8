"""Separation / FuseSeparation — reference wesep/modules/tasnet/separation.py:8-186."""
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.common import select_norm
from wesep_b200.modules.common.norm import GlobalChannelLayerNorm
from wesep_b200.modules.common.speaker import SpeakerFuseLayer
from wesep_b200.modules.tasnet.convs import Conv1DBlock, Conv1DBlock4Fuse


class Separation(nn.Module):

    def __init__(self, R, X, B, H, P, norm="gLN", causal=False, skip_con=True, start_dilation=0):
        super().__init__()
        self.separation = nn.ModuleList([])
        for _ in range(R):
            for x in range(start_dilation, X):
                self.separation.append(Conv1DBlock(B, H, P, 2 ** x, norm, causal, skip_con))
        self.skip_con = skip_con

    def forward(self, x):
        if self.skip_con:
            raise NotImplementedError("skip_con=True is outside the accelerated path (recipes use skip_con: False)")
        for blk in self.separation:
            x = blk(x)
        return x


class FuseSeparation(nn.Module):

    def __init__(self, R, X, B, H, P, norm="gLN", causal=False, skip_con=False, C_embedding=256,
                 spk_fuse_type="concatConv", multi_fuse=True):
        super().__init__()
        self.multi_fuse = multi_fuse
        self.spk_fuse_type = spk_fuse_type
        self.separation = nn.ModuleList([])
        if not multi_fuse:
            # the reference itself is broken for multi_fuse=False (separation.py:136-164 overwrites the list)
            raise NotImplementedError("multi_fuse=False is not runnable in the reference either (SURVEY App. C.15)")
        for _ in range(R):
            if spk_fuse_type == "concatConv":
                self.separation.append(Conv1DBlock4Fuse(spk_embed_dim=C_embedding, in_channels=B, conv_channels=H,
                                                        kernel_size=P, norm=norm, causal=causal, dilation=1))
                self.separation.append(Separation(1, X, B, H, P, norm=norm, causal=causal, skip_con=skip_con,
                                                  start_dilation=1))
            else:   # separation.py:116-135: [SpeakerFuseLayer, PReLU, norm, Separation(1, X)]
                self.separation.append(SpeakerFuseLayer(embed_dim=C_embedding, feat_dim=B, fuse_type=spk_fuse_type))
                self.separation.append(nn.PReLU())
                self.separation.append(select_norm(norm, B))
                self.separation.append(Separation(1, X, B, H, P, norm=norm, causal=causal, skip_con=skip_con))

    def forward(self, x, spk_embedding):
        if self.spk_fuse_type == "concatConv":
            for i in range(len(self.separation)):
                if i % 2 == 0:
                    x = self.separation[i](x, spk_embedding)
                else:
                    x = self.separation[i](x)
            return x
        for i in range(0, len(self.separation), 4):
            fuse, act, norm, sep = (self.separation[i + j] for j in range(4))
            if not isinstance(norm, GlobalChannelLayerNorm):
                raise NotImplementedError("only norm='gLN' is accelerated")
            y0, ra, rb = fuse.prepare(x, spk_embedding)
            x = ops.FusePreluGlnFn.apply(y0, ra, rb, act.weight, norm.weight, norm.bias, 0.0)
            x = sep(x)
        return x
