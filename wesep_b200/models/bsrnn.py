"""pBSRNN — parameter / constructor contract of the reference wesep/models/bsrnn.py:151-298 (class name, kwargs,
attribute paths, parameter shapes and order, hence `state_dict()` keys, checkpoint and optimizer-state layout).

STATUS: contract only.  The CUDA path for this model is not built yet — the oracle
(`oracle/bsrnn.py`) and its golden fixtures are; DESIGN.md §7.5 has the kernel plan.  `forward` raises
NotImplementedError: there is no PyTorch / CPU fallback in this package.  The nn.GroupNorm / nn.LSTM / nn.Linear /
nn.Conv1d objects below are parameter containers (same initialisation as the reference); their own forwards are never
called.
"""
import numpy as np
import torch
import torch.nn as nn


class _FuseFC(nn.Module):
    """SpeakerFuseLayer's `fc = LinearLayer(...)` (wesep/modules/common/speaker.py:52-79): keys `fc.linear.{weight,bias}`."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features)


class SpeakerFuseLayer(nn.Module):
    def __init__(self, embed_dim=256, feat_dim=512, fuse_type="concat"):
        super().__init__()
        if fuse_type not in ("concat", "additive", "multiply"):
            raise NotImplementedError("pBSRNN fuse types: concat / additive / multiply (FiLM is a Spex+ option)")
        self.fuse_type = fuse_type
        self.fc = _FuseFC(embed_dim + feat_dim if fuse_type == "concat" else embed_dim, feat_dim)


class ResRNN(nn.Module):
    """bsrnn.py:16-36: GroupNorm(1, C, eps=fp32 eps) -> bidirectional LSTM(C, 2C) -> Linear(4C, C), residual."""

    def __init__(self, input_size, hidden_size, bidirectional=True):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.eps = torch.finfo(torch.float32).eps
        self.norm = nn.GroupNorm(1, input_size, self.eps)
        self.rnn = nn.LSTM(input_size, hidden_size, 1, batch_first=True, bidirectional=bidirectional)
        self.proj = nn.Linear(hidden_size * 2, input_size)


class BSNet(nn.Module):
    """bsrnn.py:55-69."""

    def __init__(self, in_channel, nband=7, bidirectional=True):
        super().__init__()
        self.nband = nband
        self.feature_dim = in_channel // nband
        self.band_rnn = ResRNN(self.feature_dim, self.feature_dim * 2, bidirectional=bidirectional)
        self.band_comm = ResRNN(self.feature_dim, self.feature_dim * 2, bidirectional=bidirectional)


class FuseSeparation(nn.Module):
    """bsrnn.py:86-123."""

    def __init__(self, nband=7, num_repeat=6, feature_dim=128, spk_emb_dim=256, spk_fuse_type="concat", multi_fuse=True):
        super().__init__()
        self.multi_fuse, self.nband, self.feature_dim = multi_fuse, nband, feature_dim
        self.separation = nn.ModuleList([])
        if multi_fuse:
            for _ in range(num_repeat):
                self.separation.append(SpeakerFuseLayer(spk_emb_dim, feature_dim, spk_fuse_type))
                self.separation.append(BSNet(nband * feature_dim, nband))
        else:
            self.separation.append(SpeakerFuseLayer(spk_emb_dim, feature_dim, spk_fuse_type))
            for _ in range(num_repeat):
                self.separation.append(BSNet(nband * feature_dim, nband))


class BSRNN(nn.Module):

    def __init__(
        self,
        spk_emb_dim=256,
        sr=16000,
        win=512,
        stride=128,
        feature_dim=128,
        num_repeat=6,
        use_spk_transform=True,
        use_bidirectional=True,
        spk_fuse_type="concat",
        multi_fuse=True,
        joint_training=True,
        multi_task=False,
        spksInTrain=251,
        spk_model=None,
        spk_model_init=None,
        spk_model_freeze=False,
        spk_args=None,
        spk_feat=False,
        feat_type="consistent",
    ):
        super().__init__()
        if joint_training:
            raise NotImplementedError("pBSRNN with joint_training=True needs the wespeaker speaker encoder (SURVEY.md §8 row "
                                      "a22, not built); construct with joint_training=False and pass embeddings")
        if use_spk_transform:
            raise NotImplementedError("use_spk_transform=True is not on the recipe path (bsrnn.yaml:54)")
        if not use_bidirectional:
            raise NotImplementedError("the recipes use bidirectional LSTMs")
        self.sr, self.win, self.stride = sr, win, stride
        self.enc_dim = win // 2 + 1
        self.feature_dim = feature_dim
        self.eps = torch.finfo(torch.float32).eps
        self.spk_emb_dim = spk_emb_dim
        self.joint_training, self.multi_task = joint_training, multi_task
        # band layout, bsrnn.py:228-242
        bw100 = int(np.floor(100 / (sr / 2.0) * self.enc_dim))
        bw200 = int(np.floor(200 / (sr / 2.0) * self.enc_dim))
        bw500 = int(np.floor(500 / (sr / 2.0) * self.enc_dim))
        bw2k = int(np.floor(2000 / (sr / 2.0) * self.enc_dim))
        self.band_width = [bw100] * 15 + [bw200] * 10 + [bw500] * 5 + [bw2k]
        self.band_width.append(self.enc_dim - int(np.sum(self.band_width)))
        self.nband = len(self.band_width)
        self.spk_transform = nn.Identity()
        self.BN = nn.ModuleList([nn.Sequential(nn.GroupNorm(1, bw * 2, self.eps), nn.Conv1d(bw * 2, feature_dim, 1))
                                 for bw in self.band_width])
        self.separator = FuseSeparation(nband=self.nband, num_repeat=num_repeat, feature_dim=feature_dim,
                                        spk_emb_dim=spk_emb_dim, spk_fuse_type=spk_fuse_type, multi_fuse=multi_fuse)
        self.mask = nn.ModuleList([
            nn.Sequential(nn.GroupNorm(1, feature_dim, self.eps), nn.Conv1d(feature_dim, feature_dim * 4, 1), nn.Tanh(),
                          nn.Conv1d(feature_dim * 4, feature_dim * 4, 1), nn.Tanh(),
                          nn.Conv1d(feature_dim * 4, bw * 4, 1)) for bw in self.band_width])

    def forward(self, input, embeddings):
        raise NotImplementedError("wesep_b200.models.BSRNN: the sm_100a kernels for the pBSRNN path are not built yet "
                                  "(DESIGN.md §2, §7.5); there is no PyTorch fallback")
