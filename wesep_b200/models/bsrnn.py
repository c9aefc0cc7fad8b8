"""pBSRNN — parameter / constructor contract of the reference wesep/models/bsrnn.py:151-298 (class name, kwargs,
attribute paths, parameter shapes and order, hence `state_dict()` keys, checkpoint and optimizer-state layout).

STATUS: SURVEY.md §8 rows a15-a22: fuse types multiply / additive / concat, `joint_training` False (embeddings in) or True
(fbank in, wespeaker ResNet18 / ResNet34 + TSTP trained jointly, `multi_task` head included).  The forward runs on
libwesep_b200: STFT / iSTFT as framing + windowed-DFT GEMMs, band split as GroupNorm + one block-diagonal GEMM, ResRNN in
a time-major layout (input / output projections on the tcgen05 conv1x1 GEMMs, the recurrence of both directions and all
steps as ONE persistent cluster kernel per pass — `ops.LstmTmFn`, csrc/lstm_rec.cu), mask MLPs as GEMMs with tanh /
gating kernels.  There is no PyTorch
fallback: the nn.GroupNorm / nn.LSTM / nn.Linear / nn.Conv1d objects below are parameter containers (same initialisation
as the reference); their own forwards are never called.  Small torch glue remains (reflect padding, weight assembly,
concatenation of band slices, the 1 / sum(w^2) envelope).
"""
import math

import numpy as np
import torch
import torch.nn as nn

from wesep_b200 import ops


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
        self.spk_feat, self.feat_type, self.spk_model_freeze = spk_feat, feat_type, spk_model_freeze
        # band layout, bsrnn.py:228-242
        bw100 = int(np.floor(100 / (sr / 2.0) * self.enc_dim))
        bw200 = int(np.floor(200 / (sr / 2.0) * self.enc_dim))
        bw500 = int(np.floor(500 / (sr / 2.0) * self.enc_dim))
        bw2k = int(np.floor(2000 / (sr / 2.0) * self.enc_dim))
        self.band_width = [bw100] * 15 + [bw200] * 10 + [bw500] * 5 + [bw2k]
        self.band_width.append(self.enc_dim - int(np.sum(self.band_width)))
        self.nband = len(self.band_width)
        self.spk_transform = nn.Identity()
        if joint_training:                                         # bsrnn.py:216-250
            from wesep_b200.modules.speaker.resnet import get_speaker_model
            if not spk_feat and feat_type != "consistent":
                raise NotImplementedError("spk_feat=False is built for feat_type='consistent' (bsrnn.py:231-241)")
            self.spk_model = get_speaker_model(spk_model)(**(spk_args or {}))
            if spk_model_init:
                pretrained = torch.load(spk_model_init, map_location="cpu")
                state = self.spk_model.state_dict()
                for key in state.keys():
                    if key in pretrained.keys():
                        state[key] = pretrained[key]
                    else:
                        print("not %s loaded" % key)
                self.spk_model.load_state_dict(state)
            if spk_model_freeze:
                for param in self.spk_model.parameters():
                    param.requires_grad = False
            if not spk_feat:                                       # bsrnn.py:231-241: features computed inside the model
                from wesep_b200.modules.speaker.consistent import MelSpectrogram, PreEmphasis
                self.preEmphasis = PreEmphasis()
                self.spk_encoder = MelSpectrogram(sample_rate=sr, n_fft=win, hop_length=stride, f_min=20.0,
                                                  n_mels=(spk_args or {})["feat_dim"])
            else:
                self.preEmphasis = nn.Identity()
                self.spk_encoder = nn.Identity()
            self.pred_linear = nn.Linear(spk_emb_dim, spksInTrain) if multi_task else nn.Identity()
        self.BN = nn.ModuleList([nn.Sequential(nn.GroupNorm(1, bw * 2, self.eps), nn.Conv1d(bw * 2, feature_dim, 1))
                                 for bw in self.band_width])
        self.separator = FuseSeparation(nband=self.nband, num_repeat=num_repeat, feature_dim=feature_dim,
                                        spk_emb_dim=spk_emb_dim, spk_fuse_type=spk_fuse_type, multi_fuse=multi_fuse)
        self.mask = nn.ModuleList([
            nn.Sequential(nn.GroupNorm(1, feature_dim, self.eps), nn.Conv1d(feature_dim, feature_dim * 4, 1), nn.Tanh(),
                          nn.Conv1d(feature_dim * 4, feature_dim * 4, 1), nn.Tanh(),
                          nn.Conv1d(feature_dim * 4, bw * 4, 1)) for bw in self.band_width])

    # ---- constant DFT bases (band-major spectrum rows: per band bw re rows then bw im rows; padded to a multiple of 4) ----
    def _bases(self, device):
        key = str(device)
        cache = self.__dict__.setdefault("_basis_cache", {})
        if key not in cache:
            win, F = self.win, self.enc_dim
            w = torch.hann_window(win, dtype=torch.float32).double()      # fp32-rounded window, bsrnn.py:313
            k = torch.arange(win, dtype=torch.float64)
            f = torch.arange(F, dtype=torch.float64)
            ang = 2.0 * math.pi * f[:, None] * k[None, :] / win
            rows = sum(2 * b for b in self.band_width)
            R = (rows + 3) // 4 * 4
            fwd = torch.zeros(R, win, dtype=torch.float64)                # spec = fwd @ frame
            inv = torch.zeros(win, R, dtype=torch.float64)                # frame = inv @ spec  (includes the window)
            wgt = torch.full((F,), 2.0, dtype=torch.float64)
            wgt[0] = 1.0
            wgt[-1] = 1.0
            cr, ci = torch.cos(ang), -torch.sin(ang)
            icr = wgt[:, None] * torch.cos(ang) / win
            ici = -wgt[:, None] * torch.sin(ang) / win
            ici[0] = 0.0
            ici[-1] = 0.0
            offs, lo, r = [], 0, 0
            for bw in self.band_width:
                offs.append(r)
                fwd[r:r + bw] = cr[lo:lo + bw] * w
                fwd[r + bw:r + 2 * bw] = ci[lo:lo + bw] * w
                inv[:, r:r + bw] = (icr[lo:lo + bw] * w).t()
                inv[:, r + bw:r + 2 * bw] = (ici[lo:lo + bw] * w).t()
                lo += bw
                r += 2 * bw
            cache[key] = (fwd.float().to(device).contiguous(), inv.float().to(device).contiguous(), offs, R,
                          (w * w).float().to(device))
        return cache[key]

    def _resrnn_args(self, m):
        r = m.rnn
        return (m.norm.weight, m.norm.bias,
                [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, r.weight_ih_l0_reverse, r.weight_hh_l0_reverse,
                 r.bias_ih_l0_reverse, r.bias_hh_l0_reverse], m.proj.weight, m.proj.bias)

    def _fuse(self, layer, x, emb):
        N, nb = self.feature_dim, self.nband
        W, b = layer.fc.linear.weight, layer.fc.linear.bias
        if layer.fuse_type == "concat":
            # speaker.py:95-101 (4-D branch): Linear(N + E -> N) on cat([x, embed.expand], channel) at every (band, frame)
            # = W[:, :N] x (one pointwise conv over the N channels of every band) + (W[:, N:] embed + bias), a per-row
            # vector computed once and shared by all bands and frames
            B, NN, T = x.shape
            v = ops.LinearFn.apply(emb, W[:, N:], b)                                      # [B, N]
            rb = v.repeat_interleave(nb, 0)                                               # rows (b, band) of the band view
            xs = ops.as_act(x)
            ld = xs.stride(1)
            xb = xs.as_strided((B * nb, N, T), (N * ld, ld, 1))                           # bands as rows, same memory
            y = ops.Conv1x1RowBiasFn.apply(xb, W[:, :N], rb)
            ld2 = y.stride(1)
            return y.as_strided((B, NN, T), (NN * ld2, ld2, 1))
        v = ops.LinearFn.apply(emb, W, b)                                                 # [B, N], once per row
        v = v.repeat(1, nb)                                                               # same vector for every band
        if layer.fuse_type == "multiply":
            return ops.RowAffineFn.apply(x, v, None)
        return ops.RowAffineFn.apply(x, None, v)

    def _check(self, input, embeddings):
        if input.dim() != 2:
            raise RuntimeError("BSRNN expects [batch, samples]")
        if not (input.is_cuda and embeddings.is_cuda):
            raise RuntimeError("wesep_b200 kernels need CUDA tensors (no CPU fallback)")

    def _analysis(self, input):
        """STFT + band split (bsrnn.py:307-338): returns (spec [B, R, T] band-major re | im rows, features [B, nb*N, T])."""
        dev = input.device
        B, L = input.shape
        win, hop, N, nb = self.win, self.stride, self.feature_dim, self.nband
        fwd_b, inv_b, offs, R, w2 = self._bases(dev)
        T = 1 + L // hop
        with torch.no_grad():                                   # the mixture is data: no gradient through the analysis
            x = input.float()
            pad = win // 2
            xp = torch.cat([x[:, 1:pad + 1].flip(1), x, x[:, L - pad - 1:L - 1].flip(1)], 1).contiguous()
            spec = ops.conv1x1_raw(ops.frames_raw(xp, win, T, hop), fwd_b, False, R)      # [B, R, T] band-major re | im
        # band split: GroupNorm(1, 2bw) per band, then all 32 Conv1d(2bw, N, 1) as one block-diagonal GEMM
        parts = [ops.group_norm1(spec[:, o:o + 2 * bw], self.BN[i][0].weight, self.BN[i][0].bias)
                 for i, (o, bw) in enumerate(zip(offs, self.band_width))]
        if R > offs[-1] + 2 * self.band_width[-1]:
            parts.append(torch.zeros(B, R - offs[-1] - 2 * self.band_width[-1], T, device=dev))
        xhat = torch.cat(parts, 1)
        Wbig = torch.zeros(nb * N, R, device=dev)
        for i, (o, bw) in enumerate(zip(offs, self.band_width)):
            Wbig[i * N:(i + 1) * N, o:o + 2 * bw] = self.BN[i][1].weight[:, :, 0]
        bbig = torch.cat([self.BN[i][1].bias for i in range(nb)])
        return spec, ops.Conv1x1Fn.apply(xhat, Wbig, bbig, False, None)                   # [B, nb*N, T]

    def _speaker(self, spk_in, raw_wave):
        """bsrnn.py:340-360: (embedding [B, E], predict_speaker_lable).  `raw_wave`: compute the "consistent" features first."""
        predict_speaker_lable = torch.zeros((), device=spk_in.device)   # dummy, bsrnn.py:340-341 (a fill kernel: graph-capturable)
        if self.joint_training:                                    # bsrnn.py:342-357
            if raw_wave:
                from wesep_b200.modules.speaker.consistent import consistent_features
                spk_in = consistent_features(spk_in, self.preEmphasis, self.spk_encoder)
            tmp = self.spk_model(spk_in)
            spk_in = tmp[-1] if isinstance(tmp, tuple) else tmp
            if self.multi_task:
                predict_speaker_lable = ops.LinearFn.apply(spk_in, self.pred_linear.weight, self.pred_linear.bias)
            else:
                predict_speaker_lable = spk_in                      # nn.Identity
        return self.spk_transform(spk_in).float(), predict_speaker_lable

    def _separate(self, x, spec, emb, L):
        """separator -> mask heads -> iSTFT (bsrnn.py:362-391): features [B, nb*N, T] + embedding -> estimate [B, L]."""
        dev = x.device
        B = x.shape[0]
        win, hop, N, nb = self.win, self.stride, self.feature_dim, self.nband
        fwd_b, inv_b, offs, R, w2 = self._bases(dev)
        T = x.shape[2]
        sep = self.separator.separation
        if self.separator.multi_fuse:
            for r in range(len(sep) // 2):
                x = self._fuse(sep[2 * r], x, emb)
                x = ops.bsnet(x, nb, self._resrnn_args(sep[2 * r + 1].band_rnn), self._resrnn_args(sep[2 * r + 1].band_comm))
        else:
            x = self._fuse(sep[0], x, emb)
            for r in range(1, len(sep)):
                x = ops.bsnet(x, nb, self._resrnn_args(sep[r].band_rnn), self._resrnn_args(sep[r].band_comm))
        # mask head per band -> estimate spectrum (band-major)
        ests = []
        for i, (o, bw) in enumerate(zip(offs, self.band_width)):
            mk = self.mask[i]
            y = ops.group_norm1(x[:, i * N:(i + 1) * N], mk[0].weight, mk[0].bias)
            y = ops.TanhFn.apply(ops.Conv1x1Fn.apply(y, mk[1].weight[:, :, 0], mk[1].bias, False, None))
            y = ops.TanhFn.apply(ops.Conv1x1Fn.apply(y, mk[3].weight[:, :, 0], mk[3].bias, False, None))
            y = ops.Conv1x1Fn.apply(y, mk[5].weight[:, :, 0], mk[5].bias, False, None)   # [B, 4bw, T]
            ests.append(ops.MaskApplyFn.apply(y, spec[:, o:o + 2 * bw]))
        if R > offs[-1] + 2 * self.band_width[-1]:
            ests.append(torch.zeros(B, R - offs[-1] - 2 * self.band_width[-1], T, device=dev))
        est = torch.cat(ests, 1)                                                          # [B, R, T]
        frames = ops.FixedGemmFn.apply(est, inv_b, False)                                 # [B, win, T]
        n_out = win + hop * (T - 1)
        y = ops.OverlapAddFn.apply(frames, hop, n_out)                                    # [B, n_out]
        env = self.__dict__.setdefault("_env_cache", {}).get((str(dev), T, L))
        if env is None:
            e = torch.zeros(n_out, device=dev)
            for t in range(T):
                e[t * hop:t * hop + win] += w2
            env = self._env_cache[(str(dev), T, L)] = 1.0 / e[win // 2:win // 2 + L]
        return ops.ColVecMulFn.apply(y[:, win // 2:win // 2 + L], env)

    def forward(self, input, embeddings):
        self._check(input, embeddings)
        spec, x = self._analysis(input)
        emb, predict_speaker_lable = self._speaker(embeddings, raw_wave=self.joint_training and not self.spk_feat)
        return self._separate(x, spec, emb, input.shape[1]), predict_speaker_lable


class BSRNN_Multi(BSRNN):
    """wesep/models/bsrnn_multi_optim.py:155-472: pBSRNN with the self-enrollment second pass — while gradients are enabled
    the first estimate (detached) is turned into "consistent" enrollment features, embedded by the same speaker encoder and
    the separator runs again on the same band features; returns (s, self_s, predict, self_predict), and (s, predict) under
    no_grad.  Parameters and state_dict are those of BSRNN (bsrnn_multi_optim.yaml: joint_training, spk_feat False)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if not self.joint_training or self.spk_feat or self.feat_type != "consistent":
            raise NotImplementedError("BSRNN_Multi: joint_training=True, spk_feat=False, feat_type='consistent' (the recipe, "
                                      "bsrnn_multi_optim.yaml:52-68) is built; the reference's own second pass needs them too "
                                      "(bsrnn_multi_optim.py:409-431)")

    def forward(self, input, embeddings):
        self._check(input, embeddings)
        L = input.shape[1]
        spec, x = self._analysis(input)
        emb, predict_speaker_lable = self._speaker(embeddings, raw_wave=True)
        s = self._separate(x, spec, emb, L)
        if not torch.is_grad_enabled():
            return s, predict_speaker_lable
        emb2, self_predict_speaker_lable = self._speaker(s.detach(), raw_wave=True)         # bsrnn_multi_optim.py:406-431
        self_s = self._separate(x, spec, emb2, L)
        return s, self_s, predict_speaker_lable, self_predict_speaker_lable
