"""TF-GridNet — reference wesep/models/tfgridnet.py:74-312 + wesep/modules/tfgridnet/gridnet_block.py:26-284
(BASELINE config 5).

Same constructor keywords, attribute paths and parameter shapes as the reference (state_dict keys match); the forward
runs on libwesep_b200: STFT / iSTFT as framing + windowed-DFT GEMMs, the 3x3 Conv2d / ConvTranspose2d as im2col / col2im
around the tcgen05 pointwise GEMM, GroupNorm, the per-frequency speaker gain, and per GridNetBlock: LayerNorm over
channels (cLN kernel) -> BLSTM on the persistent cluster recurrence kernel -> Linear, along frequency then along time;
1x1 convolutions, PReLU + head LayerNorm (csrc/tfgridnet.cu), both attention products on the pointwise GEMMs with the row
softmax kernel in between.  The nn.* members only hold parameters.  There is no CPU path.

Built: every (emb_ks, emb_hs) the reference accepts — 1 / 1 (tfgridnet.yaml:52-53), ks != hs (the class default 4 / 1: unfold +
ConvTranspose1d) and ks == hs > 1 (ks positions packed per step) —,
n_srcs == 1, n_imics == 1, multiply fusion.
"""
import math

import torch
import torch.nn as nn

from wesep_b200 import ops


class _FuseFC(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features)


class SpeakerFuseLayer(nn.Module):
    def __init__(self, embed_dim=256, feat_dim=65, fuse_type="multiply"):
        super().__init__()
        if fuse_type != "multiply":
            raise NotImplementedError("TF-GridNet: only spk_fuse_type='multiply' (tfgridnet.yaml:57) is built")
        self.fuse_type = fuse_type
        self.fc = _FuseFC(embed_dim, feat_dim)


class LayerNormalization4DCF(nn.Module):
    """gridnet_block.py:229-252 (parameters only)."""

    def __init__(self, input_dimension, eps=1e-5):
        super().__init__()
        assert len(input_dimension) == 2
        self.gamma = nn.Parameter(torch.ones(1, input_dimension[0], 1, input_dimension[1]))
        self.beta = nn.Parameter(torch.zeros(1, input_dimension[0], 1, input_dimension[1]))
        self.eps = eps


class AllHeadPReLULayerNormalization4DCF(nn.Module):
    """gridnet_block.py:255-284 (parameters only)."""

    def __init__(self, input_dimension, eps=1e-5):
        super().__init__()
        assert len(input_dimension) == 3
        H, E, n_freqs = input_dimension
        self.gamma = nn.Parameter(torch.ones(1, H, E, 1, n_freqs))
        self.beta = nn.Parameter(torch.zeros(1, H, E, 1, n_freqs))
        self.act = nn.PReLU(num_parameters=H, init=0.25)
        self.eps, self.H, self.E, self.n_freqs = eps, H, E, n_freqs


def _lstm_args(r):
    return [r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, r.weight_ih_l0_reverse, r.weight_hh_l0_reverse,
            r.bias_ih_l0_reverse, r.bias_hh_l0_reverse]


class GridNetBlock(nn.Module):
    """gridnet_block.py:26-227."""

    def __getitem__(self, key):
        return getattr(self, key)

    def __init__(self, emb_dim, emb_ks, emb_hs, n_freqs, hidden_channels, n_head=4, approx_qk_dim=512, activation="prelu",
                 eps=1e-5):
        super().__init__()
        assert activation == "prelu"
        in_channels = emb_dim * emb_ks
        self.intra_norm = nn.LayerNorm(emb_dim, eps=eps)
        self.intra_rnn = nn.LSTM(in_channels, hidden_channels, 1, batch_first=True, bidirectional=True)
        if emb_ks == emb_hs:
            self.intra_linear = nn.Linear(hidden_channels * 2, in_channels)
        else:
            self.intra_linear = nn.ConvTranspose1d(hidden_channels * 2, emb_dim, emb_ks, stride=emb_hs)
        self.inter_norm = nn.LayerNorm(emb_dim, eps=eps)
        self.inter_rnn = nn.LSTM(in_channels, hidden_channels, 1, batch_first=True, bidirectional=True)
        if emb_ks == emb_hs:
            self.inter_linear = nn.Linear(hidden_channels * 2, in_channels)
        else:
            self.inter_linear = nn.ConvTranspose1d(hidden_channels * 2, emb_dim, emb_ks, stride=emb_hs)
        E = math.ceil(approx_qk_dim * 1.0 / n_freqs)
        assert emb_dim % n_head == 0
        self.add_module("attn_conv_Q", nn.Conv2d(emb_dim, n_head * E, 1))
        self.add_module("attn_norm_Q", AllHeadPReLULayerNormalization4DCF((n_head, E, n_freqs), eps=eps))
        self.add_module("attn_conv_K", nn.Conv2d(emb_dim, n_head * E, 1))
        self.add_module("attn_norm_K", AllHeadPReLULayerNormalization4DCF((n_head, E, n_freqs), eps=eps))
        self.add_module("attn_conv_V", nn.Conv2d(emb_dim, n_head * emb_dim // n_head, 1))
        self.add_module("attn_norm_V", AllHeadPReLULayerNormalization4DCF((n_head, emb_dim // n_head, n_freqs), eps=eps))
        self.add_module("attn_concat_proj", nn.Sequential(nn.Conv2d(emb_dim, emb_dim, 1), nn.PReLU(),
                                                          LayerNormalization4DCF((emb_dim, n_freqs), eps=eps)))
        self.emb_dim, self.emb_ks, self.emb_hs, self.n_head, self.E, self.eps = emb_dim, emb_ks, emb_hs, n_head, E, eps

    def _qkv(self, x, conv, norm, T, F):
        y = ops.Conv1x1Fn.apply(x, conv.weight.reshape(conv.weight.shape[0], -1), conv.bias, False, None)
        return ops.HeadLnFn.apply(y, norm.act.weight, norm.gamma, norm.beta, self.n_head, T, F, norm.eps)

    def _path(self, x, norm, rnn, lin):
        ks, hs, C = self.emb_ks, self.emb_hs, self.emb_dim
        if ks == 1:
            return ops.res_rnn(x, norm.weight, norm.bias, _lstm_args(rnn), lin.weight, lin.bias, layer_norm_eps=self.eps)
        if ks != hs:
            return ops.res_rnn_unfold(x, norm.weight, norm.bias, _lstm_args(rnn), lin.weight, lin.bias, ks, hs, self.eps)
        # ks == hs > 1 (gridnet_block.py:139-146): `view([B*T, -1, ks*C])` packs ks consecutive positions into the feature axis in
        # (k, c) order; the unfold kernel produces (c, k) rows, so the LSTM input columns / Linear output rows are permuted to match
        perm = (torch.arange(ks, device=x.device)[None, :] * C + torch.arange(C, device=x.device)[:, None]).reshape(-1)   # [c*ks+k] -> k*C+c
        lstm = _lstm_args(rnn)
        lstm[0], lstm[4] = lstm[0][:, perm], lstm[4][:, perm]
        rows, _, S = x.shape
        xh = ops.cln(x, norm.weight, norm.bias, self.eps)
        xn = ops.SwapOIFn.apply(ops.Unfold1dFn.apply(xh, ks, ks), 1, None)                # [S/ks, C*ks, rows]
        h = ops.LstmTmFn.apply(xn, *lstm)
        p = ops.Conv1x1Fn.apply(h, lin.weight[perm], lin.bias[perm], False, None)        # [S/ks, C*ks, rows]
        y = ops.Fold1dFn.apply(ops.SwapOIFn.apply(p, 1, None), C, S, ks, ks)
        return ops.AddFn.apply(y, x)

    def run(self, x, B, T, F):
        """x act [B, C, T*F] -> act [B, C, T*F]."""
        C, H = self.emb_dim, self.n_head
        old_T, old_F = T, F
        olp = self.emb_ks - self.emb_hs
        x4 = x.unflatten(2, (T, F))
        if self.emb_ks != 1:                                       # gridnet_block.py:124-133: zero padding to whole windows
            T = math.ceil((old_T + 2 * olp - self.emb_ks) / self.emb_hs) * self.emb_hs + self.emb_ks
            F = math.ceil((old_F + 2 * olp - self.emb_ks) / self.emb_hs) * self.emb_hs + self.emb_ks
            x4 = torch.nn.functional.pad(x4, (olp, F - old_F - olp, olp, T - old_T - olp))
        # intra (along frequency): rows (b, t), steps f — gridnet_block.py:135-161
        xi = ops.as_act(x4.permute(0, 2, 1, 3).reshape(B * T, C, F))
        yi = self._path(xi, self.intra_norm, self.intra_rnn, self.intra_linear)
        # inter (along time): rows (b, f), steps t — gridnet_block.py:163-187
        xe = ops.SwapOIFn.apply(yi, B, None)                                             # [B*F, C, T]
        ye = self._path(xe, self.inter_norm, self.inter_rnn, self.inter_linear)
        x1 = ye.unflatten(0, (B, F)).permute(0, 2, 3, 1)                                 # [B, F, C, T] -> [B, C, T, F]
        if self.emb_ks != 1:
            x1 = x1[..., olp:olp + old_T, olp:olp + old_F]                               # gridnet_block.py:190
            T, F = old_T, old_F
        x1 = ops.as_act(x1.reshape(B, C, T * F))
        # full-band self-attention over frames — gridnet_block.py:192-224
        Qc = self._qkv(x1, self["attn_conv_Q"], self["attn_norm_Q"], T, F)               # [B, H*E, T*F]
        Kc = self._qkv(x1, self["attn_conv_K"], self["attn_norm_K"], T, F)
        Vc = self._qkv(x1, self["attn_conv_V"], self["attn_norm_V"], T, F)               # [B, C, T*F]
        E, Ev = self.E, C // H
        heads = []
        for b in range(B):
            for h in range(H):
                Qm = Qc[b, h * E:(h + 1) * E].unflatten(1, (T, F)).permute(1, 0, 2).reshape(T, E * F)
                Kt = Kc[b, h * E:(h + 1) * E].unflatten(1, (T, F)).permute(0, 2, 1).reshape(E * F, T)
                Vm = Vc[b, h * Ev:(h + 1) * Ev].unflatten(1, (T, F)).permute(1, 0, 2).reshape(T, Ev * F)
                o = ops.attention_rows(Qm, Kt, Vm)                                       # [T, Ev*F]
                heads.append(o.unflatten(1, (Ev, F)).permute(1, 0, 2))                   # [Ev, T, F]
        att = ops.as_act(torch.stack(heads).reshape(B, C, T * F))
        proj = self["attn_concat_proj"]
        y = ops.Conv1x1Fn.apply(att, proj[0].weight.reshape(C, C), proj[0].bias, False, None)
        y = ops.HeadLnFn.apply(y, proj[1].weight, proj[2].gamma, proj[2].beta, 1, T, F, proj[2].eps)
        return ops.AddFn.apply(y, x1)


class TFGridNet(nn.Module):

    def __init__(
        self,
        n_srcs=1,
        sr=16000,
        n_fft=128,
        stride=64,
        window="hann",
        n_imics=1,
        n_layers=6,
        lstm_hidden_units=192,
        attn_n_head=4,
        attn_approx_qk_dim=512,
        emb_dim=48,
        emb_ks=4,
        emb_hs=1,
        activation="prelu",
        eps=1.0e-5,
        spk_emb_dim=256,
        use_spk_transform=False,
        spk_fuse_type="multiply",
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
        if n_srcs != 1 or n_imics != 1:
            raise NotImplementedError("TF-GridNet: n_srcs == n_imics == 1 (tfgridnet.yaml:44,48) is built")
        if window != "hann":
            raise NotImplementedError("TF-GridNet: the hann window of the recipe is built")
        if use_spk_transform:
            raise NotImplementedError("use_spk_transform=True is not on the recipe path (tfgridnet.yaml:56)")
        if not ops.lstm_rec_supported(lstm_hidden_units) and lstm_hidden_units % 4:
            raise NotImplementedError("TF-GridNet: lstm_hidden_units must be a multiple of 4")
        self.n_srcs, self.n_fft, self.stride, self.window, self.n_imics, self.n_layers = n_srcs, n_fft, stride, window, n_imics, n_layers
        self.spk_emb_dim, self.joint_training, self.spk_feat, self.feat_type = spk_emb_dim, joint_training, spk_feat, feat_type
        self.spk_model_freeze, self.multi_task, self.eps = spk_model_freeze, multi_task, eps
        assert n_fft % 2 == 0
        n_freqs = n_fft // 2 + 1
        self.spk_transform = nn.Identity()
        if joint_training:                                         # tfgridnet.py:127-163
            from wesep_b200.modules.speaker.resnet import get_speaker_model
            if not spk_feat and feat_type != "consistent":
                raise NotImplementedError("spk_feat=False is built for feat_type='consistent' (tfgridnet.py:144-155)")
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
            if not spk_feat:                                       # tfgridnet.py:144-155
                from wesep_b200.modules.speaker.consistent import MelSpectrogram, PreEmphasis
                self.preEmphasis = PreEmphasis()
                self.spk_encoder = MelSpectrogram(sample_rate=sr, n_fft=n_fft, hop_length=stride, f_min=20.0,
                                                  n_mels=(spk_args or {})["feat_dim"])
            else:
                self.preEmphasis = nn.Identity()
                self.spk_encoder = nn.Identity()
            self.pred_linear = nn.Linear(spk_emb_dim, spksInTrain) if multi_task else nn.Identity()
        self.spk_fuse = SpeakerFuseLayer(embed_dim=spk_emb_dim, feat_dim=n_freqs, fuse_type=spk_fuse_type)
        t_ksize = 3
        ks, padding = (t_ksize, 3), (t_ksize // 2, 1)
        self.conv = nn.Sequential(nn.Conv2d(2 * n_imics, emb_dim, ks, padding=padding), nn.GroupNorm(1, emb_dim, eps=eps))
        self.blocks = nn.ModuleList([GridNetBlock(emb_dim, emb_ks, emb_hs, n_freqs, lstm_hidden_units, n_head=attn_n_head,
                                                  approx_qk_dim=attn_approx_qk_dim, activation=activation, eps=eps)
                                     for _ in range(n_layers)])
        self.deconv = nn.ConvTranspose2d(emb_dim, n_srcs * 2, ks, padding=padding)

    def _bases(self, device):
        key = str(device)
        cache = self.__dict__.setdefault("_basis_cache", {})
        if key not in cache:
            win, F = self.n_fft, self.n_fft // 2 + 1
            w = torch.hann_window(win, dtype=torch.float32).double()
            k = torch.arange(win, dtype=torch.float64)
            f = torch.arange(F, dtype=torch.float64)
            ang = 2.0 * math.pi * f[:, None] * k[None, :] / win
            R = (2 * F + 3) // 4 * 4
            fwd = torch.zeros(R, win, dtype=torch.float64)
            inv = torch.zeros(win, R, dtype=torch.float64)
            wgt = torch.full((F,), 2.0, dtype=torch.float64)
            wgt[0] = 1.0
            wgt[-1] = 1.0
            fwd[:F] = torch.cos(ang) * w
            fwd[F:2 * F] = -torch.sin(ang) * w
            ici = -wgt[:, None] * torch.sin(ang) / win
            ici[0] = 0.0
            ici[-1] = 0.0
            inv[:, :F] = (wgt[:, None] * torch.cos(ang) / win * w).t()
            inv[:, F:2 * F] = (ici * w).t()
            cache[key] = (fwd.float().to(device).contiguous(), inv.float().to(device).contiguous(), R, (w * w).float().to(device))
        return cache[key]

    @property
    def num_spk(self):
        return self.n_srcs

    def forward(self, input, embeddings):
        if input.dim() != 2:
            raise AssertionError("TFGridNet (n_imics == 1) expects [batch, samples]")      # tfgridnet.py:214 assert
        if not (input.is_cuda and embeddings.is_cuda):
            raise RuntimeError("wesep_b200 kernels need CUDA tensors (no CPU fallback)")
        dev = input.device
        B, L = input.shape
        win, hop = self.n_fft, self.stride
        Fq = win // 2 + 1
        fwd_b, inv_b, R, w2 = self._bases(dev)
        T = 1 + L // hop
        with torch.no_grad():                                   # the mixture is data: no gradient through the analysis
            x = input.float().contiguous()
            std, inv_std = ops.row_std(x)                       # tfgridnet.py:217-218 RMS normalisation
            x = ops.RowAffineFn.apply(x[:, None, :], inv_std[:, None], None)[:, 0]
            pad = win // 2
            xp = torch.cat([x[:, 1:pad + 1].flip(1), x, x[:, L - pad - 1:L - 1].flip(1)], 1).contiguous()
            spec = ops.conv1x1_raw(ops.frames_raw(xp, win, T, hop), fwd_b, False, R)      # [B, R, T]: re rows | im rows
            batch = ops.as_act(spec[:, :2 * Fq].reshape(B, 2, Fq, T).transpose(2, 3).reshape(B, 2, T * Fq))   # [B, 2, T, F]
        batch = ops.conv3x3(batch, T, Fq, self.conv[0].weight, self.conv[0].bias, 1)
        batch = ops.group_norm1(batch, self.conv[1].weight, self.conv[1].bias, self.conv[1].eps)

        predict_speaker_lable = torch.zeros((), device=dev)          # dummy, tfgridnet.py:249-250
        spk_in = embeddings
        if self.joint_training:                                    # tfgridnet.py:251-269
            if not self.spk_feat:
                from wesep_b200.modules.speaker.consistent import consistent_features
                spk_in = consistent_features(spk_in, self.preEmphasis, self.spk_encoder)
            tmp = self.spk_model(spk_in)
            spk_in = tmp[-1] if isinstance(tmp, tuple) else tmp
            if self.multi_task:
                predict_speaker_lable = ops.LinearFn.apply(spk_in, self.pred_linear.weight, self.pred_linear.bias)
            else:
                predict_speaker_lable = spk_in
        emb = self.spk_transform(spk_in).float()
        gain = ops.LinearFn.apply(emb, self.spk_fuse.fc.linear.weight, self.spk_fuse.fc.linear.bias)       # [B, F]
        for blk in self.blocks:                                    # tfgridnet.py:274-278: the same fuse layer before every block
            batch = ops.ColScaleFn.apply(batch, gain, T, Fq)
            batch = blk.run(batch, B, T, Fq)
        out = ops.conv_transpose3x3(batch, T, Fq, self.deconv.weight, self.deconv.bias, 1)                 # [B, 2, T*F]
        est = out.reshape(B, 2, T, Fq).transpose(2, 3).reshape(B, 2 * Fq, T)
        if R > 2 * Fq:
            est = torch.cat([est, torch.zeros(B, R - 2 * Fq, T, device=dev)], 1)
        frames = ops.FixedGemmFn.apply(ops.as_act(est), inv_b, False)                       # [B, win, T]
        n_out = win + hop * (T - 1)
        y = ops.OverlapAddFn.apply(frames, hop, n_out)
        env = self.__dict__.setdefault("_env_cache", {}).get((str(dev), T, L))
        if env is None:
            e = torch.zeros(n_out, device=dev)
            for t in range(T):
                e[t * hop:t * hop + win] += w2
            env = self._env_cache[(str(dev), T, L)] = 1.0 / e[win // 2:win // 2 + L]
        s = ops.ColVecMulFn.apply(y[:, win // 2:win // 2 + L], env)
        s = ops.RowAffineFn.apply(s[:, None, :], std[:, None], None)[:, 0]                   # tfgridnet.py:297
        return s, predict_speaker_lable
