"""pDPCCN — reference wesep/models/dpccn.py:16-290 + wesep/modules/dpccn/convs.py:28-152.

Same constructor keywords, attribute paths and parameter shapes as the reference (state_dict keys match), but the
forward runs on libwesep_b200: STFT / iSTFT as framing + windowed-DFT GEMMs (as pBSRNN), every 3x3 Conv2d /
ConvTranspose2d as im2col / col2im around the tcgen05 pointwise GEMM, ELU + InstanceNorm, the depthwise dilated Conv1d,
AvgPool2d / bilinear Upsample and the per-frequency speaker gain as streaming kernels (csrc/dpccn.cu).  The nn.Conv2d /
nn.ConvTranspose2d / nn.Conv1d members only hold parameters; their own forward is never called.  There is no CPU path.

Feature maps keep the reference's NCHW order flattened: act tensors [B, C, T*F] with F contiguous.
"""
import math

import torch
import torch.nn as nn

from wesep_b200 import ops


class _FuseFC(nn.Module):
    """LinearLayer of common/speaker.py:47-60 (attribute path `fc.linear`)."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.linear = nn.Linear(in_features, out_features)


class SpeakerFuseLayer(nn.Module):
    def __init__(self, embed_dim=256, feat_dim=257, fuse_type="multiply"):
        super().__init__()
        if fuse_type != "multiply":
            raise NotImplementedError("pDPCCN: only spk_fuse_type='multiply' (dpccn.yaml:49) is built")
        self.fuse_type = fuse_type
        self.fc = _FuseFC(embed_dim, feat_dim)


class Conv2dBlock(nn.Module):
    """convs.py:28-47: Conv2d -> ELU -> InstanceNorm2d."""

    def __init__(self, in_dims=16, out_dims=32, kernel_size=(3, 3), stride=(1, 1), padding=(1, 1)):
        super().__init__()
        if tuple(kernel_size) != (3, 3) or tuple(padding) != (1, 1) or stride[0] != 1 or stride[1] not in (1, 2):
            raise NotImplementedError("pDPCCN: 3x3 convolutions with padding (1, 1) and stride (1, 1|2) are built")
        self.conv2d = nn.Conv2d(in_dims, out_dims, kernel_size, stride, padding)
        self.elu = nn.ELU()
        self.norm = nn.InstanceNorm2d(out_dims)
        self.stride = tuple(stride)

    def run(self, x, T, F):
        """x act [B, Ci, T*F] -> (y [B, Co, T*Fo], Fo)."""
        y = ops.conv3x3(x, T, F, self.conv2d.weight, self.conv2d.bias, self.stride)
        return ops.EluInFn.apply(y, 0), (F - 1) // self.stride[1] + 1


class ConvTrans2dBlock(nn.Module):
    """convs.py:50-69: ConvTranspose2d -> ELU -> InstanceNorm2d."""

    def __init__(self, in_dims=32, out_dims=16, kernel_size=(3, 3), stride=(1, 2), padding=(1, 0), output_padding=(0, 0)):
        super().__init__()
        if (tuple(kernel_size) != (3, 3) or tuple(padding) != (1, 1) or tuple(output_padding) != (0, 0) or stride[0] != 1
                or stride[1] not in (1, 2)):
            raise NotImplementedError("pDPCCN: 3x3 transposed convolutions with padding (1, 1), stride (1, 1|2) are built")
        self.convtrans2d = nn.ConvTranspose2d(in_dims, out_dims, kernel_size, stride, padding, output_padding)
        self.elu = nn.ELU()
        self.norm = nn.InstanceNorm2d(out_dims)
        self.stride = tuple(stride)

    def run(self, x, T, F):
        Fo = (F - 1) * self.stride[1] + 1                     # (F - 1) s - 2 p + k
        y = ops.conv_transpose3x3(x, T, Fo, self.convtrans2d.weight, self.convtrans2d.bias, self.stride)
        return ops.EluInFn.apply(y, 0), Fo


class DenseBlock(nn.Module):
    """convs.py:72-106."""

    def __init__(self, in_dims, out_dims, mode="enc", **kargs):
        super().__init__()
        if mode not in ["enc", "dec"]:
            raise RuntimeError("The mode option must be 'enc' or 'dec'!")
        n = 1 if mode == "enc" else 2
        self.conv1 = Conv2dBlock(in_dims=in_dims * n, out_dims=in_dims, **kargs)
        self.conv2 = Conv2dBlock(in_dims=in_dims * (n + 1), out_dims=in_dims, **kargs)
        self.conv3 = Conv2dBlock(in_dims=in_dims * (n + 2), out_dims=in_dims, **kargs)
        self.conv4 = Conv2dBlock(in_dims=in_dims * (n + 3), out_dims=in_dims, **kargs)
        self.conv5 = Conv2dBlock(in_dims=in_dims * (n + 4), out_dims=out_dims, **kargs)

    def run(self, x, T, F):
        convs = (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5)
        if any(c.stride != (1, 1) for c in convs):
            raise NotImplementedError("pDPCCN: dense blocks use stride (1, 1) (dpccn.py:140,172)")
        wb = [t for c in convs for t in (c.conv2d.weight, c.conv2d.bias)]
        return ops.DenseBlockFn.apply(x, T, F, *wb), F


class TCNBlock(nn.Module):
    """convs.py:109-152: IN - ELU - depthwise dilated Conv1d - IN - ELU - Conv1d(1x1), residual."""

    def __init__(self, in_dims=384, out_dims=384, kernel_size=3, dilation=1, causal=False):
        super().__init__()
        if causal or kernel_size != 3 or in_dims != out_dims:
            raise NotImplementedError("pDPCCN: the non-causal k=3 TCN block of the recipe is built")
        self.norm1 = nn.InstanceNorm1d(in_dims)
        self.elu1 = nn.ELU()
        pad = (dilation * (kernel_size - 1)) // 2
        self.dconv1 = nn.Conv1d(in_dims, out_dims, kernel_size, padding=pad, dilation=dilation, groups=in_dims, bias=True)
        self.norm2 = nn.InstanceNorm1d(in_dims)
        self.elu2 = nn.ELU()
        self.dconv2 = nn.Conv1d(in_dims, out_dims, 1, bias=True)
        self.causal, self.dconv_pad, self.dilation = causal, pad, dilation

    def run(self, x):
        y = ops.EluInFn.apply(x, 1)
        y = ops.DwConv1dFn.apply(y, self.dconv1.weight, self.dconv1.bias, self.dilation)
        y = ops.EluInFn.apply(y, 1)
        y = ops.Conv1x1Fn.apply(y, self.dconv2.weight[:, :, 0], self.dconv2.bias, False, None)
        return ops.AddFn.apply(x, y)


class DPCCN(nn.Module):

    def __init__(
        self,
        win=512,
        stride=128,
        spk_emb_dim=256,
        sr=16000,
        use_spk_transform=False,
        spk_fuse_type="multiply",
        feature_dim=257,
        kernel_size=(3, 3),
        stride1=(1, 1),
        stride2=(1, 2),
        paddings=(1, 1),
        output_padding=(0, 0),
        tcn_dims=384,
        tcn_blocks=10,
        tcn_layers=2,
        causal=False,
        pool_size=(4, 8, 16, 32),
        multi_fuse=False,
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
            raise NotImplementedError("use_spk_transform=True is not on the recipe path (dpccn.yaml:48)")
        if tuple(stride1) != (1, 1) or tuple(kernel_size) != (3, 3):
            raise NotImplementedError("pDPCCN: kernel (3, 3) and stride1 (1, 1) are built")
        self.win_len, self.hop_size = win, stride
        self.spk_emb_dim = spk_emb_dim
        self.joint_training, self.spk_feat, self.feat_type = joint_training, spk_feat, feat_type
        self.spk_model_freeze, self.multi_task = spk_model_freeze, multi_task
        self.feature_dim = feature_dim
        self.pool_size = tuple(pool_size)

        self.conv2d = nn.Conv2d(2, 16, kernel_size, stride1, paddings)
        self.encoder = self._build_encoder(kernel_size=kernel_size, stride=stride2, padding=paddings)
        self.spk_transform = nn.Identity()
        if joint_training:                                         # dpccn.py:70-105
            from wesep_b200.modules.speaker.resnet import get_speaker_model
            if not spk_feat and feat_type != "consistent":
                raise NotImplementedError("spk_feat=False is built for feat_type='consistent' (dpccn.py:88-99)")
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
            if not spk_feat:                                       # dpccn.py:88-99
                from wesep_b200.modules.speaker.consistent import MelSpectrogram, PreEmphasis
                self.preEmphasis = PreEmphasis()
                self.spk_encoder = MelSpectrogram(sample_rate=sr, n_fft=win, hop_length=stride, f_min=20.0,
                                                  n_mels=(spk_args or {})["feat_dim"])
            else:
                self.preEmphasis = nn.Identity()
                self.spk_encoder = nn.Identity()
            self.pred_linear = nn.Linear(spk_emb_dim, spksInTrain) if multi_task else nn.Identity()
        self.spk_fuse = SpeakerFuseLayer(embed_dim=spk_emb_dim, feat_dim=feature_dim, fuse_type=spk_fuse_type)
        self.tcn_layers = nn.Sequential(*[
            nn.Sequential(*[TCNBlock(in_dims=tcn_dims, out_dims=tcn_dims, causal=causal, dilation=2 ** b)
                            for b in range(tcn_blocks)]) for _ in range(tcn_layers)])
        self.decoder = self._build_decoder(kernel_size=kernel_size, stride=stride2, padding=paddings,
                                           output_padding=output_padding)
        self.avg_pool = nn.ModuleList([nn.Sequential(nn.AvgPool2d(sz), nn.Conv2d(32, 8, 1, 1)) for sz in pool_size])
        self.avg_proj = nn.Conv2d(64, 32, 1, 1)
        self.deconv2d = nn.ConvTranspose2d(32, 2, kernel_size, stride1, paddings)

    def _build_encoder(self, **kw):
        enc = nn.ModuleList()
        enc.append(DenseBlock(16, 16, "enc"))
        for i in range(4):
            enc.append(nn.Sequential(Conv2dBlock(in_dims=16 if i == 0 else 32, out_dims=32, **kw), DenseBlock(32, 32, "enc")))
        enc.append(Conv2dBlock(in_dims=32, out_dims=64, **kw))
        enc.append(Conv2dBlock(in_dims=64, out_dims=128, **kw))
        enc.append(Conv2dBlock(in_dims=128, out_dims=384, **kw))
        return enc

    def _build_decoder(self, **kw):
        dec = nn.ModuleList()
        dec.append(ConvTrans2dBlock(in_dims=384 * 2, out_dims=128, **kw))
        dec.append(ConvTrans2dBlock(in_dims=128 * 2, out_dims=64, **kw))
        dec.append(ConvTrans2dBlock(in_dims=64 * 2, out_dims=32, **kw))
        for i in range(4):
            dec.append(nn.Sequential(DenseBlock(32, 64, "dec"), ConvTrans2dBlock(in_dims=64, out_dims=32 if i != 3 else 16, **kw)))
        dec.append(DenseBlock(16, 32, "dec"))
        return dec

    # ---- constant DFT bases: rows [0, F) real parts, [F, 2F) imaginary parts, padded to a multiple of 4 ----
    def _bases(self, device):
        key = str(device)
        cache = self.__dict__.setdefault("_basis_cache", {})
        if key not in cache:
            win, F = self.win_len, self.win_len // 2 + 1
            w = torch.hann_window(win, dtype=torch.float32).double()          # fp32-rounded window, dpccn.py:216
            k = torch.arange(win, dtype=torch.float64)
            f = torch.arange(F, dtype=torch.float64)
            ang = 2.0 * math.pi * f[:, None] * k[None, :] / win
            R = (2 * F + 3) // 4 * 4
            fwd = torch.zeros(R, win, dtype=torch.float64)                    # spec = fwd @ frame
            inv = torch.zeros(win, R, dtype=torch.float64)                    # frame = inv @ spec (window included)
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
            cache[key] = (fwd.float().to(device).contiguous(), inv.float().to(device).contiguous(), R,
                          (w * w).float().to(device))
        return cache[key]

    @staticmethod
    def _run(mod, x, T, F):
        if isinstance(mod, nn.Sequential):
            for m in mod:
                x, F = m.run(x, T, F)
            return x, F
        return mod.run(x, T, F)

    def forward(self, input, aux):
        if input.dim() != 2:
            raise RuntimeError("DPCCN expects [batch, samples]")
        if not (input.is_cuda and aux.is_cuda):
            raise RuntimeError("wesep_b200 kernels need CUDA tensors (no CPU fallback)")
        dev = input.device
        B, L = input.shape
        win, hop = self.win_len, self.hop_size
        Fq = win // 2 + 1
        fwd_b, inv_b, R, w2 = self._bases(dev)
        T = 1 + L // hop
        with torch.no_grad():                                   # the mixture is data: no gradient through the analysis
            x = input.float()
            pad = win // 2
            xp = torch.cat([x[:, 1:pad + 1].flip(1), x, x[:, L - pad - 1:L - 1].flip(1)], 1).contiguous()
            spec = ops.conv1x1_raw(ops.frames_raw(xp, win, T, hop), fwd_b, False, R)      # [B, R, T]: re rows | im rows
            # dpccn.py:221-224: stack(real, imag) -> transpose -> [B, 2, T, F]
            spec = ops.as_act(spec[:, :2 * Fq].reshape(B, 2, Fq, T).transpose(2, 3).reshape(B, 2, T * Fq))
        out = ops.conv3x3(spec, T, Fq, self.conv2d.weight, self.conv2d.bias, 1)            # [B, 16, T*F]
        out, F = self.encoder[0].run(out, T, Fq)

        predict_speaker_lable = torch.zeros((), device=dev)          # dummy, dpccn.py:229-230
        spk_in = aux
        if self.joint_training:                                    # dpccn.py:231-249
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
        # dpccn.py:251-254 + speaker.py:117-121 (4-D multiply): a gain per (row, frequency) shared by channels and frames
        gain = ops.LinearFn.apply(emb, self.spk_fuse.fc.linear.weight, self.spk_fuse.fc.linear.bias)       # [B, F]
        out = ops.ColScaleFn.apply(out, gain, T, F)
        out_list = [(out, F)]
        for enc in self.encoder[1:]:
            out, F = self._run(enc, out, T, F)
            out_list.append((out, F))

        # [B, 384, T*F] is already the TCN's [B, N, T*F] view (dpccn.py:252-253)
        for layer in self.tcn_layers:
            for blk in layer:
                out = blk.run(out)
        out_list = out_list[::-1]
        for idx, dec in enumerate(self.decoder):
            skip, Fs = out_list[idx]
            if Fs != F or skip.shape[2] != out.shape[2]:
                raise RuntimeError("DPCCN: skip connection shape mismatch")
            out, F = self._run(dec, ops.cat_act([skip, out]), T, F)
        # pyramidal pooling, dpccn.py:260-267
        pools = [out]
        for sz, avg in zip(self.pool_size, self.avg_pool):
            p = ops.AvgPool2dFn.apply(out, T, F, sz)
            p = ops.Conv1x1Fn.apply(p, avg[1].weight.reshape(8, 32), avg[1].bias, False, None)
            pools.append(ops.Upsample2dFn.apply(p, T // sz, F // sz, T, F))
        out = ops.Conv1x1Fn.apply(ops.cat_act(pools), self.avg_proj.weight.reshape(32, 64), self.avg_proj.bias, False, None)
        out = ops.conv_transpose3x3(out, T, F, self.deconv2d.weight, self.deconv2d.bias, 1)               # [B, 2, T*F]
        # dpccn.py:271-284: [B, 2, T, F] -> [B, 2, F, T] -> complex spectrum -> iSTFT
        est = out.reshape(B, 2, T, F).transpose(2, 3).reshape(B, 2 * F, T)
        if R > 2 * F:
            est = torch.cat([est, torch.zeros(B, R - 2 * F, T, device=dev)], 1)
        frames = ops.FixedGemmFn.apply(ops.as_act(est), inv_b, False)                       # [B, win, T]
        n_out = win + hop * (T - 1)
        y = ops.OverlapAddFn.apply(frames, hop, n_out)
        env = self.__dict__.setdefault("_env_cache", {}).get((str(dev), T))
        if env is None:
            e = torch.zeros(n_out, device=dev)
            for t in range(T):
                e[t * hop:t * hop + win] += w2
            env = self._env_cache[(str(dev), T)] = 1.0 / e[win // 2:win // 2 + L]
        return ops.ColVecMulFn.apply(y[:, win // 2:win // 2 + L], env), predict_speaker_lable
