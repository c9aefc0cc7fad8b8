"""ConvTasNet (Spex+) — same class name, constructor kwargs, forward signature and state_dict
keys as the reference wesep/models/convtasnet.py:14-219; the forward runs on libwesep_b200."""
import torch
import torch.nn as nn

from wesep_b200 import ops
from wesep_b200.modules.common.speaker import SpeakerTransform
from wesep_b200.modules.tasnet import FuseSeparation, MultiDecoder, MultiEncoder, ResNet4SpExplus


class ConvTasNet(nn.Module):

    def __init__(
        self,
        N=512,
        L=16,
        B=128,
        H=512,
        P=3,
        X=8,
        R=3,
        spk_emb_dim=256,
        norm="gLN",
        activate="relu",
        causal=False,
        skip_con=False,
        spk_fuse_type="concatConv",
        multi_fuse=True,
        use_spk_transform=True,
        encoder_type="Multi",
        decoder_type="Multi",
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
        if encoder_type != "Multi" or decoder_type != "Multi":
            raise NotImplementedError("only encoder_type/decoder_type 'Multi' (Spex+) are accelerated")
        self.encoder_type = encoder_type
        self.decoder_type = decoder_type
        self.encoder = MultiEncoder(in_channels=1, middle_channels=N, out_channels=B, kernel_size=L, stride=L // 2)
        self.joint_training = joint_training
        self.spk_feat = spk_feat
        self.feat_type = feat_type
        self.spk_model_freeze = spk_model_freeze
        self.multi_task = multi_task
        if joint_training:
            if spk_feat or feat_type != "consistent":
                raise NotImplementedError("Spex+ path: spk_feat False / feat_type 'consistent' (shared encoder + ResNet)")
            self.spk_model = ResNet4SpExplus(in_channel=N, C_embedding=spk_emb_dim)
            if multi_task:
                self.pred_linear = nn.Linear(spk_emb_dim, spksInTrain)
        if not use_spk_transform:
            self.spk_transform = nn.Identity()
        else:
            self.spk_transform = SpeakerTransform()
        self.separation = FuseSeparation(R, X, B, H, P, norm=norm, causal=causal, skip_con=skip_con,
                                         C_embedding=spk_emb_dim, spk_fuse_type=spk_fuse_type, multi_fuse=multi_fuse)
        self.decoder = MultiDecoder(in_channels=B, middle_channels=N, out_channels=1, kernel_size=L, stride=L // 2)
        active_f = {"relu": nn.ReLU(), "sigmoid": nn.Sigmoid(), "softmax": nn.Softmax(dim=0)}
        self.activation = active_f[activate]

    def forward(self, x, embeddings):
        if x.dim() >= 3:
            raise RuntimeError("{} accept 1/2D tensor as input, but got {:d}".format(self.__class__.__name__, x.dim()))
        if x.dim() == 1:
            x = torch.unsqueeze(x, 0)
        w = self.encoder.filterbank(x)                      # cat[w1,w2,w3]  (convtasnet.py:171)
        e = self.encoder.proj(self.encoder.ln(w))
        predict_speaker_lable = None
        if self.joint_training:
            # reference :184 runs the whole encoder on the enrollment and discards ln+proj; only the
            # filterbank outputs are used, so only those are computed here
            if embeddings.dim() == 1:
                embeddings = embeddings.unsqueeze(0)
            aux = self.encoder.filterbank(embeddings)       # cat[aux_w1, aux_w2, aux_w3]  (:185)
            embeddings = self.spk_model(aux)                # (:190)
            if self.multi_task:
                predict_speaker_lable = ops.LinearFn.apply(embeddings, self.pred_linear.weight, self.pred_linear.bias)
        spk_embeds = self.spk_transform(embeddings.unsqueeze(-1))
        e = self.separation(e, spk_embeds)
        s = self.decoder.forward_cat(e, w) if isinstance(self.activation, nn.ReLU) else None
        if s is None:
            raise NotImplementedError("only activate='relu' is accelerated (recipe setting)")
        if self.joint_training and self.multi_task:
            s.append(predict_speaker_lable)
        return s
