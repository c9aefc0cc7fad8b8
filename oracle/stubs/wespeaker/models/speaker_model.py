"""Import stub so the reference's model files can be imported in this container.

The reference does `from wespeaker.models.speaker_model import get_speaker_model`
(wesep/models/convtasnet.py:11) but wespeaker is not installed. Spex+ never calls it
(convtasnet.py:93-98 uses the in-tree ResNet4SpExplus). Test infrastructure only.
"""
import torch
import torch.nn as nn


class _Tiny(nn.Module):
    def __init__(self, feat_dim=80, embed_dim=256, **kw):
        super().__init__()
        self.fc = nn.Linear(feat_dim, embed_dim)

    def forward(self, x):  # x [n, frames, feat_dim]
        return torch.tensor(0.0), self.fc(x.mean(1))


def get_speaker_model(name):
    return _Tiny
