"""SpeakerTransform — reference wesep/modules/common/speaker.py:26-49 (tiny 1x1 convs on the
[n, 256, 1] embedding; not a hot op, kept on torch)."""
import torch.nn as nn


class SpeakerTransform(nn.Module):

    def __init__(self, embed_dim=256, num_layers=3, hid_dim=128):
        super().__init__()
        layers = [nn.Conv1d(embed_dim, hid_dim, 1)]
        for _ in range(num_layers - 2):
            layers.append(nn.Conv1d(hid_dim, hid_dim, 1))
            layers.append(nn.Tanh())
        layers.append(nn.Conv1d(hid_dim, embed_dim, 1))
        self.transforms = nn.Sequential(*layers)

    def forward(self, x):
        if len(x.size()) == 2:
            return self.transforms(x.unsqueeze(-1)).squeeze(-1)
        return self.transforms(x)
