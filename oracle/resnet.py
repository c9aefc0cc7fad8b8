"""ORACLE (test infrastructure, never imported by the product): plain-torch restatement of the wespeaker ResNet (BasicBlock)
speaker encoder with TSTP pooling that pBSRNN calls as `self.spk_model(fbank)[-1]` (wesep/models/bsrnn.py:217,352-356;
configured by examples/librimix/tse/v2/confs/bsrnn.yaml:56-64).

PARITY UNPINNED: wespeaker is an external package that is neither vendored in /root/reference nor installable here, and the
reference holds no test or golden vector for it.  The restatement follows the published architecture
(wespeaker/models/resnet.py `ResNet`, `BasicBlock`; wespeaker/models/pooling_layers.py `TSTP`):
  x (B, T, F) -> permute (B, F, T) -> unsqueeze(1) -> relu(bn1(conv3x3(1 -> m))) -> layer1..4 of BasicBlocks
  (m, 2m, 4m, 8m channels; strides 1, 2, 2, 2; block = relu(bn2(conv3x3(relu(bn1(conv3x3(x))))) + shortcut(x)), shortcut =
  bn(conv1x1 stride s) when the shape changes) -> reshape (B, C * F / 8, T / 8) -> TSTP = cat(mean_t, sqrt(var_t (unbiased)
  + 1e-7)) -> seg_1 Linear -> embedding; returns (tensor(0.0), embedding) when two_emb_layer is False."""
import torch
import torch.nn.functional as F


def _bn(x, sd, pre, training, buffers_out=None, momentum=0.1, eps=1e-5):
    rm, rv = sd[pre + ".running_mean"], sd[pre + ".running_var"]
    if training:
        rm, rv = rm.clone(), rv.clone()
    y = F.batch_norm(x, rm, rv, sd[pre + ".weight"], sd[pre + ".bias"], training, momentum, eps)
    if training and buffers_out is not None:
        buffers_out[pre + ".running_mean"], buffers_out[pre + ".running_var"] = rm, rv
    return y


def resnet_forward(sd, x, num_blocks=(3, 4, 6, 3), prefix="", training=True, buffers_out=None):
    """sd: state dict (keys as wespeaker's ResNet, optionally under `prefix`); x [B, T, F] -> embedding [B, embed_dim]."""
    p = prefix
    h = x.permute(0, 2, 1).unsqueeze(1)
    h = F.relu(_bn(F.conv2d(h, sd[p + "conv1.weight"], None, 1, 1), sd, p + "bn1", training, buffers_out))
    for li, nb in enumerate(num_blocks, start=1):
        for bi in range(nb):
            q = f"{p}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            y = F.relu(_bn(F.conv2d(h, sd[q + "conv1.weight"], None, stride, 1), sd, q + "bn1", training, buffers_out))
            y = _bn(F.conv2d(y, sd[q + "conv2.weight"], None, 1, 1), sd, q + "bn2", training, buffers_out)
            if (q + "shortcut.0.weight") in sd:
                sc = _bn(F.conv2d(h, sd[q + "shortcut.0.weight"], None, stride, 0), sd, q + "shortcut.1", training, buffers_out)
            else:
                sc = h
            h = F.relu(y + sc)
    B = h.shape[0]
    h = h.reshape(B, h.shape[1] * h.shape[2], h.shape[3])
    stats = torch.cat([h.mean(-1), torch.sqrt(h.var(-1) + 1e-7)], 1)
    return F.linear(stats, sd[p + "seg_1.weight"], sd[p + "seg_1.bias"])
