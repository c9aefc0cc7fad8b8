"""Loss table — reference wesep/utils/losses.py:8-41 (names SISDR / SISNR / CE on this path)."""
import torch.nn as nn

from wesep_b200 import ops


class SISDRLoss(nn.Module):
    """auraloss.time.SISDRLoss() replacement (zero_mean, eps 1e-8, mean reduction) on the fused kernel."""

    def forward(self, input, target):
        if input.dim() == 3 and input.shape[1] == 1:
            input, target = input.squeeze(1), target.squeeze(1)
        losses, _ = ops.sisdr_losses([input.float()], target.float())
        return losses[0]


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() replacement (mean reduction, class-index targets) on the library kernel."""

    def forward(self, input, target):
        return ops.cross_entropy(input, target)


valid_losses = {"SISDR": SISDRLoss(), "SISNR": SISDRLoss(), "CE": CrossEntropyLoss()}


def parse_loss(loss):
    loss_functions = []
    if not isinstance(loss, list):
        loss = [loss]
    for name in loss:
        if name not in valid_losses:
            raise NotImplementedError("loss %r is outside the accelerated path" % (name,))
        loss_functions.append(valid_losses[name])
    return loss_functions
