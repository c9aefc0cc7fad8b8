"""ORACLE (test infrastructure) — loss arithmetic of the reference train step.

The reference takes SI-SDR from the third-party package ``auraloss`` (un-pinned,
requirements.txt:26; not installed in this image): ``wesep/utils/losses.py:24-25`` maps
"SISDR"/"SISNR" to ``auraloss.time.SISDRLoss()``.  Restated here from auraloss 0.4.x
``time.SISDRLoss(zero_mean=True, eps=1e-8, reduction="mean")`` (published algorithm,
SURVEY.md §8c): parity for this function is anchored on the reference call sites
(executor.py:115-122) and cross-checked with the in-tree numpy ``cal_SISNR``
(wesep/utils/score.py:7-21) restated below.
"""
import numpy as np
import torch
import torch.nn.functional as F


def sisdr_per_row(x, t, eps=1e-8, zero_mean=True):
    """Per-row SI-SDR in dB (positive = good) — auraloss.time.SISDRLoss body."""
    if zero_mean:
        x = x - torch.mean(x, dim=-1, keepdim=True)
        t = t - torch.mean(t, dim=-1, keepdim=True)
    alpha = (x * t).sum(-1) / ((t ** 2).sum(-1) + eps)
    tt = t * alpha.unsqueeze(-1)
    res = x - tt
    return 10 * torch.log10((tt ** 2).sum(-1) / ((res ** 2).sum(-1) + eps) + eps)


def sisdr_loss(x, t, eps=1e-8):
    """auraloss.time.SISDRLoss()(x, t): -mean over rows (reduction='mean')."""
    return -torch.mean(sisdr_per_row(x, t, eps))


def cal_sisnr_numpy(ref_sig, out_sig, eps=1e-8):
    """cal_SISNR — wesep/utils/score.py:7-21 (numpy, single pair of 1-D signals)."""
    assert len(ref_sig) == len(out_sig)
    ref_sig = ref_sig - np.mean(ref_sig)
    out_sig = out_sig - np.mean(out_sig)
    ref_energy = np.sum(ref_sig ** 2) + eps
    proj = np.sum(ref_sig * out_sig) * ref_sig / ref_energy
    noise = out_sig - proj
    ratio = np.sum(proj ** 2) / (np.sum(noise ** 2) + eps)
    return 10 * np.log(ratio + eps) / np.log(10.0)


def train_loss(outputs, targets, spk_label, loss_posi=((0, 1, 2), (3,)), loss_weight=((0.8, 0.1, 0.1), (0.5,)),
               multi_task=True):
    """Loss weighting of Executor.train — wesep/utils/executor.py:105-122 with
    criterion = [SISDR, CE] (spexplus.yaml:26-30); accum_grad = 1.

    Returns (loss, dict of parts)."""
    loss = 0
    parts = {}
    # criterion[0] = SISDRLoss on the listed output positions
    for ji, pos in enumerate(loss_posi[0]):
        l = sisdr_loss(outputs[pos], targets).mean()
        parts[f"sisdr{pos}"] = l
        loss = loss + loss_weight[0][ji] * l
    if multi_task and len(loss_posi) > 1:
        for ji, pos in enumerate(loss_posi[1]):
            l = F.cross_entropy(outputs[pos], spk_label).mean()
            parts[f"ce{pos}"] = l
            loss = loss + loss_weight[1][ji] * l
    return loss, parts
