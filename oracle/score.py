"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the evaluation scoring of the reference:
  * cal_SISNR / cal_SISNRi — wesep/utils/score.py:7-36 (numpy, one pair of 1-D signals);
  * the peak rule and the trim-to-shortest of the inference loop — wesep/bin/infer.py:124-129,147-152.

Pinned: tests/golden/score.npz holds the outputs of the REAL reference functions
(wesep.utils.score.cal_SISNRi imported in place, tests/golden/make_golden_score.py) on seeded
inputs; tests/test_oracle_golden.py checks this restatement against them.  infer.py is a script
(its loop is not importable), so the peak rule is pinned only by reading: it is restated from
infer.py:124-129 line by line with torch CPU ops as the reference uses.
"""
import numpy as np
import torch


def cal_sisnr(est, ref, eps=1e-8):
    """score.py:7-21."""
    assert len(est) == len(ref)
    e = est - np.mean(est)
    r = ref - np.mean(ref)
    t = np.sum(e * r) * r / (np.linalg.norm(r) ** 2 + eps)
    return 20 * np.log10(eps + np.linalg.norm(t) / (np.linalg.norm(e - t) + eps))


def cal_sisnri(est, ref, mix, eps=1e-8):
    """score.py:24-36."""
    assert len(est) == len(ref) == len(mix)
    a = cal_sisnr(est, ref, eps)
    b = cal_sisnr(mix, ref, eps)
    return a, a - b


def peak_rule(outputs):
    """infer.py:124-129 on a [n, T] fp32 torch tensor -> numpy [n, T]."""
    if torch.min(outputs.max(dim=1).values) > 0:
        return (outputs / abs(outputs).max(dim=1, keepdim=True)[0] * 0.9).cpu().numpy()
    return outputs.cpu().numpy()


def score_rows(outputs, targets, mix):
    """The per-row scoring of infer.py:144-172 for any number of rows: returns
    (waves [n, T], sisnr [n], sisnri [n]) as the reference computes them (fp32 numpy)."""
    ests = peak_rule(outputs.float())
    ref = targets.float().cpu().numpy()
    mx = mix.float().cpu().numpy()
    s, d = [], []
    for r in range(ests.shape[0]):
        if ests[r].size != ref[r].size:
            end = min(ests[r].size, ref[r].size, mx[r].size)
            a, b = cal_sisnri(ests[r][:end], ref[r][:end], mx[r][:end])
        else:
            a, b = cal_sisnri(ests[r], ref[r], mx[r])
        s.append(a)
        d.append(b)
    return ests, np.array(s, np.float64), np.array(d, np.float64)
