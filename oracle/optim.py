"""ORACLE (test infrastructure) — optimiser side of the reference train step:
per-tensor gradient clip, Adam with coupled L2 decay, exponential LR schedule."""
import math

import torch


def clip_gradients(grads, clip):
    """clip_gradients — wesep/utils/funcs.py:79-88: PER-TENSOR L2 clip, in place.

    grads: list of tensors (None entries skipped). Returns the list of norms."""
    norms = []
    for g in grads:
        if g is None:
            continue
        param_norm = g.norm(2)
        norms.append(float(param_norm))
        clip_coef = clip / (param_norm + 1e-6)
        if clip_coef < 1:
            g.mul_(clip_coef)
    return norms


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-4):
    """torch.optim.Adam(weight_decay=wd) single step (coupled L2), as constructed at
    wesep/bin/train.py:235-238.  `step` is the 1-based step count AFTER increment.
    In place on params / exp_avg / exp_avg_sq (lists of tensors)."""
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        if g is None:
            continue
        g = g + weight_decay * p
        m.mul_(beta1).add_(g, alpha=1 - beta1)
        v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def exponential_decrease_lr(cur_iter, max_iter, initial_lr=1e-3, final_lr=2.5e-5, warm_up_iter=0, scale_ratio=1.0,
                            warm_from_zero=False):
    """ExponentialDecrease.get_current_lr — wesep/utils/schedulers.py:217-222 with
    BaseClass.get_multi_process_coeff :130-140."""
    lr_coeff = 1.0 * scale_ratio
    if cur_iter < warm_up_iter:
        if warm_from_zero:
            lr_coeff = scale_ratio * cur_iter / warm_up_iter
        elif scale_ratio > 1:
            lr_coeff = (scale_ratio - 1) * cur_iter / warm_up_iter + 1.0
    return lr_coeff * initial_lr * math.exp((cur_iter / max_iter) * math.log(final_lr / initial_lr))
