"""Learning-rate schedule of the recipes as a closed form (what the reference's `ExponentialDecrease` scheduler,
wesep/utils/schedulers.py:217-222 with the warm-up coefficient of :130-140, evaluates each iteration).  The reference's own
scheduler object can drive `FusedClipAdam` unchanged (it only writes `param_groups[i]["lr"]`); this helper exists for
bench.py and the tests, which have no need for a scheduler class."""
import math


def exponential_decrease_lr(it, max_iter, initial_lr=1e-3, final_lr=2.5e-5, warm_up_iter=0, scale_ratio=1.0, warm_from_zero=False):
    """lr(it) = coeff(it) * initial_lr * (final_lr / initial_lr) ** (it / max_iter)."""
    coeff = float(scale_ratio)
    if it < warm_up_iter:
        frac = it / warm_up_iter
        if warm_from_zero:
            coeff = scale_ratio * frac
        elif scale_ratio > 1:
            coeff = 1.0 + (scale_ratio - 1.0) * frac
    return coeff * initial_lr * math.exp(math.log(final_lr / initial_lr) * it / max_iter)


def set_lr(optimizer, lr):
    for group in optimizer.param_groups:
        group["lr"] = lr
    return lr
