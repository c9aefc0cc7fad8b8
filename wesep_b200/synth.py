"""Seeded synthetic batches and parameter fills (numpy PCG64: identical on every machine).

Batch scheme = SURVEY.md §8d / BASELINE.md §3: per mixture ``s1, s2 ~ 0.1*N(0,1) [T]``,
``mix = s1 + s2``; model rows are (mixture, target) pairs like ``tse_collate_fn``
(wesep/dataset/dataset.py:217-227): row 2m = (mix_m, s1_m), row 2m+1 = (mix_m, s2_m);
enrollment ``0.1*N(0,1) [n, Te]`` raw wave (Spex+, spk_feat False); labels randint(251).
"""
import numpy as np
import torch


def make_batch(n_rows, T=64000, Te=64000, n_spk=251, seed=1234, device="cpu", dtype=torch.float32, pin=False):
    rng = np.random.default_rng(seed)
    n_mix = (n_rows + 1) // 2
    mix = np.empty((n_rows, T), np.float32)
    tgt = np.empty((n_rows, T), np.float32)
    for m in range(n_mix):
        s1 = (0.1 * rng.standard_normal(T)).astype(np.float32)
        s2 = (0.1 * rng.standard_normal(T)).astype(np.float32)
        mx = s1 + s2
        for j, s in enumerate((s1, s2)):
            r = 2 * m + j
            if r < n_rows:
                mix[r] = mx
                tgt[r] = s
    enroll = (0.1 * rng.standard_normal((n_rows, Te))).astype(np.float32)
    label = rng.integers(0, n_spk, size=(n_rows,)).astype(np.int64)
    out = dict(wav_mix=torch.from_numpy(mix).to(dtype), wav_targets=torch.from_numpy(tgt).to(dtype),
               spk_embeds=torch.from_numpy(enroll).to(dtype), spk_label=torch.from_numpy(label))
    if pin and torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    if device != "cpu":
        out = {k: v.to(device) for k, v in out.items()}
    return out


def fill_state_dict_(sd, seed=0):
    """Deterministically overwrite every entry of a state_dict (in place, any device).

    conv/linear weights ~ U(-1/sqrt(fan_in), +); biases 0.02*N; norm scales 1+0.1*N, norm
    shifts 0.1*N; PReLU slopes 0.25+0.05*N; BatchNorm buffers reset.  Perturbed norm/PReLU
    values (instead of the 1/0/0.25 defaults) make scale/shift/slope gradient bugs visible.
    """
    rng = np.random.default_rng(seed)
    for name, t in sd.items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            val = np.zeros(shape, np.int64)
        elif name.endswith("running_mean"):
            val = np.zeros(shape, np.float32)
        elif name.endswith("running_var"):
            val = np.ones(shape, np.float32)
        elif "prelu" in name.lower() or name.endswith(".act.weight") or "attn_concat_proj.1." in name:   # (TF-GridNet PReLUs)
            val = 0.25 + 0.05 * rng.standard_normal(shape)
        elif len(shape) >= 2 and not (len(shape) == 2 and shape[1] == 1):
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / np.sqrt(fan_in)
            val = rng.uniform(-b, b, size=shape)
        elif name.endswith("weight"):
            val = 1.0 + 0.1 * rng.standard_normal(shape)      # norm scale (gLN (C,1), cLN/BN (C,))
        elif any(k in name for k in ("norm", "ln.", "lnorm", "aux_enc3.0.")):
            val = 0.1 * rng.standard_normal(shape)            # norm shift
        else:
            val = 0.02 * rng.standard_normal(shape)           # conv / linear bias
        with torch.no_grad():
            t.copy_(torch.from_numpy(np.asarray(val)).to(t.dtype))
    return sd
