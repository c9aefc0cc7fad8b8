"""CPU: the oracle restatement (oracle/*.py) vs golden outputs of the REAL reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import losses as olosses
from oracle import optim as ooptim
from oracle import spexplus as ospex
from tests.util import fixture_inputs, load_fixture

SMALL_CASES = ["spex_small_train", "spex_small_eval", "spex_small_n1", "spex_small_FiLM", "spex_small_multiply",
               "spex_small_additive", "spex_small_concat"]


def _run(name, backward):
    z, meta = load_fixture(name)
    cfg, sd, b = fixture_inputs(meta)
    params = {}
    if backward:
        for k, v in sd.items():
            if v.is_floating_point() and "running_" not in k:
                v.requires_grad_(True)
                params[k] = v
    bufs = {}
    out = ospex.convtasnet_forward(sd, cfg, b["wav_mix"], b["spk_embeds"], training=meta["train"], buffers_out=bufs)
    tgt = b["wav_targets"][:, :out[0].shape[-1]]
    loss, parts = olosses.train_loss(out, tgt, b["spk_label"], multi_task=cfg["multi_task"])
    sub = meta["subsample"]
    for i in range(3):
        ref = torch.from_numpy(z[f"out{i}"])
        got = out[i].detach()[..., ::sub]
        assert got.shape == ref.shape
        assert torch.allclose(got, ref, rtol=1e-3, atol=2e-5), (name, i, float((got - ref).abs().max()))
        s = olosses.sisdr_per_row(out[i].detach().double(), tgt.double()).numpy()
        assert np.max(np.abs(s - z[f"sisdr_rows{i}"])) <= 0.01, (name, i)   # dB, north-star tolerance
    if len(out) > 3:
        assert torch.allclose(out[3].detach(), torch.from_numpy(z["out3"]), rtol=1e-3, atol=1e-4)
    assert abs(float(loss) - float(z["loss"])) <= 1e-3 * abs(float(z["loss"])) + 1e-3
    if backward and meta["backward"]:
        loss.backward()
        for k, p in params.items():
            gn = float(p.grad.double().norm())
            ref = float(z["gnorm/" + k])
            assert abs(gn - ref) <= 2e-3 * ref + 1e-6, (name, k, gn, ref)
            if ("g/" + k) in z:
                g = torch.from_numpy(z["g/" + k])
                assert (p.grad - g).norm() <= 2e-3 * g.norm() + 1e-6, (name, k)
    if meta["train"]:
        for k, v in bufs.items():
            assert torch.allclose(v, torch.from_numpy(z["buf/" + k]), rtol=1e-3, atol=1e-5), (name, k)


@pytest.mark.parametrize("name", SMALL_CASES)
def test_oracle_small(name):
    _run(name, backward=True)


def test_oracle_full_cfg1_eval():
    """BASELINE config 1: Spex+ forward + SI-SNR on one 2-speaker 4 s mixture (n=2)."""
    _run("spex_full_cfg1_eval", backward=False)


@pytest.mark.slow
def test_oracle_full_cfg1_train():
    _run("spex_full_cfg1_train", backward=True)


def test_optim_golden():
    z = np.load("tests/golden/optim.npz")
    n = 5
    params = [torch.from_numpy(z[f"p0_{i}"].copy()) for i in range(n)]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    for step in range(3):
        grads = [torch.from_numpy(z[f"g{step}_{i}"].copy()) for i in range(n)]
        norms = ooptim.clip_gradients(grads, 5.0)
        assert np.allclose(norms, z[f"norms{step}"], rtol=1e-5)
        ooptim.adam_step(params, grads, m, v, step + 1, float(z["lrs"][step]))
        for i in range(n):
            assert torch.allclose(params[i], torch.from_numpy(z[f"p{step + 1}_{i}"]), rtol=2e-5, atol=1e-7)


def test_sched_golden():
    z = np.load("tests/golden/sched.npz")
    for it, lr in zip(z["its"], z["lrs"]):
        mine = ooptim.exponential_decrease_lr(int(it), 150 * 1000)
        assert abs(mine - lr) <= 1e-12 + 1e-9 * lr


def test_sisnr_cross_check():
    """restated auraloss SISDRLoss vs in-tree cal_SISNR formula: <= 2e-6 dB."""
    z = np.load("tests/golden/sisnr.npz")
    for snr, a, c in z["rows"]:
        assert abs(a - c) < 1e-4
    rng = np.random.default_rng(3)
    t = rng.standard_normal(4000)
    x = 0.3 * t + 0.1 * rng.standard_normal(4000)
    a = float(olosses.sisdr_per_row(torch.from_numpy(x)[None], torch.from_numpy(t)[None])[0])
    c = float(olosses.cal_sisnr_numpy(t, x))
    assert abs(a - c) < 1e-6
