"""CPU: the pDPCCN oracle (oracle/dpccn.py) vs golden outputs of the REAL reference
(tests/golden/dpccn_*.npz from tests/golden/make_golden_dpccn.py)."""
import json
import os

import numpy as np
import torch

from oracle import dpccn as od
from oracle import losses as olosses
from wesep_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, json.loads(str(z["meta"]))


def inputs(meta, dtype=torch.float32):
    b = synth.make_batch(meta["n"], T=meta["L"], Te=8, seed=meta["dseed"])
    rng = np.random.default_rng(meta["dseed"] + 77)
    emb = torch.from_numpy(rng.standard_normal((meta["n"], 256)).astype(np.float32))
    return b["wav_mix"].to(dtype), b["wav_targets"].to(dtype), emb.to(dtype)


def state_dict(meta, dtype=torch.float32):
    a = meta["args"]
    sd = od.make_state_dict(tcn_blocks=a["tcn_blocks"], tcn_layers=a["tcn_layers"])
    synth.fill_state_dict_(sd, seed=meta["wseed"])
    return {k: v.to(dtype) for k, v in sd.items()}


def test_dpccn_small_golden():
    """forward, per-row SI-SDR, loss and every gradient of the small case, oracle in fp64 vs the fp32 reference run."""
    z, meta = load("dpccn_small_train")
    a = meta["args"]
    sd = {k: v.double().requires_grad_(True) for k, v in state_dict(meta).items()}
    mix, tgt, emb = inputs(meta, torch.float64)
    est = od.dpccn_forward(sd, mix, emb, tcn_blocks=a["tcn_blocks"], tcn_layers=a["tcn_layers"])
    ref = torch.from_numpy(z["out0"]).double()
    assert est.shape == ref.shape
    assert float((est.detach() - ref).norm() / ref.norm()) <= 2e-4
    rows = olosses.sisdr_per_row(est.detach(), tgt).numpy()
    assert np.abs(rows - z["sisdr_rows0"]).max() <= 0.01
    loss = olosses.sisdr_loss(est, tgt)
    assert abs(float(loss) - float(z["loss"])) <= 2e-3
    loss.backward()
    worst = 0.0
    for k, p in sd.items():
        gn, ref_n = float(p.grad.norm()), float(z["gnorm/" + k])
        worst = max(worst, abs(gn - ref_n) / (ref_n + 1e-6))
        assert abs(gn - ref_n) <= 5e-3 * ref_n + 1e-5, (k, gn, ref_n)
    print("worst relative gradient-norm difference", worst)
