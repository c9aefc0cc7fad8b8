"""CPU: the TF-GridNet oracle (oracle/tfgridnet.py) vs golden outputs of the REAL reference
(tests/golden/tfgridnet_*.npz from tests/golden/make_golden_tfgridnet.py)."""
import numpy as np
import torch

from oracle import losses as olosses
from oracle import tfgridnet as ot
from tests.test_oracle_dpccn import inputs, load
from wesep_b200 import synth


def state_dict(meta):
    a = meta["args"]
    sd = ot.make_state_dict(n_layers=a["n_layers"], emb_dim=a["emb_dim"], hidden=a["lstm_hidden_units"], n_head=a["attn_n_head"],
                            approx_qk_dim=a["attn_approx_qk_dim"], n_fft=a["n_fft"], emb_ks=a["emb_ks"], emb_hs=a["emb_hs"])
    synth.fill_state_dict_(sd, seed=meta["wseed"])
    return sd


import pytest  # noqa: E402


@pytest.mark.parametrize("name", ["tfgridnet_small_train", "tfgridnet_small_ks4", "tfgridnet_small_ks2"])
def test_tfgridnet_small_golden(name):
    """forward, per-row SI-SDR, loss and every gradient norm of the small cases (emb_ks 1 and the unfold path emb_ks 4 / emb_hs 1),
    oracle in fp64 vs the fp32 reference run."""
    z, meta = load(name)
    a = meta["args"]
    sd = {k: v.double().requires_grad_(True) for k, v in state_dict(meta).items()}
    mix, tgt, emb = inputs(meta, torch.float64)
    est = ot.tfgridnet_forward(sd, mix, emb, n_fft=a["n_fft"], stride=a["stride"], n_layers=a["n_layers"], n_head=a["attn_n_head"],
                               eps=a["eps"], emb_ks=a["emb_ks"], emb_hs=a["emb_hs"])
    ref = torch.from_numpy(z["out0"]).double()
    assert est.shape == ref.shape
    assert float((est.detach() - ref).norm() / ref.norm()) <= 2e-4
    rows = olosses.sisdr_per_row(est.detach(), tgt).numpy()
    assert np.abs(rows - z["sisdr_rows0"]).max() <= 0.01
    loss = olosses.sisdr_loss(est, tgt)
    assert abs(float(loss.detach()) - float(z["loss"])) <= 2e-3
    loss.backward()
    for k, p in sd.items():
        gn, ref_n = float(p.grad.norm()), float(z["gnorm/" + k])
        assert abs(gn - ref_n) <= 5e-3 * ref_n + 1e-5, (k, gn, ref_n)
