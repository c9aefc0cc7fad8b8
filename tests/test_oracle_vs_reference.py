"""Build-container only: oracle restatement vs the REAL reference modules imported from
/root/reference (skipped where the reference tree does not exist, e.g. the GPU box)."""
import pytest
import torch

from oracle import ref_loader
from oracle import spexplus as ospex
from wesep_b200 import synth

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.mark.parametrize("ft", ["concatConv", "FiLM", "multiply"])
def test_state_dict_contract(ft):
    ref_loader.import_reference()
    from wesep.models import get_model
    args = dict(ref_loader.SPEXPLUS_ARGS)
    args.update(spk_fuse_type=ft, B=64, H=128, X=3, R=2)
    m = get_model("ConvTasNet")(**args)
    ref = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    cfg = dict(ospex.DEFAULT_CFG)
    cfg.update(B=64, H=128, X=3, R=2, spk_fuse_type=ft)
    assert ref == ospex.state_dict_spec(cfg)


def test_forward_backward_matches_reference():
    ref_loader.import_reference()
    from wesep.models import get_model
    args = dict(ref_loader.SPEXPLUS_ARGS)
    args.update(B=64, H=128, X=3, R=2)
    m = get_model("ConvTasNet")(**args)
    synth.fill_state_dict_(m.state_dict(), seed=21)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    for v in sd.values():
        pass
    for k in [k for k, _ in m.named_parameters()]:
        sd[k].requires_grad_(True)
    b = synth.make_batch(2, T=2400, Te=1800, seed=22)
    m.train()
    out_r = m(b["wav_mix"], b["spk_embeds"])
    cfg = dict(ospex.DEFAULT_CFG)
    cfg.update(B=64, H=128, X=3, R=2)
    out_o = ospex.convtasnet_forward(sd, cfg, b["wav_mix"], b["spk_embeds"], training=True)
    for a, c in zip(out_r, out_o):
        assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)
    sum(o.square().sum() for o in out_r).backward()
    sum(o.square().sum() for o in out_o).backward()
    for k, p in m.named_parameters():
        assert torch.allclose(p.grad, sd[k].grad, rtol=1e-4, atol=1e-6), k
