"""GPU: evaluation path — batched peak rule + SI-SNR / SI-SNRi kernel and the inference loop vs
oracle/score.py (pinned to the real wesep.utils.score by tests/golden/score.npz)."""
import numpy as np
import pytest
import torch

from oracle import score as oscore
from tests.util import score_case

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_DB = 1e-4     # vs the fp64 evaluation of the reference formula
TOL_DB32 = 1e-3   # vs the reference's own fp32 numpy evaluation (its rounding, not ours)


def test_score_golden_rows():
    """Each golden row through the C-ABI kernel: the values the real reference printed."""
    from wesep_b200.utils.score import cal_SISNRi, cal_SISNR
    z = np.load("tests/golden/score.npz")
    for seed, snr, T, s, d, s64 in z["rows"]:
        est, ref, mix = (torch.from_numpy(v).to(DEV) for v in score_case(int(seed), float(snr), int(T)))
        a, b = cal_SISNRi(est, ref, mix)
        assert abs(float(a) - s64) <= TOL_DB, (seed, float(a), s64)
        assert abs(float(a) - s) <= TOL_DB32 and abs(float(b) - d) <= TOL_DB32
        assert abs(float(cal_SISNR(est, ref)) - s64) <= TOL_DB


@pytest.mark.parametrize("all_positive", [True, False])
@pytest.mark.parametrize("n,T", [(2, 64000), (5, 20001), (33, 8192 * 2 + 3)])
def test_score_batch_vs_oracle(n, T, all_positive):
    """Batch call: peak rule (both branches of infer.py:124), waves bit-exact with the reference's fp32 scaling,
    scores vs the oracle evaluated in fp32 (as the reference does) and fp64."""
    from wesep_b200.utils.score import score_batch
    g = torch.Generator().manual_seed(n * 7 + T)
    ref = torch.randn(n, T, generator=g) * 0.1
    mix = ref + torch.randn(n, T, generator=g) * 0.1
    est = ref * torch.rand(n, 1, generator=g) * 3 + torch.randn(n, T, generator=g) * 0.03
    if not all_positive:
        est[n // 2] = -est[n // 2].abs()
    est0 = est.clone()
    waves, s, d, normed = score_batch(est.to(DEV), ref.to(DEV), mix.to(DEV))
    ow, os_, od = oscore.score_rows(est, ref, mix)
    assert int(normed) == int(all_positive)
    assert torch.equal(est, est0)
    assert np.array_equal(waves.cpu().numpy(), ow)                     # same fp32 division and multiply
    assert np.abs(s.cpu().numpy() - os_).max() <= TOL_DB32
    assert np.abs(d.cpu().numpy() - od).max() <= TOL_DB32
    for r in range(0, n, max(1, n // 4)):                              # fp64 truth on the scaled waves
        a, b = oscore.cal_sisnri(ow[r].astype(np.float64), ref[r].numpy().astype(np.float64),
                                 mix[r].numpy().astype(np.float64))
        assert abs(float(s[r]) - a) <= TOL_DB and abs(float(d[r]) - b) <= TOL_DB


def test_score_ragged_and_trim():
    """Estimate longer/shorter than the targets (decoder returns (K-1)*hop+L samples): infer.py:147-152 scores the
    common prefix but scales with the peak of the whole estimate; per-row lengths for batched ragged utterances."""
    from wesep_b200.utils.score import score_batch
    g = torch.Generator().manual_seed(5)
    n, Te, Tr = 4, 9990, 10000
    ref = torch.randn(n, Tr, generator=g) * 0.1
    mix = ref + torch.randn(n, Tr, generator=g) * 0.1
    est = ref[:, :Te] * 0.8 + torch.randn(n, Te, generator=g) * 0.02
    est[:, -1] = 5.0                                                   # the peak sits inside the scored prefix's last sample
    waves, s, d, _ = score_batch(est.to(DEV), ref.to(DEV), mix.to(DEV))
    ow, os_, od = oscore.score_rows(est, ref, mix)
    assert np.array_equal(waves.cpu().numpy(), ow)
    assert np.abs(s.cpu().numpy() - os_).max() <= TOL_DB32 and np.abs(d.cpu().numpy() - od).max() <= TOL_DB32
    # est longer than the targets
    est2 = torch.cat([est, torch.full((n, 50), 7.0)], dim=1)
    waves2, s2, d2, _ = score_batch(est2.to(DEV), ref.to(DEV), mix.to(DEV))
    ow2, os2, od2 = oscore.score_rows(est2, ref, mix)
    assert np.array_equal(waves2.cpu().numpy(), ow2)
    assert np.abs(s2.cpu().numpy() - os2).max() <= TOL_DB32 and np.abs(d2.cpu().numpy() - od2).max() <= TOL_DB32
    # explicit per-row lengths
    lens = [Te, 100, 1, 7777]
    _, s3, d3, _ = score_batch(est.to(DEV), ref.to(DEV), mix.to(DEV), lengths=lens, peak_norm=False)
    for r, L in enumerate(lens):
        a, b = oscore.cal_sisnri(est[r, :L].double().numpy(), ref[r, :L].double().numpy(), mix[r, :L].double().numpy())
        assert abs(float(s3[r]) - a) <= TOL_DB and abs(float(d3[r]) - b) <= TOL_DB


def test_score_errors():
    from wesep_b200.utils.score import cal_SISNRi, score_batch
    x = torch.zeros(2, 100, device=DEV)
    with pytest.raises(AssertionError):
        cal_SISNRi(x, x[:, :50], x)
    with pytest.raises(RuntimeError):
        score_batch(x.cpu(), x.cpu(), x.cpu())
    with pytest.raises(RuntimeError):
        score_batch(x, x[:1], x)


@pytest.mark.parametrize("family", ["spex", "bsrnn"])
def test_run_inference_loop(family):
    """infer.py:108-181 end to end on a small model: eval-mode whole-utterance forward (T not a multiple of any hop),
    first output taken, peak rule, scores; compared with the oracle scoring of the same separated waves."""
    from wesep_b200 import synth
    from wesep_b200.utils.infer import run_inference
    from wesep_b200.models import get_model
    if family == "spex":
        from oracle.ref_loader import SPEXPLUS_ARGS
        m = get_model("ConvTasNet")(**dict(SPEXPLUS_ARGS, B=64, H=128, X=3, R=2)).to(DEV)
        mk = lambda seed, T: synth.make_batch(2, T=T, Te=T - 300, seed=seed)           # noqa: E731
    else:
        m = get_model("BSRNN")(spk_emb_dim=256, sr=16000, win=512, stride=128, use_spk_transform=False,
                               joint_training=False, feature_dim=16, num_repeat=2, spk_fuse_type="multiply",
                               multi_fuse=False).to(DEV)

        def mk(seed, T):
            b = synth.make_batch(2, T=T, Te=8, seed=seed)
            g = torch.Generator().manual_seed(seed)
            b["spk_embeds"] = torch.randn(2, 256, generator=g)
            return b
    synth.fill_state_dict_(m.state_dict(), seed=21)
    m.train()
    batches = []
    for i, T in enumerate((4000, 5003)):
        b = mk(40 + i, T)
        b["key"] = [f"mix{i}_a", f"mix{i}_b"]
        b["spk"] = ["s1", "s2"]
        batches.append(b)
    lines, sunk = [], {}
    res = run_inference(m, batches, device=DEV, sink=lambda name, w: sunk.__setitem__(name, w), log=lines.append)
    assert m.training                                  # mode restored
    assert res["count"] == 4 and len(lines) == 4 and len(sunk) == 4
    assert lines[0].startswith("Num=1 | Utt=mix0_a | Target speaker=s1 | SI-SNR=")
    # oracle scoring of the same forward
    m.eval()
    tot = 0.0
    with torch.no_grad():
        for bi, b in enumerate(batches):
            out = m(b["wav_mix"].to(DEV), b["spk_embeds"].to(DEV))
            out = out[0] if isinstance(out, (list, tuple)) else out
            ow, os_, od = oscore.score_rows(out.cpu(), b["wav_targets"], b["wav_mix"])
            for r in range(2):
                key, spk, s, d = res["rows"][bi * 2 + r]
                assert abs(s - os_[r]) <= TOL_DB32 and abs(d - od[r]) <= TOL_DB32
                assert np.array_equal(sunk[f"Utt{bi * 2 + r + 1}-{key}-T{spk}.wav"], ow[r])
                tot += od[r]
    assert abs(res["sisnri"] - tot / 4) <= TOL_DB32
    assert res["accept"] == sum(1 for r in res["rows"] if r[3] > 1)
