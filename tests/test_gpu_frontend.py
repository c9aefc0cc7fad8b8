"""GPU: device-side data front end — chunk + SNR mixing and Kaldi fbank + CMN kernels vs the golden
outputs of the REAL reference processors (tests/golden/frontend_*.npz) and vs oracle/frontend.py."""
import random

import numpy as np
import pytest
import torch

from oracle import frontend as ofe
from tests.util import FBANK_CASES, MIX_CASES, frontend_waves

pytestmark = pytest.mark.gpu
DEV = "cuda"
# mixing is fp32 elementwise work on top of two energy sums: the kernel's sums are fp64-accumulated (correctly rounded
# to fp32), torch's are fp32 cascade sums, so the gain can differ by an ulp or two -> relative 5e-7 on the waves
TOL_MIX = 5e-7
TOL_FBANK = 2e-4   # abs, log-mel after CMN; fp32 kernel vs the fp64 reference (torchaudio's own fp32 run is 4e-5 off)


def _pool(waves):
    start, s = [], 0
    for w in waves:
        start.append(s)
        s += len(w)
    return torch.from_numpy(np.concatenate(waves)).to(DEV), start


@pytest.mark.parametrize("case", MIX_CASES, ids=[c[0] for c in MIX_CASES])
def test_mix_golden(case):
    from wesep_b200.dataset import snr_mixer
    name, seed, lens, T, _ = case
    z = np.load("tests/golden/frontend_mix.npz")
    waves = frontend_waves(seed, lens)
    pool, start = _pool(waves)
    mix, spk = snr_mixer(pool, [start], [lens], [list(z[name + "/c0"])], T, snr_db=[list(z[name + "/snr"])])
    ref = z[name + "/mix"]
    assert mix.shape == (1, T) and spk.shape == (len(lens), 1, T)
    assert np.abs(mix[0].cpu().numpy() - ref).max() <= TOL_MIX * np.abs(ref).max(), name
    for i in range(len(lens)):
        r = z[name + f"/spk{i}"]
        assert np.abs(spk[i, 0].cpu().numpy() - r).max() <= TOL_MIX * np.abs(r).max(), (name, i)
    assert abs(max(float(mix.abs().max()), float(spk.abs().max())) - 1.0) <= 1e-6     # peak-normalised


def test_mix_batch_vs_oracle():
    """Many mixtures in one call (M = 37, S = 2, recipe chunk of 3 s), every row against the oracle; mix == sum of the
    scaled sources exactly (same fp32 adds)."""
    from wesep_b200.dataset import snr_mixer
    rng = random.Random(11)
    lens = [rng.randint(20000, 90000) for _ in range(12)]
    waves = frontend_waves(21, lens)
    pool, start = _pool(waves)
    T, M = 48000, 37
    idx = [[rng.randrange(12), rng.randrange(12)] for _ in range(M)]
    c0 = [[rng.randint(0, lens[u] - T) if lens[u] >= T else 0 for u in row] for row in idx]
    snr = [[0.0, rng.uniform(-10, 10)] for _ in range(M)]
    mix, spk = snr_mixer(pool, [[start[u] for u in r] for r in idx], [[lens[u] for u in r] for r in idx], c0, T, snr)
    mix, spk = mix.cpu(), spk.cpu()
    for m in range(M):
        chunks = [torch.from_numpy(ofe.random_chunk(waves[u], T, c))[None] for u, c in zip(idx[m], c0[m])]
        omix, ospk = ofe.snr_mixer(chunks, snr[m])
        assert (mix[m] - omix[0]).abs().max() <= TOL_MIX * float(omix.abs().max()), m
        for s in range(2):
            assert (spk[s, m] - ospk[s][0]).abs().max() <= TOL_MIX * float(ospk[s].abs().max()), (m, s)
    # the peak rule is applied after the sum: mix / scal == spk0 / scal + spk1 / scal only up to rounding, but the
    # unscaled identity holds inside the kernel; check the scaled one to 1 ulp of the peak
    assert (mix - (spk[0] + spk[1])).abs().max() <= 2e-7


def test_mix_errors():
    from wesep_b200.dataset import snr_mixer
    pool = torch.zeros(1000, device=DEV)
    with pytest.raises(RuntimeError):
        snr_mixer(pool, [[0, 900]], [[500, 500]], [[0, 0]], 400)       # second chunk runs past the pool
    with pytest.raises(RuntimeError):
        snr_mixer(pool, [[0, 10]], [[500, 0]], [[0, 0]], 400)          # empty utterance
    with pytest.raises(RuntimeError):
        snr_mixer(pool.cpu(), [[0, 10]], [[500, 100]], [[0, 0]], 400)


@pytest.mark.parametrize("case", FBANK_CASES, ids=[c[0] for c in FBANK_CASES])
def test_fbank_golden(case):
    from wesep_b200.dataset import compute_fbank
    name, seed, n_samp, _ = case
    ref = np.load("tests/golden/frontend_fbank.npz")[name]
    w = torch.from_numpy(frontend_waves(seed, [n_samp])[0]).to(DEV)
    got = compute_fbank(w, dither=0.0)[0].cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= TOL_FBANK, (name, np.abs(got - ref).max())


def test_fbank_batch_ragged_and_no_cmn():
    """Ragged batch: frames past each row's own count are zero, CMN uses the row's own frames; cmn off == raw log-mel."""
    from wesep_b200.dataset import compute_fbank
    lens = [64000, 16400, 399, 400, 31999, 7000]
    waves = frontend_waves(31, lens)
    T = max(lens)
    batch = torch.zeros(len(lens), T)
    for i, w in enumerate(waves):
        batch[i, :lens[i]] = torch.from_numpy(w)
        batch[i, lens[i]:] = 0.5                                          # garbage past the valid length must not leak
    for cmn in (True, False):
        got = compute_fbank(batch.to(DEV), lengths=lens, dither=0.0, apply_cmvn=cmn).cpu().numpy()
        assert got.shape == (len(lens), 398, 80)
        for i, w in enumerate(waves):
            ref = ofe.fbank(w, cmn=cmn)
            m = ref.shape[0]
            if m:
                assert np.abs(got[i, :m] - ref).max() <= TOL_FBANK, (i, cmn, np.abs(got[i, :m] - ref).max())
            assert not got[i, m:].any()


def test_fbank_dither_statistics():
    """dither=1.0 (bsrnn.yaml:20): N(0,1) on the int16 scale.  On a silent wave the frames are pure dither, so the
    log-mel energies are those of white noise: compare their mean over many frames with the oracle fed numpy noise;
    two seeds differ, one seed repeats."""
    from wesep_b200.dataset import compute_fbank
    w = torch.zeros(1, 160000, device=DEV)
    a = compute_fbank(w, dither=1.0, apply_cmvn=False, seed=5)
    b = compute_fbank(w, dither=1.0, apply_cmvn=False, seed=5)
    c = compute_fbank(w, dither=1.0, apply_cmvn=False, seed=6)
    assert torch.equal(a, b) and not torch.equal(a, c)
    m = a.shape[1]
    rng = np.random.default_rng(0)
    ref = ofe.fbank(np.zeros(160000), dither_noise=rng.standard_normal((m, 400)), cmn=False)
    got = a[0].cpu().numpy()
    # per-bin mean over ~1000 frames: the standard error of a log-chi2 mean is <= 0.05 for the narrowest filters
    assert np.abs(got.mean(0) - ref.mean(0)).max() <= 0.15
    assert np.abs(got.std(0) - ref.std(0)).max() <= 0.15
    # the noise is white across samples and frames: lag-1 correlation of the frame energies ~ 0
    e = got.sum(1) - got.sum(1).mean()
    assert abs(float((e[1:] * e[:-1]).mean() / (e * e).mean())) <= 0.15


def test_online_mixer_rows():
    """OnlineMixer: interferers come from other speakers, rows follow tse_collate_fn (mixture-major, speaker-minor),
    and the batch equals the oracle run on the same drawn tables."""
    from wesep_b200.dataset import OnlineMixer
    lens = [30000, 52000, 20000, 61000, 48000, 70000]
    waves = frontend_waves(41, lens)
    utts = [(f"u{i}", f"s{i % 3}", torch.from_numpy(w)) for i, w in enumerate(waves)]
    random.seed(3)
    mx = OnlineMixer(utts, DEV, chunk_len=48000, num_speakers=2, use_random_snr=True)
    tables = mx.draw(5)
    s = mx.sample(5, tables)
    idx, c0, snr = tables
    for row in idx:
        assert utts[row[0]][1] != utts[row[1]][1]
    rows = mx.rows(s)
    assert rows["wav_mix"].shape == (10, 48000) and rows["wav_targets"].shape == (10, 48000)
    assert torch.equal(rows["wav_mix"][0], rows["wav_mix"][1]) and rows["key"][0] == rows["key"][1]
    assert rows["spk"][:2] == [utts[idx[0][0]][1], utts[idx[0][1]][1]]
    for m in range(5):
        chunks = [torch.from_numpy(ofe.random_chunk(waves[u], 48000, c))[None] for u, c in zip(idx[m], c0[m])]
        omix, ospk = ofe.snr_mixer(chunks, snr[m])
        assert (rows["wav_mix"][2 * m].cpu() - omix[0]).abs().max() <= TOL_MIX * float(omix.abs().max())
        assert (rows["wav_targets"][2 * m + 1].cpu() - ospk[1][0]).abs().max() <= TOL_MIX * float(ospk[1].abs().max())
    # every utterance is a target once per pass over the pool (mix_speakers iterates its buffer)
    random.seed(4)
    seen = [r[0] for r in OnlineMixer(utts, DEV, chunk_len=48000).draw(6)[0]]
    assert sorted(seen) == list(range(6))


@pytest.mark.parametrize("name,win,hop,lens", [("w512", 512, 128, [24000, 24000]), ("w128", 128, 64, [9001])])
def test_consistent_features_golden(name, win, hop, lens):
    """In-model enrollment features (spk_feat False, feat_type consistent; bsrnn.py:345-351) vs the REAL reference PreEmphasis +
    torchaudio MelSpectrogram (tests/golden/consistent_feats.npz).  Log-mel after mean removal, fp32 on both sides: the
    low-energy TF-GridNet bins (80 mel filters over 65 frequency bins leave empty filters = log(1e-8)) agree exactly."""
    from wesep_b200.modules.speaker.consistent import MelSpectrogram, PreEmphasis, consistent_features
    ref = np.load("tests/golden/consistent_feats.npz")[name]
    x = torch.from_numpy(np.stack(frontend_waves(33, lens))).to(DEV)
    pre, enc = PreEmphasis().to(DEV), MelSpectrogram(16000, win, hop, 20.0, 80).to(DEV)
    got = consistent_features(x, pre, enc).cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-3, np.abs(got - ref).max()
    assert np.sqrt(((got - ref) ** 2).mean()) <= 1e-4
