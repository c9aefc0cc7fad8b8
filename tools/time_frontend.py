"""Time the data front end / evaluation kernels at the bench shapes (CUDA events, L2 flushed between iterations):
fbank + CMN of n enrollment waves, chunk + SNR mixing of M two-speaker mixtures, peak rule + SI-SNRi of n rows.
Prints one JSON object; algorithmic bytes are the compulsory reads + writes of each call."""
import json
import random
import sys

import torch

sys.path.insert(0, ".")
from wesep_b200.dataset import compute_fbank, snr_mixer  # noqa: E402
from wesep_b200.utils.score import score_batch  # noqa: E402

PEAK = 6568.4  # GB/s, MEASURED_PEAKS.json


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters


def main():
    n, T = 32, 64000
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    wav = torch.randn(n, T, device=dev, generator=g) * 0.1
    res = {}
    ms = timeit(lambda: compute_fbank(wav, dither=1.0, seed=1))
    by = n * T * 4 + 2 * n * 398 * 80 * 4 * 1.5
    res["fbank_cmn"] = dict(rows=n, ms=ms, rows_per_s=n / ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / PEAK)
    M = 16
    lens = [random.randint(T, 4 * T) for _ in range(64)]
    pool = torch.randn(sum(lens), device=dev, generator=g) * 0.1
    start = [sum(lens[:i]) for i in range(64)]
    idx = [[random.randrange(64), random.randrange(64)] for _ in range(M)]
    st = torch.tensor([[start[u] for u in r] for r in idx]).to(dev)
    ul = torch.tensor([[lens[u] for u in r] for r in idx], dtype=torch.int32).to(dev)
    c0 = torch.tensor([[random.randint(0, lens[u] - T) for u in r] for r in idx], dtype=torch.int32).to(dev)
    ms = timeit(lambda: snr_mixer(pool, st, ul, c0, T))
    by = M * T * 4 * (3 * 2 + 3)     # three reads of both sources, three waves written
    res["mix_2spk"] = dict(mixtures=M, ms=ms, rows_per_s=2 * M / ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / PEAK)
    est, ref, mix = (torch.randn(n, T, device=dev, generator=g) for _ in range(3))
    ms = timeit(lambda: score_batch(est, ref, mix))
    by = n * T * 4 * 6               # clone (r+w), peak read, sums: est r+w, ref, mix  -> 7, minus nothing; count 6 compulsory
    res["score"] = dict(rows=n, ms=ms, rows_per_s=n / ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / PEAK)
    try:   # the reference's CPU worker cost for the same feature: torchaudio kaldi.fbank on a float64 4 s wave + CMN
        import time
        import torchaudio.compliance.kaldi as kaldi
        w = (wav[0].cpu().double() * (1 << 15))[None]
        torch.set_num_threads(1)
        t0 = time.perf_counter()
        for _ in range(20):
            m = kaldi.fbank(w, num_mel_bins=80, frame_length=25, frame_shift=10, dither=1.0, sample_frequency=16000,
                            window_type="hamming", use_energy=False)
            m = m - m.mean(0)
        dt = (time.perf_counter() - t0) / 20
        res["cpu_torchaudio_fbank"] = dict(ms_per_row=dt * 1e3, rows_per_s_per_worker=1 / dt)
    except Exception as e:  # noqa: BLE001
        res["cpu_torchaudio_fbank"] = str(e)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
