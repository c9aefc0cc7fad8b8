"""Generate tests/golden/frontend_mix.npz and frontend_fbank.npz with the REAL reference processors
(wesep.dataset.processor.get_random_chunk / snr_mixer / compute_fbank / apply_cmvn imported in place from
/root/reference through oracle/stubs, with the real torchaudio of this image).  Build container only:

    python tests/golden/make_golden_frontend.py

Inputs are regenerated on both sides by tests.util.frontend_waves(seed, lengths).
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402
from tests.util import FBANK_CASES, MIX_CASES, frontend_waves  # noqa: E402


def main():
    ref_loader.import_reference()
    import wesep.dataset.processor as P
    fix = {}
    for name, seed, lens, T, use_snr in MIX_CASES:
        waves = frontend_waves(seed, lens)
        chunks, c0s = [], []
        for i, w in enumerate(waves):
            random.seed(seed * 100 + i)
            c0s.append(random.randint(0, len(w) - T) if len(w) >= T else 0)
            random.seed(seed * 100 + i)                                   # same draw inside the reference
            chunks.append(P.get_random_chunk([torch.from_numpy(w)[None]], T)[0])
        sample = {"num_speaker": len(waves)}
        for i, c in enumerate(chunks):
            sample[f"wav_spk{i + 1}"] = c
        random.seed(seed + 7)
        snrs = [0.0] + [random.uniform(-10, 10) if use_snr else 0.0 for _ in waves[1:]]
        random.seed(seed + 7)
        out = next(P.snr_mixer(iter([sample]), use_random_snr=use_snr))
        fix[name + "/c0"] = np.array(c0s, np.int64)
        fix[name + "/snr"] = np.array(snrs, np.float64)
        fix[name + "/mix"] = out["wav_mix"].numpy()[0]
        for i in range(len(waves)):
            fix[name + f"/spk{i}"] = out[f"wav_spk{i + 1}"].numpy()[0]
        print(name, c0s, snrs, float(np.abs(fix[name + "/mix"]).max()))
    np.savez_compressed(os.path.join(HERE, "frontend_mix.npz"), **fix)

    fix = {}
    for name, seed, n_samp, dtype in FBANK_CASES:
        w = frontend_waves(seed, [n_samp])[0].astype(dtype)
        sample = {"sample_rate": 16000, "embed_spk1": w[None]}
        s = next(P.apply_cmvn(P.compute_fbank(iter([sample]), num_mel_bins=80, frame_length=25, frame_shift=10, dither=0.0)))
        mat = s["embed_spk1"][0]
        fix[name] = np.asarray(mat, np.float64)
        print(name, mat.shape, mat.dtype, float(np.abs(mat).max()))
    np.savez_compressed(os.path.join(HERE, "frontend_fbank.npz"), **fix)


if __name__ == "__main__":
    main()
