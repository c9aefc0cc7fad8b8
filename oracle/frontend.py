"""ORACLE (test infrastructure only — never imported by the product path).

CPU restatement of the data front end of the reference:
  * get_random_chunk          — wesep/dataset/processor.py:536-573 (chunk start given, not drawn)
  * snr_mixer                 — wesep/dataset/processor.py:276-320
  * compute_fbank + apply_cmvn — wesep/dataset/processor.py:480-535.  The arithmetic of compute_fbank lives in a
    third-party dependency, torchaudio.compliance.kaldi.fbank (torchaudio is unpinned in the reference's
    requirements.txt; 2.11.0 is what this image holds).  Its published algorithm is restated below in fp64 numpy:
    frames (snip_edges) -> + dither -> - frame mean -> pre-emphasis 0.97 -> Hamming -> zero-pad to 512 -> |rfft|^2 ->
    80 triangular mel filters over [20 Hz, Nyquist] -> log(max(., eps_fp32)).

Pinned: tests/golden/frontend_*.npz hold the outputs of the REAL reference functions (wesep.dataset.processor
imported in place with the real torchaudio, tests/golden/make_golden_frontend.py); tests/test_oracle_golden.py checks
this restatement against them.
"""
import math

import numpy as np
import torch


def random_chunk(wav, chunk_len, chunk_start):
    """processor.py:549-571 for one 1-D array with the random draw replaced by `chunk_start`."""
    n = len(wav)
    if n >= chunk_len:
        return wav[chunk_start:chunk_start + chunk_len].copy()
    rep = chunk_len // n + 1
    return np.tile(wav, rep)[:chunk_len]


def snr_mixer(wavs, snrs):
    """processor.py:286-318 on a list of fp32 torch tensors [1, T] (speaker 0 = target) and a list of SNRs in dB
    (entry 0 unused).  Returns (wav_mix [1, T], [wav_spk_s [1, T]]) — same torch CPU ops, same order."""
    wavs = [w.clone() for w in wavs]
    to_mix = [wavs[0]]
    target_energy = torch.sum(to_mix[0] ** 2, dim=-1, keepdim=True)
    for i in range(1, len(wavs)):
        itf = wavs[i]
        energy = torch.sum(itf ** 2, dim=-1, keepdim=True)
        itf *= torch.sqrt(target_energy / energy) * 10 ** (snrs[i] / 20)
        to_mix.append(itf)
    stack = torch.stack(to_mix)
    mix = torch.sum(stack, 0)
    max_amp = max(torch.abs(mix).max().item(), *[x.item() for x in torch.abs(stack).max(dim=-1)[0]])
    scal = 1 / max_amp if max_amp != 0 else 1
    mix = mix * scal
    for w in wavs:
        w *= scal
    return mix, wavs


def mel_banks(num_bins=80, n_fft=512, sample_rate=16000.0, low=20.0, high=0.0):
    """kaldi.get_mel_banks (vtln_warp 1.0): fp32 torch ops as torchaudio runs them, [num_bins, n_fft/2 + 1]."""
    nyq = 0.5 * sample_rate
    if high <= 0.0:
        high += nyq
    width = sample_rate / n_fft
    mlo = 1127.0 * math.log(1.0 + low / 700.0)
    mhi = 1127.0 * math.log(1.0 + high / 700.0)
    d = (mhi - mlo) / (num_bins + 1)
    b = torch.arange(num_bins).unsqueeze(1)
    left, center, right = mlo + b * d, mlo + (b + 1.0) * d, mlo + (b + 2.0) * d
    mel = (1127.0 * torch.log(1.0 + width * torch.arange(n_fft // 2) / 700.0)).unsqueeze(0)
    bins = torch.max(torch.zeros(1), torch.min((mel - left) / (center - left), (right - mel) / (right - center)))
    return torch.nn.functional.pad(bins, (0, 1)).numpy().astype(np.float64)


def fbank(wav, num_mel_bins=80, frame_length=25, frame_shift=10, sample_rate=16000, dither_noise=None, cmn=True):
    """compute_fbank (dither given explicitly as an [m, win] array or None) + apply_cmvn(norm_mean) for one 1-D wave in
    [-1, 1); fp64 throughout, as the reference computes it when the enrollment comes from soundfile (float64)."""
    x = np.asarray(wav, np.float64) * (1 << 15)
    win = int(sample_rate * frame_length * 0.001)
    shift = int(sample_rate * frame_shift * 0.001)
    n_fft = 1 << (win - 1).bit_length()
    if len(x) < win:
        return np.zeros((0, num_mel_bins))
    m = 1 + (len(x) - win) // shift
    fr = np.stack([x[i * shift:i * shift + win] for i in range(m)])
    if dither_noise is not None:
        fr = fr + dither_noise
    fr = fr - fr.mean(axis=1, keepdims=True)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = fr - 0.97 * prev
    n = np.arange(win)
    fr = fr * (0.54 - 0.46 * np.cos(2 * np.pi * n / (win - 1)))
    fr = np.pad(fr, ((0, 0), (0, n_fft - win)))
    power = np.abs(np.fft.rfft(fr, axis=1)) ** 2
    mel = power @ mel_banks(num_mel_bins, n_fft, float(sample_rate)).T
    out = np.log(np.maximum(mel, float(np.finfo(np.float32).eps)))
    if cmn:
        out = out - out.mean(axis=0)
    return out
