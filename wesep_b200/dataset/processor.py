"""Batched GPU versions of the per-sample processors the reference runs in DataLoader workers
(wesep/dataset/processor.py): `random_chunk` + `mix_speakers` + `snr_mixer` (online mixing) and
`compute_fbank` + `apply_cmvn` (enrollment features).  At several hundred rows/s per GPU the CPU
workers (tar decode aside) become the bottleneck; here the utterance pool is
resident in HBM, the host only draws the random indices the reference draws with `random`, and
one C-ABI call per batch does the arithmetic.  There is no CPU fallback.
"""
import math
import random

import torch

from wesep_b200 import _lib, ops


# --------------------------------------------------------------------------- mixing
def snr_mixer(pool, start, ulen, chunk0, chunk_len, snr_db=None):
    """processor.py:276-320 for M mixtures of S speakers at once, with the chunk gather of
    get_random_chunk (processor.py:536-573) folded in.

    pool   : 1-D fp32 CUDA tensor, the source utterances back to back
    start  : [M, S] int64 (CPU or CUDA) — index in `pool` of each utterance
    ulen   : [M, S] int32 — utterance lengths; shorter than `chunk_len` tiles (processor.py:562-570)
    chunk0 : [M, S] int32 — chunk start inside the utterance (host-drawn random.randint)
    snr_db : [M, S] fp32 or None (0 dB, `use_random_snr=False`)
    Returns (wav_mix [M, T], wav_spk [S, M, T]) — wav_spk[s] is the reference sample's `wav_spk{s+1}`."""
    ops._check_cuda(pool)
    if pool.dtype != torch.float32 or pool.dim() != 1 or not pool.is_contiguous():
        raise RuntimeError("snr_mixer: pool must be a contiguous 1-D fp32 tensor")
    dev = pool.device
    start = torch.as_tensor(start, dtype=torch.int64)
    M, S = start.shape
    ulen = torch.as_tensor(ulen, dtype=torch.int32)
    chunk0 = torch.as_tensor(chunk0, dtype=torch.int32)
    if ulen.shape != (M, S) or chunk0.shape != (M, S):
        raise RuntimeError("snr_mixer: start / ulen / chunk0 must all be [M, S]")
    T = int(chunk_len)
    if not (start.is_cuda and ulen.is_cuda and chunk0.is_cuda):   # host tables: validate before the upload
        st, ul, c0 = start.cpu(), ulen.cpu().long(), chunk0.cpu().long()
        if int(ul.min()) <= 0:
            raise RuntimeError("snr_mixer: empty utterance")
        last = torch.where(ul >= T, c0 + T, ul)
        if int(c0.min()) < 0 or bool(((st < 0) | (st + last > pool.numel())).any()):
            raise RuntimeError("snr_mixer: chunk outside the pool")
    start, ulen, chunk0 = (t.to(dev, non_blocking=True).contiguous() for t in (start, ulen, chunk0))
    if snr_db is not None:
        snr_db = torch.as_tensor(snr_db, dtype=torch.float32).to(dev, non_blocking=True).contiguous()
        if snr_db.shape != (M, S):
            raise RuntimeError("snr_mixer: snr_db must be [M, S]")
    ld = ops.ceil4(T)
    mix = torch.empty((M, ld), dtype=torch.float32, device=dev)
    spk = torch.empty((S, M, ld), dtype=torch.float32, device=dev)
    ws = torch.empty(_lib.lib().wesep_b200_mix_ws_bytes(M) // 8, dtype=torch.float64, device=dev)
    a = ops._args("WesepMixArgs", M=M, S=S, T=T, pool=pool, start=start, ulen=ulen, chunk0=chunk0, snr_db=snr_db,
                  mix=mix, spk=spk, ld=ld, ws=ws)
    _lib.call("wesep_b200_mix", a, ops._stream())
    return mix[:, :T], spk[:, :, :T]


# --------------------------------------------------------------------------- fbank
_TABLES = {}


def _mel_scale(f):
    return 1127.0 * torch.log(1.0 + f / 700.0)


def fbank_tables(device, sample_rate=16000, frame_length=25, frame_shift=10, num_mel_bins=80, low_freq=20.0, high_freq=0.0):
    """Window + mel filterbank as torchaudio.compliance.kaldi builds them (fp32 CPU ops: hamming_window(periodic=False),
    get_mel_banks with vtln_warp 1.0), uploaded once per device."""
    key = (str(device), sample_rate, frame_length, frame_shift, num_mel_bins, low_freq, high_freq)
    t = _TABLES.get(key)
    if t is not None:
        return t
    win = int(sample_rate * frame_length * 0.001)
    shift = int(sample_rate * frame_shift * 0.001)
    n_fft = 1 << (win - 1).bit_length()                       # round_to_power_of_two
    window = torch.hamming_window(win, periodic=False, alpha=0.54, beta=0.46, dtype=torch.float32)
    nyq = 0.5 * sample_rate
    hi = high_freq + nyq if high_freq <= 0.0 else high_freq
    if not (0.0 <= low_freq < nyq and 0.0 < hi <= nyq and low_freq < hi):
        raise RuntimeError("fbank: bad low/high frequency")
    bin_width = sample_rate / n_fft
    mlo = 1127.0 * math.log(1.0 + low_freq / 700.0)
    mhi = 1127.0 * math.log(1.0 + hi / 700.0)
    delta = (mhi - mlo) / (num_mel_bins + 1)
    b = torch.arange(num_mel_bins).unsqueeze(1)
    left, center, right = mlo + b * delta, mlo + (b + 1.0) * delta, mlo + (b + 2.0) * delta
    mel = _mel_scale(bin_width * torch.arange(n_fft // 2)).unsqueeze(0)
    up = (mel - left) / (center - left)
    down = (right - mel) / (right - center)
    bins = torch.max(torch.zeros(1), torch.min(up, down))
    bins = torch.nn.functional.pad(bins, (0, 1)).float().contiguous()          # [num_mel, n_fft/2 + 1]
    nz = bins > 0
    lo = torch.where(nz.any(1), nz.float().argmax(1), torch.zeros(num_mel_bins, dtype=torch.long))
    hi_ = torch.where(nz.any(1), bins.shape[1] - nz.flip(1).float().argmax(1), torch.zeros(num_mel_bins, dtype=torch.long))
    t = dict(win=win, shift=shift, n_fft=n_fft, window=window.to(device), mel=bins.to(device),
             mel_lo=lo.int().to(device), mel_hi=hi_.int().to(device))
    _TABLES[key] = t
    return t


def num_frames(length, win=400, shift=160):
    return 1 + (length - win) // shift if length >= win else 0


def compute_fbank(wav, lengths=None, num_mel_bins=80, frame_length=25, frame_shift=10, dither=1.0, sample_rate=16000,
                  apply_cmvn=True, seed=None):
    """compute_fbank + apply_cmvn (processor.py:480-535) for a batch of enrollment waves.

    wav [n, T] CUDA fp32 in [-1, 1) (scaled by 2^15 inside, processor.py:498); lengths: valid samples per row.
    Returns [n, max_frames, num_mel_bins] fp32, frames past a row's own count zero ("max" collate padding).
    `dither` draws from this module's own counter-based generator (seeded from Python `random` unless `seed` is given),
    so dithered output agrees with the reference in distribution only; dither=0 is deterministic."""
    ops._check_cuda(wav)
    if wav.dim() == 1:
        wav = wav[None]
    wav = wav.float()
    if wav.stride(1) != 1:
        wav = wav.contiguous()
    n, T = wav.shape
    dev = wav.device
    tb = fbank_tables(dev, sample_rate, frame_length, frame_shift, num_mel_bins)
    lens = None
    longest = T
    if lengths is not None:
        host = torch.as_tensor(lengths, dtype=torch.int32)
        if host.numel() != n:
            raise RuntimeError("compute_fbank: one length per row")
        longest = min(int(host.max()), T) if not host.is_cuda else T
        lens = host.to(dev, non_blocking=True)
    mf = num_frames(longest, tb["win"], tb["shift"])
    if mf == 0:
        return torch.empty((n, 0, num_mel_bins), dtype=torch.float32, device=dev)    # kaldi.fbank returns an empty tensor
    out = torch.empty((n, mf, num_mel_bins), dtype=torch.float32, device=dev)
    if seed is None:
        seed = random.getrandbits(63) if dither != 0.0 else 0
    a = ops._args("WesepFbankArgs", n=n, T=T, wav=wav, ld_wav=wav.stride(0), len=lens, frame_len=tb["win"],
                  frame_shift=tb["shift"], n_fft=tb["n_fft"], num_mel=num_mel_bins, scale=float(1 << 15), dither=float(dither),
                  seed=int(seed), preemph=0.97, remove_dc=1, window=tb["window"], mel=tb["mel"], mel_lo=tb["mel_lo"],
                  mel_hi=tb["mel_hi"], log_floor=float(torch.finfo(torch.float32).eps), out=out, bs_out=out.stride(0),
                  max_frames=mf, cmn=int(bool(apply_cmvn)))
    _lib.call("wesep_b200_fbank", a, ops._stream())
    return out


# --------------------------------------------------------------------------- online mixing from a resident pool
class OnlineMixer:
    """The `online_mix=True` branch of the reference Dataset (wesep/dataset/dataset.py:336-353:
    random_chunk -> mix_speakers -> snr_mixer) with the utterance pool resident in HBM.

    utterances : list of (key, spk, 1-D float tensor); uploaded once.
    Every call to `sample(n_mix)` draws, with Python's `random` as the reference does, a target utterance, S-1
    interferers of other speakers (processor.py:232-241), one chunk offset per utterance (processor.py:552-553) and,
    if `use_random_snr`, an SNR in [-10, 10] dB (processor.py:296-297); the gather, scaling, summation and peak
    normalisation run in one wesep_b200_mix call.  `rows()` lays the result out as tse_collate_fn does
    (one row per speaker of every mixture: the mixture repeated, that speaker's scaled source as the target)."""

    def __init__(self, utterances, device, chunk_len=48000, num_speakers=2, use_random_snr=False):
        if not 1 <= num_speakers <= 4:
            raise RuntimeError("OnlineMixer: 1..4 speakers")
        self.keys = [u[0] for u in utterances]
        self.spks = [u[1] for u in utterances]
        if num_speakers > 1 and len(set(self.spks)) < 2:
            raise RuntimeError("OnlineMixer: mixing needs at least two speakers in the pool")
        waves = [torch.as_tensor(u[2], dtype=torch.float32).reshape(-1) for u in utterances]
        self.ulen = [w.numel() for w in waves]
        if min(self.ulen) <= 0:
            raise RuntimeError("OnlineMixer: empty utterance")
        self.start = [0]
        for ln in self.ulen[:-1]:
            self.start.append(self.start[-1] + ln)
        self.pool = torch.cat(waves).to(device)
        self.chunk_len, self.S, self.use_random_snr = int(chunk_len), int(num_speakers), bool(use_random_snr)
        self._order = []

    def _chunk0(self, u):
        ln = self.ulen[u]
        return random.randint(0, ln - self.chunk_len) if ln >= self.chunk_len else 0

    def draw(self, n_mix):
        """Host side only: the index tables of one batch (lists of [M][S])."""
        idx, c0, snr = [], [], []
        N = len(self.keys)
        for _ in range(n_mix):
            if not self._order:           # mix_speakers walks its shuffled buffer: every utterance is a target once per pass
                self._order = list(range(N))
                random.shuffle(self._order)
            t = self._order.pop()
            row = [t]
            while len(row) < self.S:
                i = random.randrange(N)
                while self.spks[i] == self.spks[t]:
                    i = random.randrange(N)
                row.append(i)
            idx.append(row)
            c0.append([self._chunk0(u) for u in row])
            snr.append([0.0] + [random.uniform(-10, 10) if self.use_random_snr else 0.0 for _ in row[1:]])
        return idx, c0, snr

    def sample(self, n_mix, tables=None):
        idx, c0, snr = tables if tables is not None else self.draw(n_mix)
        start = torch.tensor([[self.start[u] for u in row] for row in idx], dtype=torch.int64)
        ulen = torch.tensor([[self.ulen[u] for u in row] for row in idx], dtype=torch.int32)
        mix, spk = snr_mixer(self.pool, start, ulen, torch.tensor(c0, dtype=torch.int32), self.chunk_len,
                             torch.tensor(snr, dtype=torch.float32) if self.use_random_snr else None)
        keys = ["mix_" + "_".join(self.keys[u] for u in row) for row in idx]
        return dict(wav_mix=mix, wav_spk=spk, idx=idx, key=keys)

    def rows(self, s):
        """tse_collate_fn layout (dataset.py:198-226): rows ordered mixture-major, speaker-minor."""
        mix, spk = s["wav_mix"], s["wav_spk"]
        M, T = mix.shape
        S = spk.shape[0]
        wav_mix = mix[:, None, :].expand(M, S, T).reshape(M * S, T)
        wav_targets = spk.permute(1, 0, 2).reshape(M * S, T)
        spks = [self.spks[u] for row in s["idx"] for u in row]
        keys = [k for k in s["key"] for _ in range(S)]
        return dict(wav_mix=wav_mix, wav_targets=wav_targets, spk=spks, key=keys)
