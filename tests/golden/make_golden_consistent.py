"""Generate tests/golden/consistent_feats.npz: the in-model "consistent" enrollment features of the reference
(wesep/models/bsrnn.py:345-351) computed with the REAL reference PreEmphasis (wesep.modules.common.speaker) and the real
torchaudio.transforms.MelSpectrogram, for the (n_fft, hop) pairs of pBSRNN / pDPCCN (512, 128) and TF-GridNet (128, 64).
Build container only:  python tests/golden/make_golden_consistent.py"""
import os
import sys

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402
from tests.util import frontend_waves  # noqa: E402

CASES = [("w512", 512, 128, [24000, 24000]), ("w128", 128, 64, [9001])]


def main():
    ref_loader.import_reference()
    from wesep.modules.common.speaker import PreEmphasis
    fix = {}
    for name, win, hop, lens in CASES:
        x = torch.from_numpy(np.stack(frontend_waves(33, lens)))
        pre = PreEmphasis()
        enc = torchaudio.transforms.MelSpectrogram(sample_rate=16000, n_fft=win, win_length=win, hop_length=hop, f_min=20,
                                                   window_fn=torch.hamming_window, n_mels=80)
        with torch.no_grad():
            y = enc(pre(x)) + 1e-8
            y = y.log()
            y = y - torch.mean(y, dim=-1, keepdim=True)
            y = y.permute(0, 2, 1)
        fix[name] = y.numpy()
        print(name, y.shape, float(y.abs().max()))
    np.savez_compressed(os.path.join(HERE, "consistent_feats.npz"), **fix)


if __name__ == "__main__":
    main()
