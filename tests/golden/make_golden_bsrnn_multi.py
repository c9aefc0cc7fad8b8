"""Golden fixture for BSRNN_Multi: outputs / loss / gradient summaries of the REAL reference
``wesep.models.bsrnn_multi_optim.BSRNN_Multi`` (imported in place from /root/reference through oracle/stubs) on seeded inputs.
Build container only:  python tests/golden/make_golden_bsrnn_multi.py

The speaker encoder is a two-parameter stand-in (mean over frames of the squared features + Linear(80, 256); tests/util.SqTiny)
patched into the reference module: wespeaker itself is an external package; the test swaps the same module into the CUDA model.
(The plain mean of oracle/stubs would be degenerate here: the features are mean-normalised over frames.)  Only PARAMETERS are filled from the seed
(the pre-emphasis / window / mel-filterbank buffers keep their constructed values).  Loss as in bsrnn_multi_optim.yaml:34-37:
0.4 SISDR(s) + 0.6 SISDR(self_s).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader, losses as olosses  # noqa: E402
from wesep_b200 import synth  # noqa: E402

ARGS = dict(sr=16000, win=512, stride=128, feature_dim=16, num_repeat=2, spk_fuse_type="multiply", use_spk_transform=False,
            multi_fuse=False, joint_training=True, spk_model="ResNet18", spk_model_init=False,
            spk_args=dict(feat_dim=80, embed_dim=256), spk_emb_dim=256, spk_model_freeze=False, spk_feat=False,
            feat_type="consistent", multi_task=False)
N, L, TE, WSEED, DSEED = 2, 4000, 5000, 91, 92


def params_only(sd):
    return {k: v for k, v in sd.items() if not k.startswith(("preEmphasis", "spk_encoder"))}


def main():
    ref_loader.import_reference()
    import wesep.models.bsrnn_multi_optim as M
    from tests.util import SqTiny
    M.get_speaker_model = lambda name: SqTiny
    BSRNN_Multi = M.BSRNN_Multi
    torch.manual_seed(0)
    m = BSRNN_Multi(**ARGS)
    synth.fill_state_dict_(params_only(m.state_dict()), seed=WSEED)
    b = synth.make_batch(N, T=L, Te=TE, seed=DSEED)
    m.train(True)
    s, self_s, _, _ = m(b["wav_mix"], b["spk_embeds"])
    loss = 0.4 * olosses.sisdr_loss(s, b["wav_targets"]) + 0.6 * olosses.sisdr_loss(self_s, b["wav_targets"])
    loss.backward()
    fix = dict(out0=s.detach().numpy(), out1=self_s.detach().numpy(), loss=np.float64(loss.item()),
               sisdr_rows0=olosses.sisdr_per_row(s.detach().double(), b["wav_targets"].double()).numpy(),
               sisdr_rows1=olosses.sisdr_per_row(self_s.detach().double(), b["wav_targets"].double()).numpy())
    for k, p in m.named_parameters():
        g64 = p.grad.double()
        fix["gnorm/" + k] = np.float64(g64.norm().item())
        fix["ghead/" + k] = p.grad.detach().reshape(-1)[:256].numpy().copy()
    with torch.no_grad():
        m.eval()
        out = m(b["wav_mix"], b["spk_embeds"])
        assert len(out) == 2
        fix["eval_out0"] = out[0].numpy()
    fix["meta"] = np.array(json.dumps(dict(args=ARGS, n=N, L=L, Te=TE, wseed=WSEED, dseed=DSEED)))
    np.savez_compressed(os.path.join(HERE, "bsrnn_multi_small.npz"), **fix)
    print("wrote bsrnn_multi_small", float(loss), fix["sisdr_rows0"], fix["sisdr_rows1"])


if __name__ == "__main__":
    main()
