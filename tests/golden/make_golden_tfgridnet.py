"""Golden fixtures for the TF-GridNet path: outputs / loss / gradient summaries of the REAL reference
``wesep.models.tfgridnet.TFGridNet`` (imported in place from /root/reference through oracle/stubs) on seeded inputs.
Build container only:

    python tests/golden/make_golden_tfgridnet.py [case]

``joint_training=False``: the separator is fed a seeded 256-d embedding.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from oracle import ref_loader  # noqa: E402
from tests.golden.make_golden_dpccn import run_case  # noqa: E402


def main():
    ref_loader.import_reference()
    from wesep.models.tfgridnet import TFGridNet
    base = dict(n_srcs=1, n_fft=128, stride=64, window="hann", n_imics=1, emb_ks=1, emb_hs=1, activation="prelu", eps=1e-5,
                use_spk_transform=False, spk_fuse_type="multiply", joint_training=False)
    # small: 2 blocks, 16 channels, hidden 32, 2 heads x E = 4; 2 rows, T = 33 frames
    run_case(TFGridNet, "tfgridnet_small_train", dict(base, n_layers=2, emb_dim=16, lstm_hidden_units=32, attn_n_head=2,
                                                      attn_approx_qk_dim=260), 2, 2048 + 41, 71, 81)
    # the class default window: emb_ks 4 / emb_hs 1 (unfold + ConvTranspose1d paths), 1 block
    run_case(TFGridNet, "tfgridnet_small_ks4", dict(base, n_layers=1, emb_dim=16, lstm_hidden_units=32, attn_n_head=2,
                                                    attn_approx_qk_dim=260, emb_ks=4, emb_hs=1), 2, 1500, 73, 83)
    # ks == hs == 2: two positions packed per recurrent step (the reshaping Linear path)
    run_case(TFGridNet, "tfgridnet_small_ks2", dict(base, n_layers=1, emb_dim=16, lstm_hidden_units=32, attn_n_head=2,
                                                    attn_approx_qk_dim=260, emb_ks=2, emb_hs=2), 2, 1500, 74, 84)
    # the recipe network (tfgridnet.yaml:44-55: 6 blocks, 128 channels, hidden 192, 4 heads, qk 512) on 0.5 s, one row
    run_case(TFGridNet, "tfgridnet_full_train_05s", dict(base, n_layers=6, emb_dim=128, lstm_hidden_units=192, attn_n_head=4,
                                                         attn_approx_qk_dim=512), 1, 8000, 72, 82, subsample=2)


if __name__ == "__main__":
    main()
