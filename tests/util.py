"""Shared helpers for the parity tests (tests may import oracle/)."""
import json
import os

import numpy as np
import torch

from oracle import spexplus as ospex
from wesep_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    return z, meta


def cfg_from_args(args):
    cfg = dict(ospex.DEFAULT_CFG)
    for k in ("N", "L", "B", "H", "P", "X", "R", "spk_emb_dim", "spk_fuse_type", "multi_task", "spksInTrain"):
        if k in args:
            cfg[k] = args[k]
    return cfg


def fixture_inputs(meta, dtype=torch.float32, device="cpu"):
    cfg = cfg_from_args(meta["args"])
    sd = ospex.make_state_dict(cfg, dtype=dtype, device="cpu")
    synth.fill_state_dict_(sd, seed=meta["wseed"])
    sd = {k: v.to(device) for k, v in sd.items()}
    batch = synth.make_batch(meta["n"], T=meta["T"], Te=meta["Te"], seed=meta["dseed"], dtype=dtype)
    batch = {k: v.to(device) for k, v in batch.items()}
    return cfg, sd, batch


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
