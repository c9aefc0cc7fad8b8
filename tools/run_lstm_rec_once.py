"""One forward + one backward launch of the recurrence kernels (for ncu): python tools/run_lstm_rec_once.py S Q Hd seqs"""
import sys

import torch

sys.path.insert(0, ".")
from wesep_b200 import _lib, ops   # noqa: E402
from wesep_b200.ops import _args, _stream   # noqa: E402

S, Q, Hd, seqs = (int(v) for v in sys.argv[1:5])
dev = "cuda"
g = torch.Generator().manual_seed(0)
G = ops.new_act(S, 8 * Hd, Q, dev)
G.copy_(torch.randn(S, 8 * Hd, Q, generator=g))
H = ops.new_act(S, 2 * Hd, Q, dev)
Cs = ops.new_act(S, 2 * Hd, Q, dev)
dH = ops.new_act(S, 2 * Hd, Q, dev)
dH.copy_(torch.randn(S, 2 * Hd, Q, generator=g))
W = [(torch.randn(4 * Hd, Hd, generator=g) * Hd ** -0.5).to(dev) for _ in range(2)]
a = _args("WesepLstmRecArgs", S=S, Q=Q, Hd=Hd, ld=G.stride(1), bsG=G.stride(0), bsH=H.stride(0), G=G, H=H, C=Cs,
          Whh_f=W[0], Whh_r=W[1], dH=dH, seqs_per_cluster=seqs)
for _ in range(2):
    _lib.call("wesep_b200_lstm_rec_fwd", a, _stream())
    _lib.call("wesep_b200_lstm_rec_bwd", a, _stream())
torch.cuda.synchronize()
print("ok")
