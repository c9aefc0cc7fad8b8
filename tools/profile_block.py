"""Short workload for `ncu --set full`: a few forward/backward passes of one Spex+ TCN block
(B=256, H=512, K=6399 frames) at n rows.  Usage: python tools/profile_block.py [n] [dil] [reps]"""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from wesep_b200 import ops, synth  # noqa: E402
from wesep_b200.modules.tasnet.convs import Conv1DBlock  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dil = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = "cuda"
blk = Conv1DBlock(256, 512, 3, dil, "gLN", False, False)
synth.fill_state_dict_(blk.state_dict(), seed=1)
blk = blk.to(dev)
x = ops.new_act(n, 256, 6399, dev)
x.normal_()
x.requires_grad_(True)
g = ops.new_act(n, 256, 6399, dev)
g.normal_()
for _ in range(reps):
    y = blk(x)
    torch.autograd.grad(y, [x] + list(blk.parameters()), g)
torch.cuda.synchronize()
print("done")
