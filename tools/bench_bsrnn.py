"""pBSRNN train-step timing (BASELINE config 3: 4 s @ 16 kHz, batch 16, bsrnn.yaml network, joint_training=False):
forward + SISDR loss + backward + clip/Adam on one GPU, CUDA events.  The recurrence is still driven from Python (one GEMM
+ one cell launch per time step), so this is a first-correct-path number, not a tuned one.
Usage: python tools/bench_bsrnn.py [rows=16] [seconds=4] [steps=2]"""
import json, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from wesep_b200 import _lib, ops, synth
from wesep_b200.models import get_model
from wesep_b200.utils.optim import FusedClipAdam

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
L = int(16000 * secs)
dev = "cuda"
m = get_model("BSRNN")(spk_emb_dim=256, sr=16000, win=512, stride=128, feature_dim=128, num_repeat=6, use_spk_transform=False,
                       spk_fuse_type="multiply", multi_fuse=False, joint_training=False)
synth.fill_state_dict_(m.state_dict(), seed=1)
m = m.to(dev).train()
opt = FusedClipAdam(m.parameters(), lr=1e-3, weight_decay=1e-4, clip=5.0)
b = synth.make_batch(n, T=L, Te=8, seed=3, device=dev)
emb = torch.from_numpy(np.random.default_rng(5).standard_normal((n, 256)).astype(np.float32)).to(dev)


def step():
    opt.zero_grad()
    est, _ = m(b["wav_mix"], emb)
    losses, _ = ops.sisdr_losses([est], b["wav_targets"])
    losses[0].backward()
    opt.step()
    return losses[0]


step()
torch.cuda.synchronize()
l0 = _lib.launch_count()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
s.record()
for _ in range(steps):
    loss = step()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / steps
print(json.dumps(dict(metric="utterances/sec pBSRNN train step (%gs@16kHz)" % secs, value=n / (ms * 1e-3), unit="utterances/s",
                      n_gpus=1, rows=n, steps=steps, ms_per_step=ms, wall_ms_per_step=(time.perf_counter() - t0) * 1e3 / steps,
                      gpu_launches_per_step=(_lib.launch_count() - l0) // steps, loss=float(loss),
                      peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
                      note="first correct path: Python-driven recurrence (launch-bound), 3xTF32 GEMMs")))
