"""Per-kernel GPU time of ONE pBSRNN train step (CUPTI activity records via torch.profiler) + step timing.
Usage: python tools/kernel_times_bsrnn.py [rows=16] [seconds=4] [out.md]"""
import collections, os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from wesep_b200 import _lib, ops, synth
from wesep_b200.models import get_model
from wesep_b200.utils.optim import FusedClipAdam

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 4.0
out = sys.argv[3] if len(sys.argv) > 3 else None
L = int(16000 * secs)
dev = "cuda"
import bench
m = get_model("BSRNN")(**bench.BSRNN_ARGS)
synth.fill_state_dict_(m.state_dict(), seed=1)
m = m.to(dev).train()
opt = FusedClipAdam(m.parameters(), lr=1e-3, weight_decay=1e-4, clip=5.0)
b = synth.make_batch(n, T=L, Te=8, seed=3, device=dev)
emb = torch.from_numpy(np.random.default_rng(5).standard_normal((n, bench.BSRNN_FBANK_FRAMES, 80)).astype(np.float32)).to(dev)


def step():
    opt.zero_grad()
    est, _ = m(b["wav_mix"], emb)
    losses, _ = ops.sisdr_losses([est], b["wav_targets"])
    losses[0].backward()
    opt.step()
    return losses[0]


for _ in range(2):
    step()
torch.cuda.synchronize()
l0 = _lib.launch_count()
t0 = time.perf_counter()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
step()
t_enq = time.perf_counter() - t0
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e)
launches = _lib.launch_count() - l0
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
small = collections.defaultdict(lambda: [0, 0.0])     # launches shorter than 60 us: the per-band glue
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    name = ev.name
    if name.startswith("Memcpy") or name.startswith("Memset"):
        name = name.split(" ")[0]
    a = agg[name]
    a[0] += 1
    a[1] += ev.time_range.end - ev.time_range.start
    if ev.time_range.end - ev.time_range.start < 60:
        small[name][0] += 1
        small[name][1] += ev.time_range.end - ev.time_range.start
tot = sum(v[1] for v in agg.values())
lines = [f"# one pBSRNN train step, n={n}, {secs:g} s: CUPTI kernel activity (torch.profiler)", "",
         f"step {ms:.1f} ms by CUDA events ({n / ms * 1e3:.1f} utt/s), host enqueue {t_enq * 1e3:.1f} ms, {launches} wesep_b200 launches; "
         f"kernel time sum {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} activities; "
         f"peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB", "",
         "| share | total ms | avg us | count | kernel |", "|---:|---:|---:|---:|---|"]
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    lines.append(f"| {100 * t / tot:.2f} % | {t / 1e3:.2f} | {t / c:.1f} | {c} | `{name[:100]}` |")
st = sum(v[1] for v in small.values())
lines += ["", f"Launches shorter than 60 us: {sum(v[0] for v in small.values())} activities, {st / 1e3:.1f} ms ({100 * st / tot:.1f} % of kernel time):", "",
          "| total ms | avg us | count | kernel |", "|---:|---:|---:|---|"]
for name, (c, t) in sorted(small.items(), key=lambda kv: -kv[1][1])[:14]:
    lines.append(f"| {t / 1e3:.2f} | {t / c:.1f} | {c} | `{name[:100]}` |")
txt = "\n".join(lines)
print(txt)
if out:
    with open(out, "w") as f:
        f.write(txt + "\n")
