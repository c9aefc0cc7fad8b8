"""Per-kernel GPU time of ONE Spex+ train step (n rows), via CUPTI activity records (torch.profiler).

Much cheaper than an ncu launch list (the step runs at full speed); durations are warm and overlapped-as-run,
so they complement (not replace) the ncu pass.  Also reports the host-side enqueue time of a step.
Usage: python tools/kernel_times.py [rows=32] [out.md]
"""
import os, sys, time, collections
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from wesep_b200 import synth
from wesep_b200.models import get_model
from wesep_b200.utils.executor import train_step
from wesep_b200.utils.optim import FusedClipAdam

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 32
out = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = get_model("ConvTasNet")(**bench.SPEX_ARGS).to(dev).train()
opt = FusedClipAdam(model.parameters(), lr=1e-3, weight_decay=1e-4, clip=5.0)
batch = {k: v.to(dev) for k, v in synth.make_batch(rows, T=bench.T_SAMPLES, Te=bench.T_SAMPLES, seed=1234, pin=False).items()}
for _ in range(3):
    train_step(model, batch, opt)
torch.cuda.synchronize()
t0 = time.perf_counter()
train_step(model, batch, opt)
t_enq = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    train_step(model, batch, opt)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
first, last = None, None
for ev in prof.events():
    if ev.device_type.name != "CUDA":
        continue
    name = ev.name
    if name.startswith("Memcpy") or name.startswith("Memset"):
        name = name.split(" ")[0]
    a = agg[name]
    a[0] += 1
    a[1] += ev.time_range.end - ev.time_range.start
    s, e = ev.time_range.start, ev.time_range.end
    first = s if first is None else min(first, s)
    last = e if last is None else max(last, e)
tot = sum(v[1] for v in agg.values())
lines = [f"# one Spex+ train step, n={rows}: CUPTI kernel activity (torch.profiler)", "",
         f"host enqueue {t_enq * 1e3:.1f} ms, step wall (enqueue + drain) {t_all * 1e3:.1f} ms; "
         f"kernel time sum {tot / 1e3:.1f} ms over {sum(v[0] for v in agg.values())} activities, "
         f"first-to-last span {(last - first) / 1e3:.1f} ms", "",
         "| share | total ms | avg us | count | kernel |", "|---:|---:|---:|---:|---|"]
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    lines.append(f"| {100 * t / tot:.2f} % | {t / 1e3:.2f} | {t / c:.1f} | {c} | `{name[:110]}` |")
txt = "\n".join(lines)
print(txt)
if out:
    with open(out, "w") as f:
        f.write(txt + "\n")
