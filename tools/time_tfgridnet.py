"""Time one TF-GridNet train step (tfgridnet.yaml network, BASELINE config 5, speaker embedding given, 4 s @ 16 kHz) on one GPU and list the kernel
shares of one step (CUPTI through torch.profiler).  Usage: python tools/time_tfgridnet.py [rows] [samples]
Prints one JSON object: utterances/s, ms per step, peak memory, the top kernels."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from wesep_b200 import ops, synth  # noqa: E402
from wesep_b200.models import get_model  # noqa: E402
from wesep_b200.utils.optim import FusedClipAdam  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 64000
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = get_model("TFGridNet")(n_srcs=1, n_fft=128, stride=64, window="hann", n_imics=1, n_layers=6, lstm_hidden_units=192,
                               attn_n_head=4, attn_approx_qk_dim=512, emb_dim=128, emb_ks=1, emb_hs=1, activation="prelu", eps=1e-5,
                               use_spk_transform=False, spk_fuse_type="multiply", joint_training=False).to(dev).train()
    opt = FusedClipAdam(m.parameters(), lr=1e-3, weight_decay=1e-4, clip=5.0)
    b = synth.make_batch(n, T=L, Te=8, seed=1, device=dev)
    emb = torch.from_numpy(np.random.default_rng(2).standard_normal((n, 256)).astype(np.float32)).to(dev)

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
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    t0.record()
    for _ in range(K):
        loss = step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / K
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type.name == "CUDA":
            c, t = agg.get(e.name, (0, 0.0))
            agg[e.name] = (c + 1, t + e.device_time)
    tot = sum(v[1] for v in agg.values())
    top = [dict(kernel=k[:60], count=v[0], ms=v[1] / 1e3, share=v[1] / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]]
    print(json.dumps(dict(rows=n, samples=L, ms_per_step=ms, utt_per_s=n / ms * 1e3, loss=float(loss),
                          peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, kernel_ms=tot / 1e3, top=top)))


if __name__ == "__main__":
    main()
