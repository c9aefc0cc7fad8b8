"""Time the persistent LSTM recurrence kernels at the pBSRNN shapes (band_rnn: Q = 32 n, S = 501; band_comm: Q = 501 n,
S = 32; Hd = 256) with CUDA events, and compare h / da against the step-by-step path on a short sequence."""
import sys
import time

import torch

sys.path.insert(0, ".")
from wesep_b200 import _lib, ops   # noqa: E402
from wesep_b200.ops import _args, _stream   # noqa: E402


def run(S, Q, Hd, reps=5, seqs=0):
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    G = ops.new_act(S, 8 * Hd, Q, dev)
    G.copy_(torch.randn(S, 8 * Hd, Q, generator=g))
    H = ops.new_act(S, 2 * Hd, Q, dev)
    Cs = ops.new_act(S, 2 * Hd, Q, dev)
    dH = ops.new_act(S, 2 * Hd, Q, dev)
    dH.copy_(torch.randn(S, 2 * Hd, Q, generator=g))
    W = [(torch.randn(4 * Hd, Hd, generator=g) * Hd ** -0.5).to(dev) for _ in range(2)]
    G0 = G.clone()
    a = _args("WesepLstmRecArgs", S=S, Q=Q, Hd=Hd, ld=G.stride(1), bsG=G.stride(0), bsH=H.stride(0), G=G, H=H, C=Cs,
              Whh_f=W[0], Whh_r=W[1], dH=dH, seqs_per_cluster=seqs)
    tf, tb = [], []
    for _ in range(reps):
        G.copy_(G0)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        _lib.call("wesep_b200_lstm_rec_fwd", a, _stream())
        e[1].record()
        _lib.call("wesep_b200_lstm_rec_bwd", a, _stream())
        e[2].record()
        torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1]))
        tb.append(e[1].elapsed_time(e[2]))
    flops = 2.0 * 4 * Hd * Hd * Q * S * 2      # both directions, algorithmic
    tf_, tb_ = min(tf), min(tb)
    print(f"S={S} Q={Q} Hd={Hd} seqs/cluster={seqs}: fwd {tf_:.3f} ms ({tf_ * 1e3 / S:.2f} us/step, {flops / tf_ / 1e9:.1f} TFLOP/s alg), "
          f"bwd {tb_:.3f} ms ({tb_ * 1e3 / S:.2f} us/step)", flush=True)


def profile(S, Q, Hd, seqs=0):
    """Per-phase SM-clock stamps of cluster 0 / CTA 0 (steps 8..23)."""
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    G = ops.new_act(S, 8 * Hd, Q, dev)
    G.copy_(torch.randn(S, 8 * Hd, Q, generator=g))
    H = ops.new_act(S, 2 * Hd, Q, dev)
    Cs = ops.new_act(S, 2 * Hd, Q, dev)
    dH = ops.new_act(S, 2 * Hd, Q, dev)
    dH.copy_(torch.randn(S, 2 * Hd, Q, generator=g))
    W = [(torch.randn(4 * Hd, Hd, generator=g) * Hd ** -0.5).to(dev) for _ in range(2)]
    for name in ("wesep_b200_lstm_rec_fwd", "wesep_b200_lstm_rec_bwd"):
        prof = torch.zeros(16 * 16, dtype=torch.int64, device=dev)
        a = _args("WesepLstmRecArgs", S=S, Q=Q, Hd=Hd, ld=G.stride(1), bsG=G.stride(0), bsH=H.stride(0), G=G, H=H, C=Cs,
                  Whh_f=W[0], Whh_r=W[1], dH=dH, prof=prof, seqs_per_cluster=seqs)
        _lib.call(name, a, _stream())
        torch.cuda.synchronize()
        pr = prof.view(16, 16).cpu()
        print(name, f"S={S} Q={Q} seqs={seqs}: stamps relative to slot 0 of each step (cycles); last column = step period")
        for i in range(2, 14):
            row = pr[i]
            rel = [(int(row[k] - row[0]) if int(row[k]) else None) for k in range(11)]
            period = int(pr[i + 1][0] - row[0])
            print("  ", rel, period)


if __name__ == "__main__":
    print("max clusters fwd/bwd Hd=256:", _lib.lib().wesep_b200_lstm_rec_max_clusters(256, 0),
          _lib.lib().wesep_b200_lstm_rec_max_clusters(256, 1), " Hd=128:", _lib.lib().wesep_b200_lstm_rec_max_clusters(128, 0),
          " Hd=64:", _lib.lib().wesep_b200_lstm_rec_max_clusters(64, 0))
    profile(64, 512, 256, 64)
    profile(64, 512, 256, 128)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    t0 = time.time()
    for seqs in (64, 128):
        run(501, 32 * n, 256, seqs=seqs)
    for seqs in (64, 128):
        run(32, 501 * n, 256, seqs=seqs)
    run(501, 32 * n, 128)
    run(501, 32 * n, 192)
    print(f"wall {time.time() - t0:.1f} s")
