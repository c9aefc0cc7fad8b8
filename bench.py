#!/usr/bin/env python
"""bench.py — Spex+ train-step throughput (utterances/sec) on synthetic 4 s @ 16 kHz two-speaker
mixtures, batch = 32 model rows per GPU (BASELINE.json configs[1]; configs[3] for N > 1).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows 32]

A "step" = forward + 0.8/0.1/0.1 SI-SDR + 0.5 CE loss + backward + gradient all-reduce (N > 1) +
per-tensor clip + Adam, the loop body of the reference Executor.train (wesep/utils/executor.py:70-134).
Prints ONE JSON line on rank 0.  `value`: inputs already resident in HBM; `e2e`: same step through
the public API from pinned host buffers (H2D every step, D2H of the loss every step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SPEX_ARGS = dict(B=256, H=512, L=20, N=256, P=3, R=4, X=8, spk_emb_dim=256, activate="relu", causal=False, norm="gLN",
                 skip_con=False, spk_fuse_type="concatConv", use_spk_transform=False, multi_fuse=True,
                 encoder_type="Multi", decoder_type="Multi", joint_training=True, multi_task=True, spksInTrain=251)
T_SAMPLES = 64000
METRIC = "utterances/sec Spex+ train step (4s@16kHz)"
# dram__bytes_read.sum + dram__bytes_write.sum of the K2 GEMM at n=32 (ncu --set full, profiles/r01_ncu_full_tcn_block_n32.md)
K2_DRAM_BYTES_PER_LAUNCH = 579.5e6


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons),
                    samples=len(sm))


# ----------------------------------------------------------------------------- reference / CPU arm
def cpu_threads():
    """Intra-op threads for the CPU arm.  torch's CPU kernels on this model stop scaling (and then collapse:
    302 s/step with 128 threads vs 14 s/step with 8 on the same code) well below the box's core count, so the
    reference arm uses min(cores, 32) and says so in `cores`."""
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_train_rows_per_s(rows, steps, warmup, threads):
    """The reference's algorithm on the host cores: oracle port (plain torch CPU fp32) of the same train
    step incl. reference-style per-tensor clip + Adam.  Bounded sample: `rows` model rows per step."""
    from oracle import losses as olosses
    from oracle import optim as ooptim
    from oracle import spexplus as ospex
    from wesep_b200 import synth
    torch.set_num_threads(threads)
    cfg = dict(ospex.DEFAULT_CFG)
    sd = ospex.make_state_dict(cfg)
    synth.fill_state_dict_(sd, seed=0)
    names = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    P = [sd[k].requires_grad_(True) for k in names]
    m = [torch.zeros_like(p) for p in P]
    v = [torch.zeros_like(p) for p in P]
    b = synth.make_batch(rows, T=T_SAMPLES, Te=T_SAMPLES, seed=1234)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        bufs = {}
        out = ospex.convtasnet_forward(sd, cfg, b["wav_mix"], b["spk_embeds"], training=True, buffers_out=bufs)
        loss, _ = olosses.train_loss(out, b["wav_targets"], b["spk_label"])
        grads = [g.clone() for g in torch.autograd.grad(loss, P)]
        ooptim.clip_gradients(grads, 5.0)
        with torch.no_grad():
            ooptim.adam_step(P, grads, m, v, it + 1, 1e-3)
            sd.update(bufs)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return rows / med, med, float(loss)


def run_reference(args, rank):
    if rank != 0:
        return
    threads = cpu_threads()
    rows = args.ref_rows
    val, med, loss = cpu_train_rows_per_s(rows, min(args.steps, 5), min(args.warmup, 1), threads)
    sample = f"{rows} rows x {T_SAMPLES} samples per step, {args.steps} timed steps (median), oracle port on host CPU"
    line = dict(metric=METRIC, value=val, unit="utterances/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=med * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload="Spex+ train step, 4s@16kHz, CPU sample of %d rows" % rows, rows_per_step=rows),
                cpu_baseline=dict(value=val, unit="utterances/s", cores=threads, kind="port", sample=sample),
                e2e=dict(value=val, unit="utterances/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0, loss=loss)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- our arm
def kernel_rooflines(n, dev, pk):
    """Time the TCN-block kernels alone at the workload shape (CUDA events on the launching stream;
    operands >> L2 so every launch streams from HBM).  Algorithmic bytes: DESIGN.md §kernels."""
    from wesep_b200 import ops, synth
    from wesep_b200.modules.tasnet.convs import Conv1DBlock
    B, H, K = 256, 512, 6399
    blk = Conv1DBlock(B, H, 3, 8, "gLN", False, False)
    synth.fill_state_dict_(blk.state_dict(), seed=1)
    blk = blk.to(dev)
    x = ops.new_act(n, B, K, dev)
    x.normal_()
    x.requires_grad_(True)
    gy = ops.new_act(n, B, K, dev)
    gy.normal_()

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps * 1e-3

    out = {}
    # K2-shaped GEMM (256 -> 512, bias, gLN statistics in the epilogue)
    W = blk.conv1x1.weight.detach().reshape(H, B)
    stats = torch.zeros((n, 2), dtype=torch.float64, device=dev)
    u = ops.new_act(n, H, K, dev)
    xd = x.detach()
    t = timed(lambda: ops.conv1x1_raw(xd, W, False, H, bias=blk.conv1x1.bias.detach(), Y=u, out_stats=stats,
                                      out_alpha=blk.PReLU_1.weight.detach()), 10)
    flops = 2.0 * H * B * K * n
    bytes_ = 4.0 * (B + H) * K * n
    out["gemm_wx_k2"] = dict(seconds=t, alg_tflops=flops / t / 1e12, exec_tflops=3 * flops / t / 1e12,
                             alg_gbs=bytes_ / t / 1e9)
    # whole block forward / backward
    tf = timed(lambda: blk(xd), 5)
    y = blk(x)
    tb = timed(lambda: torch.autograd.grad(y, [x] + list(blk.parameters()), gy, retain_graph=True), 5)
    bf = (2 * B + 4 * H) * K * 4.0 * n
    bb = (3 * B + 8 * H) * K * 4.0 * n
    out["tcn_block_fwd"] = dict(seconds=tf, alg_gbs=bf / tf / 1e9, frac_hbm=bf / tf / 1e9 / pk["hbm"])
    out["tcn_block_bwd"] = dict(seconds=tb, alg_gbs=bb / tb / 1e9, frac_hbm=bb / tb / 1e9 / pk["hbm"])
    return out


def run_ours(args, rank, world, local):
    from wesep_b200 import _lib, synth
    from wesep_b200.distributed import GradAllReducer, broadcast_params
    from wesep_b200.models import get_model
    from wesep_b200.utils.executor import train_step
    from wesep_b200.utils.optim import FusedClipAdam
    from wesep_b200.utils.schedulers import ExponentialDecrease
    import torch.distributed as dist

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    pk = peaks()
    n = args.rows
    torch.manual_seed(42 + rank)                                # train.py:88 per-rank seed
    model = get_model("ConvTasNet")(**SPEX_ARGS).to(dev).train()
    opt = FusedClipAdam(model.parameters(), lr=1e-3, weight_decay=1e-4, clip=5.0)
    broadcast_params(opt.arena.flat_p)
    reducer = GradAllReducer(opt.arena.flat_g, n_buckets=3) if world > 1 else None
    sched = ExponentialDecrease(opt, num_epochs=150, epoch_iter=1000, initial_lr=1e-3, final_lr=2.5e-5, warm_up_epoch=0)
    host = synth.make_batch(n, T=T_SAMPLES, Te=T_SAMPLES, seed=1234 + rank, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    graphed = None
    if args.cuda_graph:
        if world > 1:
            raise SystemExit("--cuda-graph is single-process (the gradient all-reduce is not captured)")
        from wesep_b200.utils.executor import GraphedTrainStep
        graphed = GraphedTrainStep(model, opt, resident, warmup=3)   # 3 eager steps, then ONE capture of the whole step

    def timed_region(batch, read_loss):
        it = [0]

        def one():
            sched.step(it[0])
            it[0] += 1
            loss = graphed(batch) if graphed is not None else train_step(model, batch, opt, reducer)
            if read_loss:
                return loss.item()                              # D2H of the step's result
            return loss
        for _ in range(args.warmup):
            one()
        barrier()
        l0 = _lib.launch_count()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        last = None
        for _ in range(args.steps):
            last = one()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        launches = _lib.launch_count() - l0
        if graphed is not None:                                 # replays do not pass through the host-side counter
            launches = graphed.launches_per_step * args.steps
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms, launches, float(last)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_res, launches, loss_res = timed_region(resident, read_loss=False)
    ms_e2e, _, loss_e2e = timed_region(host, read_loss=True)
    clocks = sampler.summary() if sampler else None
    value = n * world * args.steps / (ms_res * 1e-3)
    e2e = n * world * args.steps / (ms_e2e * 1e-3)
    if rank != 0:
        return
    roof = kernel_rooflines(n, dev, pk)
    dom = roof["gemm_wx_k2"]
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads()
        v, med, _ = cpu_train_rows_per_s(args.ref_rows, 2, 1, threads)
        cpu = dict(value=v, unit="utterances/s", cores=threads, kind="port",
                   sample=f"{args.ref_rows} rows x {T_SAMPLES} samples, 1 warm-up + 2 timed steps (median {med:.2f} s), "
                          "oracle port (plain torch fp32) on host CPU")
    line = dict(
        metric=METRIC, value=value, unit="utterances/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_res / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
        data="synthetic", impl="ours",
        config=dict(workload="Spex+ (ConvTasNet, examples/librimix/tse/v2/confs/spexplus.yaml) full train step, "
                             "4s@16kHz, %d model rows per GPU" % n,
                    rows_per_gpu=n, global_rows=n * world, samples=T_SAMPLES, parallelism="dp%d" % world,
                    gemm_mode="3xTF32 split (fp32-grade); tcgen05.mma kind::tf32 cta_group::2 + TMA + TMEM GEMMs, mma.sync for odd shapes", l2="inputs and activations >> L2 (126 MB)",
                    loss="0.8/0.1/0.1 SI-SDR + 0.5 CE", optimizer="per-tensor clip 5.0 + Adam(wd 1e-4), exp-decay lr",
                    launch="one CUDA-graph replay per step" if graphed is not None else "eager (one launch per kernel)"),
        e2e=dict(value=e2e, unit="utterances/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                 ms_per_step=ms_e2e / args.steps),
        gpu_launches=launches, clocks=clocks, loss=loss_res, loss_e2e=loss_e2e,
        roofline=dict(kernel="gemm_wx_tc2_kernel<0,0,6> (tcgen05 cta_group::2; K2 shape 256->512, n=%d, K=6399)" % n,
                      bound="tensor", achieved=dom["exec_tflops"], peak=pk["tf_burst"], unit="TFLOP/s",
                      frac=dom["exec_tflops"] / pk["tf_burst"], frac_of_tf32_peak=dom["exec_tflops"] / (0.5 * pk["tf_burst"]),
                      traffic=K2_DRAM_BYTES_PER_LAUNCH if n == 32 else None,
                      note="executed = 3x algorithmic flops (3xTF32 split, fp32-grade); peak = measured bf16 burst (%s); "
                           "kind::tf32 runs at half the bf16 rate, so frac_of_tf32_peak is the pipe utilisation; traffic = "
                           "dram read+write of this kernel per launch from profiles/r01_ncu_full_tcn_block_n32.md" % pk["src"],
                      algorithmic_tflops=dom["alg_tflops"], alg_gbs=dom["alg_gbs"]),
        roofline_tcn_block=dict(bound="hbm", peak=pk["hbm"], unit="GB/s", fwd=roof["tcn_block_fwd"],
                                bwd=roof["tcn_block_bwd"],
                                note="algorithmic bytes per row per block: fwd (2B+4H)*K*4 = 65.5 MB, bwd (3B+8H)*K*4 = "
                                     "124.5 MB (SURVEY 8d)"),
        cpu_baseline=cpu)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=32, help="model rows (utterances) per GPU per step")
    ap.add_argument("--ref-rows", type=int, default=1, help="rows per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="capture the whole train step in a CUDA graph and time replays (single GPU)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a GPU (no CPU fallback)")
    from wesep_b200.distributed import init_from_env
    rank, world, local = init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    run_ours(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
