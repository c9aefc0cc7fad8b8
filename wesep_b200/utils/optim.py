"""Flat parameter/gradient arena + fused per-tensor clip + Adam (one kernel pair per step, no host
sync), replacing clip_gradients (wesep/utils/funcs.py:79-88: one .item() per parameter tensor)
followed by torch.optim.Adam(weight_decay) (wesep/bin/train.py:237-238)."""
import torch

from wesep_b200 import _lib
from wesep_b200.ops import _args, _stream

_CHUNK = 4096


class ParamArena:
    """Re-homes parameters (and their .grad) as views of two flat fp32 buffers. Segment offsets are
    multiples of 4 floats; padding stays zero."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise RuntimeError("ParamArena: no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("ParamArena needs CUDA parameters (no CPU fallback)")
        offs, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise RuntimeError("ParamArena: fp32 parameters on one device expected")
            offs.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.total = off
        self.offsets = offs + [off]
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, offs):
            n = p.numel()
            self.flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[o:o + n].view(p.shape)
            p.grad = self.flat_g[o:o + n].view(p.shape)
        chunk_seg, chunk_off = [], []
        for i in range(len(self.params)):
            for c in range(self.offsets[i], self.offsets[i + 1], _CHUNK):
                chunk_seg.append(i)
                chunk_off.append(c)
        self.seg_off = torch.tensor(self.offsets, dtype=torch.int64, device=dev)
        self.chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32, device=dev)
        self.chunk_off = torch.tensor(chunk_off, dtype=torch.int64, device=dev)
        self.n_chunk = len(chunk_seg)

    def zero_grad(self):
        self.flat_g.zero_()
        for p, o in zip(self.params, self.offsets):   # re-attach if someone set grads to None
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)


class FusedClipAdam:
    """Adam(betas, eps, weight_decay as coupled L2) with the reference's per-tensor clip folded in.
    `param_groups[0]["lr"]` is read every step (the scheduler writes it, executor.py:80-81)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, clip=0.0):
        self.arena = params if isinstance(params, ParamArena) else ParamArena(list(params))
        a = self.arena
        self.param_groups = [dict(params=a.params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip=clip)]
        dev = a.flat_p.device
        self.exp_avg = torch.zeros_like(a.flat_p)
        self.exp_avg_sq = torch.zeros_like(a.flat_p)
        self.sumsq = torch.empty(len(a.params), dtype=torch.float64, device=dev)
        self.norms = torch.empty(len(a.params), dtype=torch.float32, device=dev)
        self.step_count = 0
        self.grad_scale = 1.0
        self._dyn = None           # device [3] {lr, 1-b1^t, sqrt(1-b2^t)} once enable_device_scalars() was called
        self._dyn_host = None

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    # ---- CUDA-graph support: the schedule-dependent scalars move to device memory --------------------------------
    def enable_device_scalars(self):
        """After this, the kernels read lr and the Adam bias corrections from a 3-float device buffer that
        `push_scalars()` refreshes (one small pinned H2D copy), so a captured `launch()` can be replayed every step."""
        if self._dyn is None:
            dev = self.arena.flat_p.device
            self._dyn = torch.zeros(3, dtype=torch.float32, device=dev)
            # A RING of pinned staging slots, each guarded by an event recorded after its copy: the host may run several
            # steps ahead of the device (graph replays, no .item()), and rewriting a pinned source whose async copy has
            # not executed yet would hand a later step's lr / bias corrections to an earlier step.
            self._dyn_host = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(8)]
            self._dyn_evt = [None] * len(self._dyn_host)
            self._dyn_slot = 0
        return self

    def push_scalars(self):
        """Advance the step counter and upload {lr, 1-beta1^t, sqrt(1-beta2^t)} for the NEXT launch()."""
        g = self.param_groups[0]
        self.step_count += 1
        b1, b2 = float(g["betas"][0]), float(g["betas"][1])
        i = self._dyn_slot
        self._dyn_slot = (i + 1) % len(self._dyn_host)
        if self._dyn_evt[i] is not None:
            self._dyn_evt[i].synchronize()             # the copy that last read this slot has executed
        h = self._dyn_host[i]
        h[0] = float(g["lr"])
        h[1] = 1.0 - b1 ** self.step_count
        h[2] = (1.0 - b2 ** self.step_count) ** 0.5
        self._dyn.copy_(h, non_blocking=True)
        evt = self._dyn_evt[i] or torch.cuda.Event()
        evt.record(torch.cuda.current_stream(self._dyn.device))
        self._dyn_evt[i] = evt

    def launch(self):
        """The two kernels of a step, with no host-side state change (what a CUDA graph captures)."""
        if self._dyn is None:
            raise RuntimeError("FusedClipAdam.launch() needs enable_device_scalars()")
        g = self.param_groups[0]
        a = self.arena
        args = _args("WesepClipAdamArgs", total=a.total, n_seg=len(a.params), seg_off=a.seg_off, chunk_seg=a.chunk_seg,
                     chunk_off=a.chunk_off, n_chunk=a.n_chunk, param=a.flat_p, grad=a.flat_g, exp_avg=self.exp_avg,
                     exp_avg_sq=self.exp_avg_sq, sumsq=self.sumsq, norms=self.norms, grad_scale=float(self.grad_scale),
                     clip=float(g["clip"] or 0.0), lr=float(g["lr"]), beta1=float(g["betas"][0]),
                     beta2=float(g["betas"][1]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]),
                     step=max(self.step_count, 1), dyn=self._dyn)
        _lib.call("wesep_b200_clip_adam", args, _stream())
        return self.norms

    def step(self):
        if self._dyn is not None:      # device-scalar mode: same arithmetic, scalars through memory
            self.push_scalars()
            return self.launch()
        g = self.param_groups[0]
        a = self.arena
        self.step_count += 1
        args = _args("WesepClipAdamArgs", total=a.total, n_seg=len(a.params), seg_off=a.seg_off, chunk_seg=a.chunk_seg,
                     chunk_off=a.chunk_off, n_chunk=a.n_chunk, param=a.flat_p, grad=a.flat_g, exp_avg=self.exp_avg,
                     exp_avg_sq=self.exp_avg_sq, sumsq=self.sumsq, norms=self.norms, grad_scale=float(self.grad_scale),
                     clip=float(g["clip"] or 0.0), lr=float(g["lr"]), beta1=float(g["betas"][0]),
                     beta2=float(g["betas"][1]), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]),
                     step=self.step_count)
        _lib.call("wesep_b200_clip_adam", args, _stream())
        return self.norms

    # torch.optim.Adam-compatible checkpoint format (wesep/utils/checkpoint.py:94-105 stores optimizer.state_dict())
    def state_dict(self):
        a = self.arena
        state = {}
        for i, (p, o) in enumerate(zip(a.params, a.offsets)):
            n = p.numel()
            state[i] = dict(step=torch.tensor(float(self.step_count)), exp_avg=self.exp_avg[o:o + n].view(p.shape).clone(),
                            exp_avg_sq=self.exp_avg_sq[o:o + n].view(p.shape).clone())
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(a.params)))
        return dict(state=state, param_groups=[g])

    def load_state_dict(self, sd):
        a = self.arena
        for i, (p, o) in enumerate(zip(a.params, a.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
            self.step_count = int(float(st["step"]))
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v


def clip_gradients(model, clip):
    raise NotImplementedError("per-tensor clipping is fused into wesep_b200.utils.optim.FusedClipAdam(clip=...); "
                              "the reference's host loop (wesep/utils/funcs.py:79-88) is not reproduced")
