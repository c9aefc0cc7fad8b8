"""Data-parallel plumbing: one process per GPU, ONE collective on the data path — the gradient
all-reduce (reference: torch DDP at wesep/bin/train.py:227-228).  The flat gradient arena is reduced
in place with NCCL over NVLink/NVSwitch (SUM); the 1/world_size average is folded into the fused
clip+Adam kernel (`grad_scale`).  Works with the gloo backend on CPU tensors for the host-logic tests."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """torchrun-style rendezvous (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        # a mismatched collective must fail fast (default watchdog: 10 min of a box's GPU time per hang)
        import datetime
        dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=int(os.environ.get("WESEP_DIST_TIMEOUT_S", "180"))))
    return rank, world, local


def shard_rows(n_total, rank, world):
    """Utterance sharding: rank r owns rows [lo, hi) (reference partitions shard files data[rank::world],
    wesep/dataset/dataset.py:98-102; the synthetic bench uses contiguous blocks)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class NcclComm:
    """An NCCL communicator owned through the C-ABI (`wesep_b200_nccl_*`, include/wesep_b200.h): rank 0 draws the unique id,
    the 128 bytes travel over the existing torch.distributed group (any backend), every rank joins.  `all_reduce(t)` enqueues
    the in-place SUM on the current CUDA stream — no torch.distributed call on the data path."""

    def __init__(self, group=None):
        import ctypes
        from wesep_b200 import _lib
        if not (dist.is_initialized() and torch.cuda.is_available()):
            raise RuntimeError("NcclComm needs an initialised process group (for the id exchange) and CUDA")
        L = _lib.lib()
        if not L.wesep_b200_nccl_available():
            raise RuntimeError("libnccl.so.2 could not be loaded")
        self._L, self._ct = L, ctypes
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        buf = ctypes.create_string_buffer(128)
        if self.rank == 0 and L.wesep_b200_nccl_unique_id(buf) != 0:
            raise RuntimeError("nccl_unique_id: " + L.wesep_b200_last_error().decode())
        box = [bytes(buf.raw)]
        dist.broadcast_object_list(box, src=0, group=group)
        self._comm = ctypes.c_void_p()
        rc = L.wesep_b200_nccl_comm_init_rank(ctypes.byref(self._comm), self.world, ctypes.c_char_p(box[0]), self.rank)
        if rc != 0:
            raise RuntimeError("nccl_comm_init_rank: " + L.wesep_b200_last_error().decode())

    def all_reduce(self, t):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise RuntimeError("NcclComm.all_reduce: contiguous fp32 CUDA tensor expected")
        ct = self._ct
        self._L.wesep_b200_nccl_allreduce_flat.argtypes = [ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p]
        rc = self._L.wesep_b200_nccl_allreduce_flat(t.data_ptr(), t.numel(), self._comm, torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            raise RuntimeError("nccl_allreduce_flat: " + self._L.wesep_b200_last_error().decode())

    def close(self):
        if self._comm:
            self._L.wesep_b200_nccl_comm_destroy(self._comm)
            self._comm = None


class GradAllReducer:
    """All-reduce(SUM) of a flat gradient buffer, in `n_buckets` contiguous pieces (reverse order, like
    DDP's reverse-registration buckets) so later rounds can overlap them with the backward pass.
    `direct=True` (or WESEP_NCCL_DIRECT=1) routes the collective through the C-ABI NCCL wrappers on the current stream
    instead of torch.distributed (CUDA tensors only)."""

    def __init__(self, flat_grad, group=None, n_buckets=1, direct=None):
        self.flat = flat_grad
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = flat_grad.numel()
        step = (n + n_buckets - 1) // n_buckets
        step = (step + 3) // 4 * 4
        self.bounds = [(lo, min(lo + step, n)) for lo in range(0, n, step)]
        if direct is None:
            direct = os.environ.get("WESEP_NCCL_DIRECT", "0") == "1"
        self.comm = NcclComm(group) if (direct and self.world > 1 and flat_grad.is_cuda) else None

    def all_reduce(self, async_op=False):
        if self.world == 1:
            return []
        if self.comm is not None:                       # stream-ordered: nothing to wait for on the host
            for lo, hi in reversed(self.bounds):
                self.comm.all_reduce(self.flat[lo:hi])
            return []
        works = []
        for lo, hi in reversed(self.bounds):
            w = dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
            if async_op:
                works.append(w)
        return works

    @property
    def grad_scale(self):
        return 1.0 / self.world


def broadcast_params(flat_param, src=0, group=None):
    """DDP constructor semantics (train.py:227): every rank starts from rank 0's parameters."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(flat_param, src=src, group=group)
