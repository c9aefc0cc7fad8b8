"""Time the TCN-shaped GEMMs under different tcgen05 kernel flags (CUDA events, n=32)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from wesep_b200 import _lib, ops
DEV = "cuda"
n, K = 32, 6399
x256 = ops.new_act(n, 256, K, DEV); x256.normal_()
x512 = ops.new_act(n, 512, K, DEV); x512.normal_()
W1 = torch.randn(512, 256, device=DEV) / 16
W3 = torch.randn(256, 512, device=DEV) / 22
y512 = ops.new_act(n, 512, K, DEV); y256 = ops.new_act(n, 256, K, DEV)
stats = torch.zeros(n, 2, dtype=torch.float64, device=DEV)
al = torch.tensor([0.25], device=DEV)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


ref = None
for flags, name in [(2, "1cta g2"), (4, "2cta g1"), (0, "2cta g2 (default)")]:
    _lib.lib().wesep_b200_set_tc_flags(flags)
    k2 = timed(lambda: ops.conv1x1_raw(x256, W1, False, 512, Y=y512, out_stats=stats, out_alpha=al))
    k4 = timed(lambda: ops.conv1x1_raw(x512, W3, False, 256, Y=y256, epi=2, R=x256))
    yy = ops.conv1x1_raw(x256, W1, False, 512)
    if ref is None:
        ref = yy.clone()
    err = float((yy - ref).abs().max())
    print(f"{name}: K2 {k2:.1f} us  K4 {k4:.1f} us  maxdiff vs first {err:.2e}", flush=True)
_lib.lib().wesep_b200_set_tc_flags(0)

# epilogue-pattern experiment: same GEMM as K2 (M=512, K=256) with (a) plain store, (b) residual load+store (EPI 2),
# under the debug flags of the 2-CTA kernel (bit4 no epilogue loads, bit5 no epilogue stores, bit6 no transform work)
r512 = ops.new_act(n, 512, K, DEV); r512.normal_()
for dbg in [0, 1, 2, 3, 4, 7]:
    _lib.lib().wesep_b200_set_tc_flags(dbg << 4)
    ka = timed(lambda: ops.conv1x1_raw(x256, W1, False, 512, Y=y512))
    kb = timed(lambda: ops.conv1x1_raw(x256, W1, False, 512, Y=y512, epi=2, R=r512))
    kc = timed(lambda: ops.conv1x1_raw(x512, W3, False, 256, Y=y256, epi=2, R=x256))
    print(f"dbg={dbg} (1 noload 2 nostore 4 noxf)  M=512 K=256: store-only {ka:.1f} us  load+store {kb:.1f} us ;  M=256 K=512 load+store {kc:.1f} us", flush=True)
_lib.lib().wesep_b200_set_tc_flags(0)
