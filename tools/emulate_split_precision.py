"""CPU emulation of the split-product GEMM variants (operand rounding only, exact accumulation), relative L2 vs fp64:
1xTF32, today's 3xTF32 (lo.hi + hi.lo + hi.hi, all kind::tf32), and the round-2 candidate hi.hi in tf32 + the two cross
terms with bf16 operands (kind::f16 at twice the MMA rate).  Measured here: 7.7e-4, 3.5e-7, 1.3e-6 (fp32 GEMM: 2.9e-7);
on the GPU the tensor core's accumulation adds ~1e-6 to all of them (3xTF32 measured 1.5e-6)."""
import torch

torch.manual_seed(0)


def tf32_trunc(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def rel(a, b):
    return float((a - b).norm() / b.norm())


bf = lambda t: t.bfloat16().float()  # noqa: E731
for K in (256, 512):
    M, N = 512, 2048
    W = torch.randn(M, K) / K ** 0.5
    X = torch.randn(K, N)
    exact = W.double() @ X.double()
    Wh, Xh = tf32_trunc(W), tf32_trunc(X)
    Wl, Xl = W - Wh, X - Xh
    one = Wh.double() @ Xh.double()
    t3 = one + tf32_trunc(Wl).double() @ Xh.double() + Wh.double() @ tf32_trunc(Xl).double()
    mix = one + bf(Wl).double() @ bf(Xh).double() + bf(Wh).double() @ bf(Xl).double()
    print(f"K={K}: 1xTF32 {rel(one, exact):.2e}  3xTF32 {rel(t3, exact):.2e}  tf32 + 2 x bf16 cross terms {rel(mix, exact):.2e}  "
          f"fp32 {rel((W @ X).double(), exact):.2e}")
