"""GPU box: time forward (contiguous / permuted view) and backward at N = 2^24."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
B, T, D, K = 1 << 20, 16, 32, 256
phys = 0.1 * torch.randn(B, D, T, device=dev)
zperm = phys.permute(0, 2, 1)
zc = zperm.contiguous()
w = (torch.rand(K, D, device=dev) * 2 - 1) / K
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print(f"fwd contiguous auto : {timeit(lambda: ops.forward(zc, w, 0.25)):.3f} ms")
print(f"fwd permuted  auto  : {timeit(lambda: ops.forward(zperm, w, 0.25)):.3f} ms (includes the packing copy)")
print(f"fwd permuted  fma   : {timeit(lambda: ops.forward(zperm, w, 0.25, path='fma')):.3f} ms (in place)")
a = ops.forward(zc, w, 0.25); b = ops.forward(zperm, w, 0.25); c = ops.forward(zperm, w, 0.25, path="fma")
print("permuted results equal:", torch.equal(a[3], b[3]), torch.equal(a[3], c[3]), torch.equal(a[1], b[1]), torch.equal(a[1], c[1]))
idx = a[3]
g = torch.randn(B * T, D, device=dev); gl = torch.tensor(1.7, device=dev)
ms = timeit(lambda: ops.backward(g, gl, zc, idx, w, 0.25))
n = B * T
print(f"bwd contiguous      : {ms:.3f} ms -> {n * (12 * D + 8) / ms / 1e6:.0f} GB/s algorithmic ({n * (12 * D + 8) / ms / 1e6 / 6448.4 * 100:.1f}% of HBM peak)")
ms = timeit(lambda: ops.backward(g, gl, zperm, idx, w, 0.25))
print(f"bwd permuted z      : {ms:.3f} ms")
ms = timeit(lambda: ops.one_hot(idx, K), 3)
print(f"one-hot (N,K) fp32  : {ms:.3f} ms -> {n * K * 4 / ms / 1e6:.0f} GB/s")
ms = timeit(lambda: ops.backward(g, gl, zc, idx, w, 0.25, need_e=False))
print(f"bwd grad_z only     : {ms:.3f} ms -> {n * (12 * D + 8) / ms / 1e6:.0f} GB/s algorithmic (phase 1 alone)")
ms = timeit(lambda: ops.backward(g, gl, zc, idx, w, 0.25, need_z=False))
print(f"bwd grad_E only     : {ms:.3f} ms (reads z, idx; no g_zq read? no grad_z write)")
# through autograd (VQStraightThrough): forward + backward on the reference encoder's permuted layout vs contiguous rows
def fb(zin):
    zr = zin.detach().requires_grad_(True)
    wr = w.detach().requires_grad_(True)
    loss, zq, ppl, _i, _c = ops.VQStraightThrough.apply(zr, wr, 0.25, "auto")
    torch.autograd.backward([loss, zq], [gl, g.view_as(zq)])
print(f"autograd fwd+bwd contiguous : {timeit(lambda: fb(zc)):.3f} ms")
print(f"autograd fwd+bwd permuted z : {timeit(lambda: fb(zperm)):.3f} ms (one packing copy shared by forward and backward)")
