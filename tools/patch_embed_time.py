"""GPU box: the fused patch-embedding kernel against the stock sequence it replaces (T = 2^20 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
B, L, C, P, H = 1 << 16, 200, 2, 25, 512
x = torch.randn(B, L, C, device=dev)
w = torch.randn(H, 1, P, device=dev) * 0.2; b = torch.randn(H, device=dev) * 0.1
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def stock():
    patches = x.permute(0, 2, 1).reshape(-1, P)
    h = torch.matmul(patches, w[:, 0, :].t())
    return h, ops.token_bias_gelu(h, b)
n = B * 16
ms = timeit(lambda: ops.patch_embed(x, w, b, P))
print(f"vqb_patch_embed            : {ms:.3f} ms  ({n * (4 * P + 6 * H) / ms / 1e6:.0f} GB/s algorithmic, {2 * n * P * H / ms / 1e9:.1f} TFLOP/s fp32)")
print(f"permute + SGEMM + bias/GELU: {timeit(stock):.3f} ms")
h, a = ops.patch_embed(x, w, b, P); h2, a2 = stock()
print("max |h - h_stock|", float((h - h2).abs().max()), " bf16 act equal fraction", float((a == a2).float().mean()))
