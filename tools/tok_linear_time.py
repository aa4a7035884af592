"""GPU box: device time of the fused per-token layer kernel (csrc/tok_linear.cu), both modes, T tokens x 512 x 512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = N = 512
a = torch.randn(T, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * 0.06).to(torch.bfloat16)
b = torch.zeros(N, device=dev); h = torch.randn(T, N, device=dev); out = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fl = 2.0 * T * K * N
for mode, fn, bytes_ in ((0, lambda: ops.token_linear(a, w, b, out=out, mode=0), T * (K * 2 + N * 2)),
                         (1, lambda: ops.token_linear(a, w, b, h=h, out=out, mode=1), T * (K * 2 + N * 10))):
    ms = timeit(fn)
    print(f"mode {mode}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {bytes_ / ms / 1e6:.0f} GB/s algorithmic")
ms = timeit(lambda: torch.matmul(a, w.t()))
print(f"torch bf16 matmul alone: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
