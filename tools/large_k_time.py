"""GPU box: per-call time of the chunked tcgen05 path (K > 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
n = 1 << 24; D = int(sys.argv[2]) if len(sys.argv) > 2 else 32; K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
z = 0.1 * torch.randn(n, D, device=dev); w = ((torch.rand(K, D) * 2 - 1) / K).to(dev)
for _ in range(3): ops.forward(z, w, 0.25)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.forward(z, w, 0.25)
e1.record(); torch.cuda.synchronize()
print(f"K={K} D={D}: {e0.elapsed_time(e1)/5:.3f} ms per call")
