"""GPU box, for ncu: a few launches of vqb_encoder_chain on T tokens (default 2^18), nothing else."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
H, L = 512, 16
g = torch.Generator(device=dev).manual_seed(0)
h0 = torch.randn(T, H, device=dev, generator=g)
w = (torch.randn(L, H, H, device=dev, generator=g) * (1.0 / H) ** 0.5).to(torch.bfloat16)
b = 0.1 * torch.randn(L, H, device=dev, generator=g)
a0 = torch.nn.functional.gelu(h0).to(torch.bfloat16)
for _ in range(4):
    ops.encoder_chain(a0, h0.clone(), w, b)
torch.cuda.synchronize()
print("done")
