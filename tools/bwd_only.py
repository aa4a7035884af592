"""GPU box: a few backward calls at N = 2^24 and nothing else (the command ncu is pointed at)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
n, D, K = 1 << 24, 32, 256
zc = 0.1 * torch.randn(n, D, device=dev)
w = (torch.rand(K, D, device=dev) * 2 - 1) / K
idx = ops.forward(zc, w, 0.25)[3]
g = torch.randn(n, D, device=dev); gl = torch.tensor(1.7, device=dev)
for _ in range(4): ops.backward(g, gl, zc, idx, w, 0.25)
torch.cuda.synchronize()
