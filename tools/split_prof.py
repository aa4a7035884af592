"""GPU box: a few launches of the split (fp32-faithful) layer kernel for ncu.  python tools/split_prof.py [log2 T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
T = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
K = N = 512
ap = torch.randn(T, 2 * K, device=dev).to(torch.bfloat16); wp = ops.bf16_pair(torch.randn(N, K, device=dev) * 0.06)
b = torch.zeros(N, device=dev); h = torch.randn(T, N, device=dev); outp = torch.empty(T, 2 * N, dtype=torch.bfloat16, device=dev)
for _ in range(2):
    ops.token_linear_split(ap, wp, b, out=outp, mode=0)
    ops.token_linear_split(ap, wp, b, h=h, out=outp, mode=1)
torch.cuda.synchronize()
