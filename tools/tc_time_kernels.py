"""GPU box: device time of the tcgen05 forward (all its kernels) via the library's event hooks."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0"); lib = vqb200._lib.load()
n = 1 << 24
z = 0.1 * torch.randn(n, 32, device=dev); w = (torch.rand(256, 32, device=dev) * 2 - 1) / 256
for _ in range(3): ops.forward(z, w, 0.25, path="tc")
torch.cuda.synchronize()
lib.vqb_profile_enable(1)
for _ in range(10): ops.forward(z, w, 0.25, path="tc")
torch.cuda.synchronize(); lib.vqb_profile_enable(0)
ms, k = ctypes.c_double(), ctypes.c_int()
lib.vqb_profile_collect(ctypes.byref(ms), ctypes.byref(k))
print(f"tc main+fixup kernels: {ms.value / k.value:.4f} ms per call  -> {n*264/ (ms.value/k.value) / 1e6:.0f} GB/s, {n*264/(ms.value/k.value)/1e6/6448.4*100:.1f}% of 6448.4")
