"""GPU box: the tcgen05 path against the exact CUDA-core FMA kernel on 2^24 vectors per case (ids, histogram, z_q bit
for bit) -- run with VQB_TF32=1 / VQB_PAIR=1 to check a variant of the forward kernel at full size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
n = 1 << 24
bad = 0
for K, D, scale_z, cb in ((256, 32, 0.1, "uniform"), (256, 32, 0.1, "randn"), (256, 32, 1.0, "randn"), (200, 16, 0.1, "randn"),
                         (256, 24, 3.0, "uniform"), (129, 32, 0.02, "randn")):
    g = torch.Generator(device=dev).manual_seed(K + D)
    z = scale_z * torch.randn(n, D, device=dev, generator=g)
    w = ((torch.rand(K, D, device=dev, generator=g) * 2 - 1) / K) if cb == "uniform" else 0.1 * torch.randn(K, D, device=dev, generator=g)
    a = ops.forward(z, w, 0.25, path="tc", want_stats=True)
    b = ops.forward(z, w, 0.25, path="fma")
    mism = int((a[3] != b[3]).sum().item())
    zq_same = bool(torch.equal(a[1], b[1]))
    cnt_same = bool(torch.equal(a[4], b[4]))
    bad += mism + (not zq_same) + (not cnt_same)
    print(f"K={K} D={D} z~{scale_z}*randn codebook={cb}: id mismatches {mism} of {n}, z_q equal {zq_same}, counts equal {cnt_same}, "
          f"loss {a[0].item():.6e} vs {b[0].item():.6e}, uncertified {int(a[5][1].item())} ({100.0 * a[5][1].item() / n:.2f} %)")
    del z, a, b
print("ALL EQUAL" if bad == 0 else f"DIFFERENCES: {bad}")
