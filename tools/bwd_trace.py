"""GPU box: clock64 timeline of the TMA-ring backward kernel (CTA 0, tiles 40..103) from a -DBW_TRACE=1 build:
    python tools/ab_build.py trace:-DBW_TRACE=1 && gpurun -- python tools/bwd_trace.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("VQB_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vq-vae-transformer-arc-welding_b200", "ab_trace.so"))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
n, D, K = 1 << 24, 32, 256
zc = 0.1 * torch.randn(n, D, device=dev)
w = (torch.rand(K, D, device=dev) * 2 - 1) / K
idx = ops.forward(zc, w, 0.25)[3]
g = torch.randn(n, D, device=dev); gl = torch.tensor(1.7, device=dev)
for _ in range(3): ops.backward(g, gl, zc, idx, w, 0.25)
torch.cuda.synchronize()
import ctypes, numpy as np
lib = ctypes.CDLL(os.environ["VQB_LIB_PATH"])
buf = (ctypes.c_longlong * (6 * 64))()
print("rc", lib.vqb_debug_bw_trace(buf))
t = np.array(buf[:], dtype=np.int64).reshape(6, 64)
names = ["load_issue", "full_seen", "computed", "p2_start", "p2_done", "slot_free"]
for i in range(20, 30):
    print(i + 40, " ".join(f"{names[e]}={t[e, i] - t[0, 20]:7d}" for e in range(6)))
d = t[:, 8:60]
for a, b in ((1, 0), (2, 1), (3, 2), (4, 3), (5, 4), (5, 0)):
    v = d[a] - d[b]
    print(f"{names[b]:>10s} -> {names[a]:<10s} p10 {np.percentile(v, 10):7.0f} p50 {np.percentile(v, 50):7.0f} p90 {np.percentile(v, 90):7.0f}")
print("tile period", np.median(np.diff(d[2])))
