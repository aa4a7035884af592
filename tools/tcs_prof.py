"""GPU box: kernel-by-kernel split of one large-codebook / wide-vector call (torch profiler), and the loop ncu captures.
    python tools/tcs_prof.py K D [log2 N] [full|ids]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
K, D = int(sys.argv[1]), int(sys.argv[2])
n = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 24)
kw = {"want_zq": False, "want_loss": False} if (len(sys.argv) > 4 and sys.argv[4] == "ids") else {}
z = 0.1 * torch.randn(n, D, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
w = ((torch.rand(K, D, generator=torch.Generator().manual_seed(0)) * 2 - 1) / K).to(dev)
for _ in range(3):
    ops.forward(z, w, 0.25, **kw)
torch.cuda.synchronize()
if os.environ.get("TCS_NO_PROFILER") == "1":
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.forward(z, w, 0.25, **kw)
    torch.cuda.synchronize()
print(f"K={K} D={D} N={n} {kw}")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
