"""GPU box: per-launch device time of back-to-back forward calls after an idle gap -- does the 20-launch burst bench.py
times run at one speed, or does the board's power limiter pull the clock down inside the burst?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
n = 1 << 24
z = 0.1 * torch.randn(n, 32, device=dev); w = (torch.rand(256, 32, device=dev) * 2 - 1) / 256
for _ in range(5): ops.forward(z, w, 0.25, path="tc")
torch.cuda.synchronize()
for gap in (0.0, 0.5, 2.0):
    time.sleep(gap)
    m = 60
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(m + 1)]
    ev[0].record()
    for i in range(m):
        ops.forward(z, w, 0.25, path="tc")
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(m)]
    print(f"idle {gap:.1f} s before the burst: launch 1-5 {sum(t[0:5])/5:.3f} ms, 6-10 {sum(t[5:10])/5:.3f}, 11-20 {sum(t[10:20])/10:.3f}, "
          f"21-40 {sum(t[20:40])/20:.3f}, 41-60 {sum(t[40:60])/20:.3f}; first 20 mean {sum(t[:20])/20:.3f}")
