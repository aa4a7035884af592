"""GPU box: SM clock / power while a kernel runs back to back for a few seconds
(python tools/clock_probe.py [tc] [fma] [bwd] [bwdz] [copy] [tl0] [tl1]; tl0 / tl1 = the fused encoder layer, modes 0 / 1)."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml, vqb200
from vqb200 import ops
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = torch.device("cuda:0")
n = 1 << 24
z = 0.1 * torch.randn(n, 32, device=dev); w = (torch.rand(256, 32, device=dev) * 2 - 1) / 256
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.05)
idx = ops.forward(z, w, 0.25)[3]
g = torch.randn(n, 32, device=dev); gl = torch.tensor(1.7, device=dev); dst = torch.empty_like(z)
T = 1 << 20
a_t = torch.randn(T, 512, device=dev).to(torch.bfloat16); w_t = (torch.randn(512, 512, device=dev) * 0.06).to(torch.bfloat16)
b_t = torch.zeros(512, device=dev); h_t = torch.randn(T, 512, device=dev); o_t = torch.empty(T, 512, dtype=torch.bfloat16, device=dev)
def work(path):
    if path == "tl0": return ops.token_linear(a_t, w_t, b_t, out=o_t, mode=0)
    if path == "tl1": return ops.token_linear(a_t, w_t, b_t, h=h_t, out=o_t, mode=1)
    if path == "bwd": return ops.backward(g, gl, z, idx, w, 0.25)
    if path == "bwdz": return ops.backward(g, gl, z, idx, w, 0.25, need_e=False)
    if path == "copy": return dst.copy_(z)
    return ops.forward(z, w, 0.25, path=path)
for path in (sys.argv[1:] or ["tc", "fma"]):
    samples.clear(); stop = False
    t = threading.Thread(target=sampler); t.start()
    torch.cuda.synchronize(); t0 = time.time(); it = 0
    while time.time() - t0 < 3.0:
        for _ in range(50): work(path)
        torch.cuda.synchronize(); it += 50
    dt = time.time() - t0; stop = True; t.join()
    mid = samples[len(samples)//4:]
    print(f"{path}: {dt/it*1e3:.3f} ms/call; sm clock min/median/max {min(s[1] for s in mid)}/{sorted(s[1] for s in mid)[len(mid)//2]}/{max(s[1] for s in mid)} MHz; "
          f"power median {sorted(s[2] for s in mid)[len(mid)//2]:.0f} W; throttle reasons {sorted(set(hex(s[3]) for s in mid))}")
