"""GPU box: time + check every experiment variant  (python tools/ab_run.py name1 name2 ...; 'base' = product lib)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vq-vae-transformer-arc-welding_b200")
CHILD = r'''
import os, sys, ctypes
sys.path.insert(0, %r)
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0"); lib = vqb200._lib.load()
n = 1 << 24
torch.manual_seed(1)
z = 0.1 * torch.randn(n, 32, device=dev); w = (torch.rand(256, 32, device=dev) * 2 - 1) / 256
ref = ops.forward(z[:1 << 20], w, 0.25, path="fma"); out = ops.forward(z[:1 << 20], w, 0.25, path="tc", want_stats=True)
ok = bool(torch.equal(ref[3], out[3]) and torch.equal(ref[1], out[1]) and torch.equal(ref[4], out[4]))
for _ in range(3): ops.forward(z, w, 0.25, path="tc")
torch.cuda.synchronize()
best = 1e9; tot = 0
for rep in range(3):
    lib.vqb_profile_enable(1)
    for _ in range(20): ops.forward(z, w, 0.25, path="tc")
    torch.cuda.synchronize(); lib.vqb_profile_enable(0)
    ms, k = ctypes.c_double(), ctypes.c_int()
    lib.vqb_profile_collect(ctypes.byref(ms), ctypes.byref(k))
    t = ms.value / k.value; best = min(best, t); tot += t
import time, threading, pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples = []; stop = False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.05)
sus = ""
if os.environ.get("VQB_SUSTAINED"):
    th = threading.Thread(target=sampler); th.start()
    torch.cuda.synchronize(); t0 = time.time(); it = 0
    lib.vqb_profile_enable(1)
    while time.time() - t0 < 2.5:
        for _ in range(50): ops.forward(z, w, 0.25, path="tc")
        torch.cuda.synchronize(); it += 50
    lib.vqb_profile_enable(0)
    ms, k = ctypes.c_double(), ctypes.c_int()
    lib.vqb_profile_collect(ctypes.byref(ms), ctypes.byref(k))
    stop = True; th.join(); mid = samples[len(samples) // 3:]
    sus = f" | sustained kernel {ms.value / k.value:.4f} ms, clk {sorted(s[0] for s in mid)[len(mid)//2]} MHz, {sorted(s[1] for s in mid)[len(mid)//2]:.0f} W"
st = ops.forward(z, w, 0.25, path="tc", want_stats=True)[5]; clk = int(st[3]); tiles = n // 128 / 148
print(f"{os.environ.get('VQB_NAME'):24s} exact={ok} slow_rows={int(out[5][1])} clk/tile {clk / tiles:.0f} ({clk / (best * 1e-3) / 1e9:.2f} GHz at best)  mean {tot/3:.4f} ms best {best:.4f} ms -> {n*264/(tot/3)/1e6/6448.4*100:.1f}%% of HBM roofline" + sus)
''' % ROOT
for name in sys.argv[1:]:
    env = dict(os.environ, VQB_NAME=name)
    if name != "base":
        env["VQB_LIB_PATH"] = os.path.join(PKG, f"ab_{name}.so")
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    print(r.stdout.strip() or ("FAILED " + name + ": " + r.stderr[-800:]), flush=True)
