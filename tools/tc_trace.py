"""Timeline of the tcgen05 kernel's pipeline (GPU box): per-tile clock64() stamps of
0 TMA issued, 1 converter saw the tile, 2 converter done, 3 MMA issue, 4 epilogue saw the accumulator,
5 filter done, 6 outputs done, 7 slot released."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
lib = vqb200._lib.load()
n = 1 << 24
z = 0.1 * torch.randn(n, 32, device=dev)
w = (torch.rand(256, 32, device=dev) * 2 - 1) / 256
for _ in range(3):
    ops.forward(z, w, 0.25, path="tc")
words = lib.vqb_debug_tc_trace_words()
buf = torch.zeros(words, dtype=torch.int64, device=dev)
lib.vqb_debug_set_tc_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.forward(z, w, 0.25, path="tc"); e1.record()
torch.cuda.synchronize()
print(f"traced call: {e0.elapsed_time(e1):.3f} ms for {n // 128} tiles on 148 CTAs = {e0.elapsed_time(e1) * 1e6 / (n / 128 / 148):.0f} ns per tile and CTA")
lib.vqb_debug_set_tc_trace(None)
t = buf.cpu().numpy().reshape(4, -1, 8).astype(np.int64)
names = ["tma", "cv_in", "cv_out", "mma", "ep_in", "flt_out", "out_done", "slot_free"]
for cta in range(2):
    a = t[cta]
    base = a[0, 0]
    print(f"--- CTA {cta}: absolute stamps (clk since first TMA), tiles 40..59")
    for i in range(40, 60):
        print(i, " ".join(f"{names[e]}={a[i, e] - base:7d}" for e in range(8)))
    d = a[20:200]
    print("median stage latencies (clk):")
    print("  tma->cv_in", np.median(d[:, 1] - d[:, 0]), " cv_in->cv_out", np.median(d[:, 2] - d[:, 1]),
          " cv_out->mma", np.median(d[:, 3] - d[:, 2]), " mma->ep_in", np.median(d[:, 4] - d[:, 3]),
          " ep_in->flt_out", np.median(d[:, 5] - d[:, 4]), " flt_out->out_done", np.median(d[:, 6] - d[:, 5]),
          " out_done->slot_free", np.median(d[:, 7] - d[:, 6]), " total", np.median(d[:, 7] - d[:, 0]))
    print("  tile period (mma issue to next mma issue):", np.median(np.diff(d[:, 3])), " tma period", np.median(np.diff(d[:, 0])))
    for name, col in (("tma->cv_in", (1, 0)), ("cv", (2, 1)), ("cv_out->mma", (3, 2)), ("mma->ep_in", (4, 3)),
                      ("filter", (5, 4)), ("R", (6, 5)), ("store", (7, 6)), ("total", (7, 0))):
        v = d[:, col[0]] - d[:, col[1]]
        print(f"  {name:12s} p10 {np.percentile(v,10):7.0f} p50 {np.percentile(v,50):7.0f} p90 {np.percentile(v,90):7.0f} p99 {np.percentile(v,99):7.0f} mean {v.mean():7.0f}")
    per = np.diff(d[:, 5])
    print(f"  filter-done period: p50 {np.percentile(per,50):.0f} mean {per.mean():.0f}; span per tile {(a[200,7]-a[20,0])/180:.0f}")
    h2 = a[22:200, 3] - a[20:198, 5]      # accumulator released by tile i (last warp) -> MMA of tile i + 2 issued
    print(f"  accumulator hand-back (flt_out(i) -> mma(i+2)): p10 {np.percentile(h2,10):.0f} p50 {np.percentile(h2,50):.0f} p90 {np.percentile(h2,90):.0f}")
    cyc = a[22:200, 3] - a[20:198, 3]
    print(f"  TMEM buffer cycle (mma(i) -> mma(i+2)): p50 {np.percentile(cyc,50):.0f} = mma->ep_in {np.median(d[:,4]-d[:,3]):.0f} + filter(last warp) {np.median(d[:,5]-d[:,4]):.0f} + hand-back {np.percentile(h2,50):.0f}")
    last = int((a[:, 7] > 0).sum()) - 1
    print(f"  whole kernel, CTA {cta}: {last + 1} tiles, first TMA -> last slot free {a[last, 7] - a[0, 0]} clocks = {(a[last, 7] - a[0, 0]) / (last + 1):.0f} per tile")
    for lo, hi in ((10, 60), (60, 120), (120, 180), (180, 250), (250, 450), (450, 650), (650, last)):
        print(f"  tiles {lo}-{hi}: mma period {np.median(np.diff(a[lo:hi, 3])):.0f}  slot_free period {np.median(np.diff(a[lo:hi, 7])):.0f}  mean {(a[hi, 7] - a[lo, 7]) / (hi - lo):.0f}")
