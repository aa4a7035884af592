"""GPU box: clock64 timeline of the tile-stationary kernel (csrc/vq_fwd_tcs.cu, TRACE instantiation, D <= 32):
per (tile, chunk) item  0 issuer ready, 1 MMAs issued, 2 epilogue saw the accumulator, 3 accumulator handed back,
4 item reduced, and per tile  5 code list handed to the emit warps, 6 emit warps saw it, 7 z_q walk done.
    python tools/tcs_trace.py K [ids]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
lib = vqb200._lib.load()
K = int(sys.argv[1]) if len(sys.argv) > 1 else 512
mode = sys.argv[2] if len(sys.argv) > 2 else "full"
kw = {"ids": {"want_zq": False, "want_loss": False}, "loss": {"want_zq": False}, "zq": {"want_loss": False}}.get(mode, {})
n, D = 1 << 24, 32
nc = (K + 255) // 256
z = 0.1 * torch.randn(n, D, device=dev)
w = ((torch.rand(K, D) * 2 - 1) / K).to(dev)
for _ in range(3):
    ops.forward(z, w, 0.25, path="tc", **kw)
buf = torch.zeros(lib.vqb_debug_tc_trace_words(), dtype=torch.int64, device=dev)
lib.vqb_debug_set_tc_trace(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.forward(z, w, 0.25, path="tc", **kw); e1.record()
torch.cuda.synchronize()
lib.vqb_debug_set_tc_trace(None)
print(f"K={K} {kw}: traced call {e0.elapsed_time(e1):.3f} ms, {nc} chunks per tile")
t = buf.cpu().numpy().reshape(4, -1, 8).astype(np.int64)
def pct(v): return f"p10 {np.percentile(v,10):6.0f} p50 {np.percentile(v,50):6.0f} p90 {np.percentile(v,90):6.0f}"
for cta in range(2):
    a = t[cta]
    # items in issue order: pairs of tiles, chunk-major inside a pair
    order = []
    for g0 in range(0, 1024 // nc - 2, 2):
        for c in range(nc):
            for j in range(2):
                order.append((g0 + j) * nc + c)
    order = np.array(order[8:400])
    b = a[order]
    print(f"--- CTA {cta}")
    print("  issuer ready -> MMAs issued      ", pct(b[:, 1] - b[:, 0]))
    print("  MMAs issued -> epilogue saw them  ", pct(b[:, 2] - b[:, 1]))
    print("  accumulator drain (2 -> 3)        ", pct(b[:, 3] - b[:, 2]))
    print("  rest of the item (3 -> 4)         ", pct(b[:, 4] - b[:, 3]))
    print("  item period (issue to issue)      ", pct(np.diff(b[:, 1])), " mean", np.diff(b[:, 1]).mean())
    print("  hand-back: 3(n) -> issued(n+2)    ", pct(b[2:, 1] - b[:-2, 3]))
    print("  hand-back: 3(n) -> ready(n+2)     ", pct(b[2:, 0] - b[:-2, 3]))
    print("  buffer cycle issued(n)->issued(n+2)", pct(b[2:, 1] - b[:-2, 1]))
    if not kw:
        last = np.array([i * nc + nc - 1 for i in range(8, 300)])
        d = a[last]
        print("  tile: last item reduced -> codes handed", pct(d[:, 5] - d[:, 4]))
        print("  tile: codes handed -> emit saw them    ", pct(d[:, 6] - d[:, 5]))
        print("  tile: z_q walk                          ", pct(d[:, 7] - d[:, 6]))
        print("  tile: walk period                       ", pct(np.diff(d[:, 7])))
    print("  first items (absolute):")
    base = a[0, 0]
    for k in order[:12]:
        print("   item", k, " ".join(f"{a[k, e] - base:8d}" for e in range(8)))
