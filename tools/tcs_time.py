"""GPU box: the tile-stationary tcgen05 kernel (csrc/vq_fwd_tcs.cu) on BASELINE configs[1] shapes with K > 256 or D > 64:
ms per call for the full output set and for ids only, uncertified vectors, clock64 span of the longest CTA.
    python tools/tcs_time.py [K,D ...]      (VQB_CHUNK_PASSES=1: the pass-per-chunk schedule of round 1 for comparison)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
n = int(os.environ.get("TCS_N", 1 << 24))
shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(512, 32), (1024, 32), (8192, 32), (256, 128)]
for K, D in shapes:
    g = torch.Generator(device=dev).manual_seed(1234)
    z = 0.1 * torch.randn(n, D, device=dev, generator=g)
    w = ((torch.rand(K, D, generator=torch.Generator().manual_seed(0)) * 2 - 1) / K).to(dev)
    res = {"K": K, "D": D, "N": n}
    for name, kw in (("full", {}), ("ids_only", {"want_zq": False, "want_loss": False})):
        for _ in range(2):
            out = ops.forward(z, w, 0.25, want_stats=True, **kw)
        torch.cuda.synchronize()
        reps = 3 if K * D > 65536 else 8
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = ops.forward(z, w, 0.25, want_stats=True, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        res[name + "_ms"] = round(ms, 4)
        res[name + "_alg_tflops"] = round(2.0 * K * D * n / ms / 1e9, 1)
        st = out[5].cpu().tolist()
        res["uncertified"] = st[1]; res["cta_clocks"] = st[3]
    res["hbm_frac_full"] = round(n * (8 * D + 8) / res["full_ms"] / 1e6 / 6448.4, 4)
    print(json.dumps(res), flush=True)
    del z, out
