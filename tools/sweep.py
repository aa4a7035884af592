"""GPU box: BASELINE configs[1] sweep -- standalone VectorQuantizer.forward on 2^24 synthetic vectors for
codebook sizes K and latent dims D, default-init codebook, path 'auto' (tcgen05 where the shape allows,
else the exact CUDA-core FMA kernel).  One JSON line per (K, D): device ms per call (CUDA events), vectors/s,
the algorithmic HBM GB/s (N*(8D+8)/t) and FLOP/s (2KD*N/t) against the measured peaks, and the bound that
applies (arithmetic intensity 2KD/(8D+8) against the ridge of the pipe that runs it)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
lib = vqb200._lib.load()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
HBM, TENSOR, FMA32 = peaks["hbm_gbs"], peaks["bf16_tflops_sustained"], 74.0     # fp32 FMA: 148 SMs x 128 lanes x 2 x 1.965 GHz
n = int(os.environ.get("SWEEP_N", 1 << 24))
Ks = [int(k) for k in os.environ.get("SWEEP_K", "64,128,256,512,1024,2048,4096,8192").split(",")]
Ds = [int(d) for d in os.environ.get("SWEEP_D", "8,16,32,64,128").split(",")]
for D in Ds:
    g = torch.Generator(device=dev).manual_seed(1234)
    z = 0.1 * torch.randn(n, D, device=dev, generator=g)
    for K in Ks:
        w = ((torch.rand(K, D, generator=torch.Generator().manual_seed(0)) * 2 - 1) / K).to(dev)
        path = "tc" if lib.vqb_select_path(0, n, K, D, D, 1) == 2 else "fma"
        flop = 2.0 * K * D * n
        reps = 3 if flop > 2e12 else 10
        for _ in range(2):
            out = ops.forward(z, w, 0.25)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = ops.forward(z, w, 0.25)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = n * (8 * D + 8) / ms / 1e6
        tfl = flop / ms / 1e9
        ai = 2.0 * K * D / (8 * D + 8)
        pipe_peak = TENSOR if path == "tc" else FMA32
        bound = "hbm" if ai < pipe_peak * 1e3 / HBM else ("tensor" if path == "tc" else "fp32-fma")
        print(json.dumps({"K": K, "D": D, "N": n, "path": path, "ms": round(ms, 4), "vectors_per_s": n / ms * 1e3,
                          "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / HBM, 4), "tflops": round(tfl, 2),
                          "fma_frac": round(tfl / FMA32, 4), "arith_intensity": round(ai, 1), "bound": bound,
                          "counts_total": int(out[4].sum().item())}), flush=True)
        del out
    del z
