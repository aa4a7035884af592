"""Debug helper (GPU box): compare the tcgen05 path against the FMA path on the same input."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vqb200
from vqb200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(n, K, D=32, zscale=0.1, cb="default", reps=1):
    z = zscale * torch.randn(n, D, device=dev)
    if cb == "default":
        w = (torch.rand(K, D, device=dev) * 2 - 1) / K
    else:
        w = 0.1 * torch.randn(K, D, device=dev)
    ref = ops.forward(z, w, 0.25, path="fma", want_stats=True)
    torch.cuda.synchronize()
    out = ops.forward(z, w, 0.25, path="tc", want_stats=True)
    torch.cuda.synchronize()
    mism = int((ref[3] != out[3]).sum())
    zq_ok = bool(torch.equal(ref[1], out[1]))
    cnt_ok = bool(torch.equal(ref[4], out[4]))
    print(f"n={n} K={K} z={zscale} cb={cb}: idx mismatches={mism} zq_equal={zq_ok} counts_equal={cnt_ok} "
          f"loss {ref[0].item():.8e} vs {out[0].item():.8e} ppl {ref[2].item():.5f} vs {out[2].item():.5f} "
          f"slow_rows={int(out[5][1])} nonfinite_rows={int(out[5][2])}", flush=True)
    if mism:
        bad = (ref[3] != out[3]).view(-1).nonzero().view(-1)[:8]
        print("  first bad rows", bad.tolist(), "ref", ref[3].view(-1)[bad].tolist(), "tc", out[3].view(-1)[bad].tolist())
    if reps > 1:
        for path in ("fma", "tc"):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(reps):
                ops.forward(z, w, 0.25, path=path)
            torch.cuda.synchronize()
            print(f"  {path}: {(time.perf_counter()-t0)/reps*1e3:.3f} ms/call")

if __name__ == "__main__" and "--timing" not in sys.argv:
    run(128, 256)
    run(1000, 256)
    run(4096, 256, cb="randn")
    run(4096, 100)
    run(65536, 256, zscale=1.0)
    run(1 << 20, 256, reps=5)
    run(1 << 24, 256, reps=5)


def timing(n=1 << 24, K=256, reps=10):
    z = 0.1 * torch.randn(n, 32, device=dev)
    w = (torch.rand(K, 32, device=dev) * 2 - 1) / K
    for want_zq in (True, False):
        for _ in range(3):
            ops.forward(z, w, 0.25, path="tc", want_zq=want_zq, want_loss=want_zq)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.forward(z, w, 0.25, path="tc", want_zq=want_zq, want_loss=want_zq)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_ = n * (264 if want_zq else 136)
        print(f"timing n={n} zq={want_zq}: {ms:.3f} ms/call  {bytes_/ms/1e6:.0f} GB/s algorithmic", flush=True)

if __name__ == "__main__" and "--timing" in sys.argv:
    timing()
