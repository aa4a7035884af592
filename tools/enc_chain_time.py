"""GPU box: the one-launch residual-block chain (vqb_encoder_chain) against the layer-at-a-time kernels
(vqb_token_linear): equality of the residual stream and time per 2^20 tokens."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops

dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
H = int(sys.argv[2]) if len(sys.argv) > 2 else 512
L = int(sys.argv[3]) if len(sys.argv) > 3 else 16
g = torch.Generator(device=dev).manual_seed(0)
h0 = torch.randn(T, H, device=dev, generator=g)
w = (torch.randn(L, H, H, device=dev, generator=g) * (1.0 / H) ** 0.5).to(torch.bfloat16)
b = 0.1 * torch.randn(L, H, device=dev, generator=g)
a0 = torch.nn.functional.gelu(h0).to(torch.bfloat16)


def layerwise(h):
    a = a0.clone()
    u = torch.empty_like(a)
    for i in range(L // 2):
        ops.token_linear(a, w[2 * i], b[2 * i], out=u, mode=0)
        ops.token_linear(u, w[2 * i + 1], b[2 * i + 1], h=h, out=a if 2 * i + 2 < L else None, mode=1)
    return h


ref = layerwise(h0.clone())
out = ops.encoder_chain(a0, h0.clone(), w, b)
torch.cuda.synchronize()
diff = (out - ref).abs().max().item()
print(f"T={T} H={H} L={L}: max |chain - layerwise| = {diff:.3e} (scale {ref.abs().max().item():.3f}), equal={torch.equal(out, ref)}")
wp = torch.randn(32, H, device=dev, generator=g) * (1.0 / H) ** 0.5
bp = torch.zeros(32, device=dev)
stack = torch.cat([w.reshape(-1, H), ops.projection_rows(wp)]).contiguous()
zp = ops.encoder_chain(a0, h0.clone(), stack, b, proj_bias=bp)
zr = out @ wp.t()
print(f"  fused projection: max |z - ref| = {(zp - zr).abs().max().item():.3e} (scale {zr.abs().max().item():.3f})")
for name, fn in (("chain", lambda: ops.encoder_chain(a0, hbuf, w, b)),
                 ("chain+proj", lambda: ops.encoder_chain(a0, hbuf, stack, b, proj_bias=bp)),
                 ("layerwise", lambda: layerwise(hbuf))):
    hbuf = h0.clone()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"  {name:10s} {ms:8.3f} ms  = {T / ms / 1e3:7.1f} M tokens/s, {2.0 * H * H * L * T / ms / 1e9:7.1f} TFLOP/s")
