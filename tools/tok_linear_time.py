"""GPU box: device time of the fused per-token layer kernel (csrc/tok_linear.cu), both modes, T tokens x 512 x 512."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from vqb200 import ops
dev = torch.device("cuda:0")
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = N = 512
a = torch.randn(T, K, device=dev).to(torch.bfloat16); w = (torch.randn(N, K, device=dev) * 0.06).to(torch.bfloat16)
b = torch.zeros(N, device=dev); h = torch.randn(T, N, device=dev); out = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
fl = 2.0 * T * K * N
for mode, fn, bytes_ in ((0, lambda: ops.token_linear(a, w, b, out=out, mode=0), T * (K * 2 + N * 2)),
                         (1, lambda: ops.token_linear(a, w, b, h=h, out=out, mode=1), T * (K * 2 + N * 10))):
    ms = timeit(fn)
    print(f"mode {mode}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s  {bytes_ / ms / 1e6:.0f} GB/s algorithmic")
ms = timeit(lambda: torch.matmul(a, w.t()))
print(f"torch bf16 matmul alone: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
# the fp32-faithful (split) form: bf16 hi + lo pairs, three products per layer
ap = torch.randn(T, 2 * K, device=dev).to(torch.bfloat16); wp = ops.bf16_pair(torch.randn(N, K, device=dev) * 0.06)
outp = torch.empty(T, 2 * N, dtype=torch.bfloat16, device=dev)
for mode, fn, bytes_ in ((0, lambda: ops.token_linear_split(ap, wp, b, out=outp, mode=0), T * (K * 4 + N * 4)),
                         (1, lambda: ops.token_linear_split(ap, wp, b, h=h, out=outp, mode=1), T * (K * 4 + N * 12))):
    ms = timeit(fn)
    print(f"split mode {mode}: {ms:.3f} ms  {3 * fl / ms / 1e9:.0f} issued TFLOP/s ({fl / ms / 1e9:.0f} algorithmic)  {bytes_ / ms / 1e6:.0f} GB/s algorithmic")
hf = torch.randn(T, N, device=dev)
ms = timeit(lambda: ops.token_pair(hf, gelu=True, out=outp))
print(f"token_pair: {ms:.3f} ms  {T * N * 8 / ms / 1e6:.0f} GB/s")
torch.backends.cuda.matmul.allow_tf32 = False
af = torch.randn(T, K, device=dev); wf = torch.randn(N, K, device=dev)
ms = timeit(lambda: torch.matmul(af, wf.t()), reps=3)
print(f"torch fp32 matmul alone: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
# whole encoder, 65536 cycles: fp32 PyTorch layers / fused_fp32 / fused_bf16
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, patch_size=25, batch_norm=False).to(dev).eval()
x = torch.randn(65536, 200, 2, device=dev)
res = {}
with torch.no_grad():
    for m in ("torch", "fused_fp32", "fused_bf16"):
        model.encoder_mode = m
        ms = timeit(lambda: model.encode_ids(x), reps=3)
        res[m] = model.encode_ids(x)
        z = model.encode(x[:4096]).double()
        if m == "torch":
            z0 = z
        print(f"encode_ids {m}: {ms:.2f} ms per 65536 cycles = {65536 * 16 / ms / 1e3:.1f} M patches/s; ids equal to torch: "
              f"{(res[m] == res['torch']).double().mean().item():.7f}; max |z - z_torch| / range = {(z - z0).abs().max().item() / z0.abs().max().item():.2e}")
from torch.profiler import profile, ProfilerActivity
model.encoder_mode = "fused_fp32"
with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        model.encode_ids(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
