"""GPU box: the decoder half on the fused layer kernels (decoder_mode = "fused_bf16") at bulk size -- ms per call against the
stock PyTorch decoder (fp32 and bf16 autocast), kernel-by-kernel split, deviation on identical latents.
    python tools/decoder_time.py [cycles]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, batch_norm=False).to(dev).eval()
z_q = 0.1 * torch.randn(n, 16, 32, device=dev)
flop = n * 16 * (2 * 32 * 512 + 16 * 2 * 1536 * 512 + 2 * 512 * 2560 + 5 * 2 * 512 * 5)
def timed(fn, reps=5):
    for _ in range(2): out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
with torch.no_grad():
    model.decoder_mode = "fused_bf16"
    ms_f, got = timed(lambda: model.decode(z_q))
    print(f"fused_bf16 decoder: {ms_f:.3f} ms per {n} cycles = {flop / ms_f / 1e9:.0f} TFLOP/s, {n * 16 / ms_f / 1e3:.1f} M patches/s")
    model.decoder_mode = "fused_fp32"
    ms_s, got32 = timed(lambda: model.decode(z_q))
    print(f"fused_fp32 decoder: {ms_s:.3f} ms per {n} cycles = {3 * flop / ms_s / 1e9:.0f} issued TFLOP/s, {n * 16 / ms_s / 1e3:.1f} M patches/s")
    model.decoder_mode = "torch"
    torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
    want_s = model.decode(z_q[:1024])
    print(f"max |fused_fp32 - fp32| = {(got32[:1024] - want_s).abs().max().item():.3e} at max |x_hat| = {want_s.abs().max().item():.3e}")
    model.decoder_mode = "fused_bf16"
    if os.environ.get("DEC_NO_PROFILER") != "1":
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            model.decode(z_q); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=6, max_name_column_width=70))
        model.decoder_mode = "torch"
        torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
        small = z_q[: min(n, 4096)]
        ms_t, want = timed(lambda: model.decode(small), reps=2)
        print(f"PyTorch fp32 decoder: {ms_t * n / small.shape[0]:.1f} ms per {n} cycles (timed on {small.shape[0]})")
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_a, _ = timed(lambda: model.decode(small), reps=2)
        print(f"PyTorch bf16-autocast decoder: {ms_a * n / small.shape[0]:.1f} ms per {n} cycles")
        dev_abs = (got[: small.shape[0]] - want).abs().max().item()
        print(f"max |fused - fp32| = {dev_abs:.3e} at max |x_hat| = {want.abs().max().item():.3e}")
