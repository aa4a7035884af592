"""GPU box: BASELINE configs[0] -- the VQ-VAE-Patch reconstruction forward (train_reconstruction_embedding.py path:
VQVAEPatch.forward -> embedding_loss, x_hat, perplexity) at batch 256, repo-default model, fp32, eval, no_grad.
SURVEY.md section 8(d) config 1 measured the unmodified reference at 955 ms on 8 host cores (4.3 k patches/s); this
prints the same call on one B200 through the drop-in module (encoder/decoder: PyTorch layers; quantiser: fused kernel),
eager and replayed from a CUDA graph."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
x = torch.randn(256, 200, 2, generator=torch.Generator().manual_seed(0)).to(dev)
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
def timeit(fn, reps=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = {}
with torch.no_grad():
    out["eager_ms"] = timeit(lambda: model(x))
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): model(x)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        res = model(x)
    out["graph_ms"] = timeit(g.replay)
    loss, x_hat, ppl = res
out.update({"batch": 256, "patches": 256 * 16, "patches_per_s_eager": 256 * 16 / (out["eager_ms"] * 1e-3),
            "patches_per_s_graph": 256 * 16 / (out["graph_ms"] * 1e-3), "embedding_loss": float(loss), "perplexity": float(ppl),
            "x_hat_shape": list(x_hat.shape), "reference_cpu_ms_survey": 955.0})
print(json.dumps(out))
