"""GPU box: does running the fused encoder chain over L2-sized sub-chunks (CUDA-graph replays) beat one pass per
layer over the whole 65536-cycle chunk?  (activations of a sub-chunk stay in the 126 MB L2 between the 19 kernels)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
model.encoder_mode = "fused_bf16"
n = 1 << 16
x = torch.randn(n, 200, 2, device=dev)

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

ms, ref = timed(lambda: model.encode_ids(x))
print(f"whole chunk ({n} cycles): {ms:.2f} ms -> {n*16/ms/1e3:.1f} M patches/s", flush=True)
for sub in (592, 1184, 2368, 4736, 9472):
    xs = torch.zeros(sub, 200, 2, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2): model.encode_ids(xs)
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            ids_s = model.encode_ids(xs)
    except Exception as e:
        print("capture failed:", repr(e)[:300]); break
    ids = torch.empty(n, 16, dtype=torch.int64, device=dev)
    def run():
        for o in range(0, n, sub):
            m = min(sub, n - o)
            xs[:m].copy_(x[o:o + m])
            g.replay()
            ids[o:o + m].copy_(ids_s[:m])
        return ids
    ms, out = timed(run)
    print(f"sub-chunks of {sub} cycles ({sub*16} tokens), graph replay: {ms:.2f} ms -> {n*16/ms/1e3:.1f} M patches/s, "
          f"ids equal: {bool(torch.equal(out, ref))}", flush=True)
