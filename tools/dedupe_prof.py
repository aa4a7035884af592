"""GPU box: where a batch of the de-duplicating bulk loop spends its time (512 windows x 20 cycles, stride of one cycle)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vqb200
from vqb200.dataloader import latentspace_dataloader as L
dev = "cuda:0"
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
enc = L.LatentSpaceEncoder(model, window_size=200, device=dev)
n_cycles, SEQ, BATCH = 40000, 20, 512
GROUP = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
stream = torch.randn(n_cycles * 200, 2)
nw = n_cycles - SEQ + 1
windows = stream.as_strided((nw, SEQ * 200, 2), (400, 2, 1))
loader = [windows[i:i + BATCH] for i in range(0, nw, BATCH)]
def wall(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3 / len(loader)
def only_prefetch():
    for cyc, items in L._DevicePrefetcher(loader, dev, SEQ, 200, no_labels=True, group_cycles=GROUP):
        pass
def prefetch_dedupe():
    for cyc, items in L._DevicePrefetcher(loader, dev, SEQ, 200, no_labels=True, group_cycles=GROUP):
        L.dedupe_rows(cyc)
def prefetch_dedupe_gather():
    for cyc, items in L._DevicePrefetcher(loader, dev, SEQ, 200, no_labels=True, group_cycles=GROUP):
        rep, inv = L.dedupe_rows(cyc); u = cyc[rep]
def prefetch_dedupe_encode():
    with torch.no_grad():
        for cyc, items in L._DevicePrefetcher(loader, dev, SEQ, 200, no_labels=True, group_cycles=GROUP):
            rep, inv = L.dedupe_rows(cyc); ids = enc.get_latent_space_IDs(cyc[rep], True).view(rep.numel(), -1)[inv]
small = torch.randn(531, 200, 2, device=dev)
def encode_small():
    with torch.no_grad():
        for _ in loader:
            enc.get_latent_space_IDs(small, True)
big = torch.randn(10240, 200, 2, device=dev)
def encode_big():
    with torch.no_grad():
        for _ in loader:
            enc.get_latent_space_IDs(big, True)
for name, fn in (("prefetch only", only_prefetch), ("+ dedupe_rows", prefetch_dedupe), ("+ gather", prefetch_dedupe_gather),
                 ("+ encode + scatter", prefetch_dedupe_encode), ("encode 531 cycles alone", encode_small),
                 ("encode 10240 cycles alone", encode_big)):
    print(f"{name:44s} {wall(fn):.3f} ms per batch")

# the kernels alone (CUDA events, 65536 cycles = 105 MB: larger than... no, smaller than L2 -- so two buffers alternate)
from vqb200 import ops
rows = [torch.randn(65536 * 4, 400, device=dev) for _ in range(2)]        # 2 x 419 MB: larger than L2
mult = L._hash_weight_pair(400, torch.device(dev))
def ev_time(fn, reps=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = ev_time(lambda i: ops.row_keys(rows[i & 1], mult))
nbytes = rows[0].numel() * 4
print(f"vqb_row_keys      {rows[0].shape[0]} rows x 1600 B: {ms:.4f} ms = {nbytes / ms / 1e6:.0f} GB/s")
keys = ops.row_keys(rows[0], mult)
dup = keys[torch.randint(0, 20000, (keys.shape[0],), device=dev)].contiguous()        # ~13 rows per distinct key
for name, k in (("all distinct", keys), ("20000 distinct", dup)):
    ms = ev_time(lambda i: ops.dedupe_first(k))
    print(f"vqb_dedupe_first  {k.shape[0]} keys, {name}: {ms:.4f} ms = {k.shape[0] / ms / 1e3:.0f} M rows/s")
