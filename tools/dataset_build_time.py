"""GPU box: the reference's data-set geometry -- windows of 20 cycles with a stride of ONE cycle over a stream of
cycles (dataloader/asimow_dataloader.py:185-206), loader batches of 512 windows (train_transformer_mtasks.py:214) --
through LatentSpaceEncoder.create_latent_space_dataset_VQ_VAE_IDs with dedupe off / per batch / per data set.  The
loader batches are overlapping strided views of one host stream (what a DataLoader's collate would materialise), so
every batch goes through the prefetcher's pinned staging copy.  Prints wall-clock ms, windows/s and patches/s (counted
the way the reference counts: every cycle of every window) per mode, and checks that the three id arrays are equal;
then the same data set through create_latent_space_dataset_from_cycles (the cycle stream in, every cycle encoded once).

    python tools/dataset_build_time.py [n_cycles [group_cycles]]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vqb200
from vqb200.dataloader import LatentSpaceEncoder

dev = torch.device("cuda:0")
n_cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
SEQ, BATCH = 20, 512
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
enc = LatentSpaceEncoder(model, window_size=200, device=str(dev))
if len(sys.argv) > 2:
    enc.group_cycles = int(sys.argv[2])
stream = torch.randn(n_cycles * 200, 2, generator=torch.Generator().manual_seed(1))
n_windows = n_cycles - SEQ + 1
windows = stream.as_strided((n_windows, SEQ * 200, 2), (200 * 2, 2, 1))
loader = [windows[i:i + BATCH] for i in range(0, n_windows, BATCH)]
if os.environ.get("SHUFFLE") == "1":       # the reference's train loader is shuffled (dataloader/asimow_dataloader.py:141-142)
    perm = torch.randperm(n_windows, generator=torch.Generator().manual_seed(2))
    loader = [windows[perm[i:i + BATCH]] for i in range(0, n_windows, BATCH)]      # materialised, contiguous, pageable
if os.environ.get("PINNED") == "1":        # what DataLoader(pin_memory=True) hands over: contiguous pinned batches, no staging copy
    loader = [b.contiguous().pin_memory() for b in loader]
out, ref = {"pinned": os.environ.get("PINNED") == "1", "shuffled": os.environ.get("SHUFFLE") == "1", "group_cycles": enc.group_cycles, "n_cycles": n_cycles, "n_windows": n_windows, "batches": len(loader)}, None
for mode in (False, True, "dataset"):
    enc.dedupe = mode
    best = None
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ids, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=SEQ, has_patch_embed=True, no_labels=True)
        dt = time.perf_counter() - t0
        best = dt if best is None or (rep and dt < best) else best
    if ref is None:
        ref = ids
    out[str(mode)] = {"ms": best * 1e3, "windows_per_s": n_windows / best, "patches_per_s": n_windows * SEQ * 16 / best,
                      "ms_per_batch": best * 1e3 / len(loader), "ids_equal_to_plain": bool(np.array_equal(ids, ref))}
if os.environ.get("SHUFFLE") == "1":
    print(json.dumps(out)); sys.exit(0)
# the same data set from the cycle stream: every cycle encoded once, windows as a sliding view of the ids (the reference's
# create_sequence_ds keeps n - seq_len windows: the last window of the loader above is not part of it)
cycles = stream.view(n_cycles, 200, 2)
best = None
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ids_s, _ = enc.create_latent_space_dataset_from_cycles(cycles, None, seq_len=SEQ, has_patch_embed=True)
    dt = time.perf_counter() - t0
    best = dt if best is None or (rep and dt < best) else best
nw = ids_s.shape[0]
out["from_cycles"] = {"ms": best * 1e3, "windows_per_s": nw / best, "patches_per_s": nw * SEQ * 16 / best,
                      "ids_equal_to_plain": bool(np.array_equal(ids_s, ref[:nw])), "windows": nw}
best = None
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    view, _ = enc.create_latent_space_dataset_from_cycles(cycles, None, seq_len=SEQ, has_patch_embed=True, materialize=False)
    dt = time.perf_counter() - t0
    best = dt if best is None or (rep and dt < best) else best
out["from_cycles_view"] = {"ms": best * 1e3, "patches_per_s": nw * SEQ * 16 / best, "ids_equal_to_plain": bool(np.array_equal(view, ref[:nw]))}
print(json.dumps(out))
