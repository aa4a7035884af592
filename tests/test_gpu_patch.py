"""VQ-VAE-Patch with the fused VQ on the GPU against outputs of the unmodified reference
model (tests/golden/patch_golden.npz): full forward, one training step, the id-encode call
and the bulk latent-dataset loops (dataloader/latentspace_dataloader.py:144-263)."""
import numpy as np
import pytest
import torch

import cases as C
import vqb200
from vqb200.dataloader import LatentSpaceEncoder

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _load(case, patch_golden):
    model = vqb200.VQVAEPatch(hidden_dim=case["hidden_dim"], input_dim=case["input_dim"],
                              num_embeddings=case["num_embeddings"], embedding_dim=case["embedding_dim"],
                              n_resblocks=case["n_resblocks"], learning_rate=1e-3, dropout_p=0.0,
                              patch_size=case["patch_size"], seq_len=case["seq_len"],
                              batch_norm=case["batch_norm"], beta=case["beta"])
    pre = f"{case['name']}/sd/"
    sd = {k[len(pre):]: torch.from_numpy(patch_golden[k]) for k in patch_golden.files if k.startswith(pre)}
    model.load_state_dict(sd, strict=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return model.to(DEV)


@pytest.mark.parametrize("case", C.PATCH_CASES, ids=[c["name"] for c in C.PATCH_CASES])
def test_forward_matches_reference(case, patch_golden):
    name = case["name"]
    model = _load(case, patch_golden).eval()
    x = torch.from_numpy(C.make_cycles(case)).to(DEV)
    with torch.no_grad():
        z_e = model.encode(x)
        emb_loss, x_hat, ppl = model(x)
        ids = model.encode_ids(x)
    assert z_e.is_contiguous()                     # (B, T, D) rows: straight into the tcgen05 path
    np.testing.assert_allclose(z_e.contiguous().cpu().numpy(), patch_golden[f"{name}/z_e"], rtol=1e-4, atol=1e-5)
    ref_ids = patch_golden[f"{name}/idx"].astype(np.int64)
    got = ids.cpu().numpy().reshape(-1)
    assert ids.shape == (x.shape[0], model.enc_out_len)
    # the encoder runs through cuBLAS here and MKL in the fixture: ids may differ only where the
    # two nearest codes are equidistant to within the encoder's fp32 rounding (tolerance 1e-5 relative)
    E = patch_golden[f"{name}/sd/vector_quantization.embedding.weight"]
    assert C.unexplained_mismatches(patch_golden[f"{name}/z_e"], E, got, ref_ids, rel=1e-5) == 0
    assert (got == ref_ids).mean() >= 0.99
    if np.array_equal(got, ref_ids):
        assert emb_loss.item() == pytest.approx(float(patch_golden[f"{name}/emb_loss"]), rel=1e-4)
        assert ppl.item() == pytest.approx(float(patch_golden[f"{name}/perplexity"]), rel=1e-4)
        np.testing.assert_allclose(x_hat.cpu().numpy(), patch_golden[f"{name}/x_hat"], rtol=1e-3, atol=1e-4)


def test_training_step_matches_reference(patch_golden):
    case = C.PATCH_CASES[0]
    name = case["name"]
    model = _load(case, patch_golden).train()
    x = torch.from_numpy(C.make_cycles(case)).to(DEV)
    out = model.training_step(x, 0)
    out["loss"].backward()
    assert out["loss"].item() == pytest.approx(float(patch_golden[f"{name}/train_total"]), rel=1e-4)
    assert out["recon_error"].item() == pytest.approx(float(patch_golden[f"{name}/train_recon"]), rel=1e-4)
    for key, param in (("grad_codebook", model.vector_quantization.embedding.weight),
                       ("grad_enc_proj", model.encoder[1].shared_conv.weight),
                       ("grad_patch_proj", model.patch_embed.proj.weight)):
        ref = patch_golden[f"{name}/{key}"]
        got = param.grad.cpu().numpy()
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=2e-3, atol=2e-4 * np.abs(ref).max())
    # the two outer taps of the k=3 encoder convs only ever see zero padding (appendix A.7)
    w = model.encoder[0].shared_conv[0].block[1].weight.grad
    assert torch.count_nonzero(w[:, :, 0]) == 0 and torch.count_nonzero(w[:, :, 2]) == 0
    opt = model.configure_optimizers()
    opt.step()


def test_bulk_loops_match_the_reference_loops(patch_golden, bulk_golden):
    """a15: the three bulk builders against arrays the UNMODIFIED reference loops produced
    (dataloader/latentspace_dataloader.py:171-263 run by oracle/make_golden.py on overlapping windows)."""
    case = C.BULK_CASE
    name = case["name"]
    mcase = next(c for c in C.PATCH_CASES if c["name"] == case["model"])
    model = _load(mcase, patch_golden).eval()
    E = patch_golden[f"{mcase['name']}/sd/vector_quantization.embedding.weight"]
    win, labels, slices = C.make_windows(case)
    win_t, lab_t = torch.from_numpy(win), torch.from_numpy(labels)
    loader = [(win_t[lo:hi], lab_t[lo:hi]) for lo, hi in slices]
    ref_ids, ref_zq = bulk_golden[f"{name}/ids"], bulk_golden[f"{name}/zq"]
    for dedupe in (False, True):
        enc = LatentSpaceEncoder(model, window_size=200, device=DEV)
        enc.dedupe = dedupe
        ids, y = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=case["seq_len"], has_patch_embed=True)
        assert ids.shape == ref_ids.shape and ids.dtype == ref_ids.dtype
        assert np.array_equal(y, bulk_golden[f"{name}/labels"]) and y.dtype == bulk_golden[f"{name}/labels"].dtype
        # fp32 encoder on cuBLAS here, MKL in the fixture: equal ids except fp32 near-ties (none on this fixture)
        assert C.unexplained_mismatches(bulk_golden[f"{name}/z_e"], E, ids.reshape(-1), ref_ids.reshape(-1), rel=1e-5) == 0
        assert (ids == ref_ids).mean() >= 0.995
    ar, y3 = enc.create_latent_space_dataset_VQ_VAE_autoreggressive(
        [w for w, _ in loader], seq_len=case["seq_len"], has_patch_embed=True, task="autoregressive_ids")
    assert ar.shape == bulk_golden[f"{name}/ar_ids"].shape and np.array_equal(ar, bulk_golden[f"{name}/ar_ids"])
    assert np.array_equal(y3, bulk_golden[f"{name}/ar_labels"])
    zq, y2 = enc.create_latent_space_dataset_VQ_VAE(loader, seq_len=case["seq_len"], has_patch_embed=True)
    assert zq.shape == ref_zq.shape and zq.dtype == ref_zq.dtype
    np.testing.assert_allclose(zq, ref_zq, rtol=0, atol=1e-6)
    assert np.array_equal(y2, bulk_golden[f"{name}/labels_zq"])


def test_bulk_loops_equal_per_cycle_reference_loop(patch_golden):
    """create_latent_space_dataset_VQ_VAE_IDs over windows of 3 cycles == encoding each cycle
    slice on its own, the way the reference loop does (:225-238)."""
    case = C.PATCH_CASES[0]
    model = _load(case, patch_golden).eval()
    enc = LatentSpaceEncoder(model, window_size=200, device=DEV)
    rs = np.random.RandomState(5)
    seq_len, n_windows = 3, 10
    data = torch.from_numpy(rs.standard_normal((n_windows, seq_len * 200, 2)).astype(np.float32))
    labels = torch.from_numpy(rs.randint(0, 2, n_windows).astype(np.float32))
    loader = [(data[i:i + 4], labels[i:i + 4]) for i in range(0, n_windows, 4)]
    ids, y = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True)
    assert ids.shape == (n_windows, seq_len, model.enc_out_len) and ids.dtype == np.int64
    assert np.array_equal(y, labels.numpy().astype(np.float64))
    want = np.empty_like(ids)
    with torch.no_grad():
        for i in range(seq_len):
            x_i = data[:, i * 200:(i + 1) * 200, :].clone().to(DEV)
            want[:, i, :] = enc.get_latent_space_IDs(x_i, True).cpu().numpy().reshape(n_windows, -1)
    assert np.array_equal(ids, want)
    assert int(enc.code_counts.sum()) == ids.size
    flat, y2 = enc.create_latent_space_dataset_VQ_VAE_autoreggressive(
        [d for d, _ in loader], seq_len=seq_len, has_patch_embed=True, task="autoregressive_ids")
    assert flat.shape == (n_windows, seq_len * model.enc_out_len) and np.array_equal(flat, ids.reshape(n_windows, -1))
    assert np.array_equal(y2, np.zeros(n_windows))
    zq, y3 = enc.create_latent_space_dataset_VQ_VAE(loader, seq_len=seq_len, has_patch_embed=True)
    assert zq.shape == (n_windows, seq_len, model.embedding_dim * model.enc_out_len) and zq.dtype == np.float64
    E = model.vector_quantization.embedding.weight.detach().cpu().numpy()
    np.testing.assert_allclose(zq.reshape(n_windows, seq_len, model.enc_out_len, -1), E[ids], rtol=0, atol=1e-6)


def test_bulk_loop_prefetch_over_every_kind_of_loader_batch(patch_golden):
    """The bulk loops copy batch i + 1 to the device on a copy stream while batch i is encoded (_DevicePrefetcher):
    pageable and pinned tensors, numpy arrays, windows longer than seq_len cycles (strided source), batches of changing
    size (slots are re-sized), a single batch and an empty loader, consecutive batches grouped into one encoder call or
    not -- ids equal to one direct encode call per batch, many batches so that every slot is recycled several times."""
    case = C.PATCH_CASES[0]
    model = _load(case, patch_golden).eval()
    enc = LatentSpaceEncoder(model, window_size=200, device=DEV)
    rs = np.random.RandomState(11)
    seq_len = 2
    sizes = [3, 7, 1, 16, 5, 16, 2, 9, 16, 4, 11, 6]
    kinds = ["pageable", "pinned", "numpy", "long", "long_pinned", "cuda"]
    loader, direct = [], []
    for i, b in enumerate(sizes):
        kind = kinds[i % len(kinds)]
        extra = 37 if kind.startswith("long") else 0              # the loop reads the first seq_len * 200 samples only
        x = torch.from_numpy(rs.standard_normal((b, seq_len * 200 + extra, 2)).astype(np.float32))
        with torch.no_grad():
            cyc = x[:, : seq_len * 200, :].reshape(b * seq_len, 200, 2).to(DEV)
            direct.append(enc.get_latent_space_IDs(cyc, True).cpu().numpy().reshape(b, seq_len, -1))
        if kind in ("pinned", "long_pinned"):
            x = x.pin_memory()
        elif kind == "numpy":
            x = x.numpy()
        elif kind == "cuda":
            x = x.to(DEV)
        loader.append(x)
    want = np.concatenate(direct, axis=0)
    for group in (65536, 65536, 1, 30, 33):                      # one call per run of host batches ... one per loader batch
        enc.group_cycles = group
        ids, y = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
        assert ids.shape == want.shape and np.array_equal(ids, want)
        assert np.array_equal(y, np.zeros(sum(sizes)))
    enc.group_cycles = 24
    one, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader[3:4], seq_len=seq_len, has_patch_embed=True, no_labels=True)
    assert np.array_equal(one, direct[3])
    none, y0 = enc.create_latent_space_dataset_VQ_VAE_IDs([], seq_len=seq_len, has_patch_embed=True, no_labels=True)
    assert none.shape == (0, seq_len, model.enc_out_len) and y0.shape == (0,)
    labelled = [(x, np.full(len(x), float(i))) for i, x in enumerate(loader)]
    zq, yl = enc.create_latent_space_dataset_VQ_VAE(labelled, seq_len=seq_len, has_patch_embed=True)
    E = model.vector_quantization.embedding.weight.detach().cpu().numpy()
    np.testing.assert_allclose(zq.reshape(sum(sizes), seq_len, model.enc_out_len, -1), E[want], rtol=0, atol=1e-6)
    assert np.array_equal(yl, np.concatenate([np.full(b, float(i)) for i, b in enumerate(sizes)]))


def test_data_sets_from_the_cycle_stream_match_the_reference_window_builder(patch_golden, bulk_golden):
    """create_latent_space_dataset_from_cycles (every cycle encoded once, windows as a sliding view of the result) against
    the arrays the unmodified reference produced with its own window builder (ASIMoWDataLoader.create_sequence_ds,
    dataloader/asimow_dataloader.py:185-206) and its own loops over those windows (tests/golden: bulk_overlap/seq_*)."""
    case = C.BULK_CASE
    name = case["name"]
    model = _load(next(c for c in C.PATCH_CASES if c["name"] == case["model"]), patch_golden).eval()
    enc = LatentSpaceEncoder(model, window_size=200, device=DEV, encoder_mode="torch")
    stream, cycle_labels = C.make_stream(case)
    calls = []
    inner = enc.get_latent_space_IDs
    enc.get_latent_space_IDs = lambda x, p=False: (calls.append(x.shape[0]), inner(x, p))[1]
    for cyc in (stream, stream.astype(np.float64), torch.from_numpy(stream)):
        ids, y = enc.create_latent_space_dataset_from_cycles(cyc, cycle_labels, seq_len=case["seq_len"], has_patch_embed=True,
                                                             batch=4)
        assert ids.dtype == np.int64 and np.array_equal(ids, bulk_golden[f"{name}/seq_ids"])
        assert y.dtype == np.float64 and np.array_equal(y, bulk_golden[f"{name}/seq_labels"])
    assert calls == [4, 4, 4, 3] * 3                    # 15 cycles, each encoded once (the reference: 11 windows x 4)
    ar, y0 = enc.create_latent_space_dataset_from_cycles(stream, None, seq_len=case["seq_len"], has_patch_embed=True, kind="ar_ids")
    assert np.array_equal(ar, bulk_golden[f"{name}/seq_ids"].reshape(ar.shape[0], -1)) and np.array_equal(y0, np.zeros(len(ar)))
    zq, y2 = enc.create_latent_space_dataset_from_cycles(stream, cycle_labels, seq_len=case["seq_len"], has_patch_embed=True,
                                                         kind="latents")
    assert zq.dtype == np.float64 and zq.shape == bulk_golden[f"{name}/seq_zq"].shape
    np.testing.assert_allclose(zq, bulk_golden[f"{name}/seq_zq"], rtol=0, atol=1e-6)
    assert np.array_equal(y2, bulk_golden[f"{name}/seq_labels"])
    # the windows of the older fixture (n - seq_len + 1 of them) are the same sliding view, one window longer
    assert np.array_equal(bulk_golden[f"{name}/ids"][: len(ids)], ids)
    # seq_len = 1: the cycles themselves (dataloader/asimow_dataloader.py:173); fewer cycles than a window: nothing
    one, y1 = enc.create_latent_space_dataset_from_cycles(stream, cycle_labels, seq_len=1, has_patch_embed=True)
    assert one.shape == (len(stream), 1, model.enc_out_len) and np.array_equal(y1, cycle_labels)
    assert np.array_equal(one[:, 0, :][np.arange(len(ids))[:, None] + np.arange(case["seq_len"])[None, :]], ids)
    none, yn = enc.create_latent_space_dataset_from_cycles(stream[:3], cycle_labels[:3], seq_len=4, has_patch_embed=True)
    assert none.shape == (0, 4, model.enc_out_len) and yn.shape == (0,)
    with pytest.raises(ValueError):
        enc.create_latent_space_dataset_from_cycles(stream, cycle_labels[:-1], seq_len=2, has_patch_embed=True)
    with pytest.raises(ValueError):
        enc.create_latent_space_dataset_from_cycles(stream, None, seq_len=2, has_patch_embed=True, kind="one_hot")



    """Overlapping windows (stride of one cycle): with `dedupe = True` every distinct cycle is encoded once and the ids
    array is identical to the plain loop's (SURVEY.md section 8(f) row 2)."""
    case = C.PATCH_CASES[0]
    model = _load(case, patch_golden).eval()
    enc = LatentSpaceEncoder(model, window_size=200, device=DEV)
    rs = np.random.RandomState(9)
    seq_len, n_windows = 5, 24
    stream = torch.from_numpy(rs.standard_normal((n_windows + seq_len - 1, 200, 2)).astype(np.float32))
    data = torch.stack([stream[i:i + seq_len].reshape(seq_len * 200, 2) for i in range(n_windows)])
    loader = [data[i:i + 8] for i in range(0, n_windows, 8)]
    plain, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
    enc.group_cycles = 1                    # one encoder call per loader batch (the default groups consecutive batches)
    assert np.array_equal(plain, enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True,
                                                                            no_labels=True)[0])
    calls = []
    inner = enc.get_latent_space_IDs
    enc.get_latent_space_IDs = lambda x, p=False: (calls.append(x.shape[0]), inner(x, p))[1]
    enc.dedupe = True
    fast, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
    assert np.array_equal(plain, fast)
    assert calls == [12, 12, 12]            # 8 windows x 5 cycles = 40 rows hold 12 distinct cycles per batch
    # once per data set: 28 distinct cycles in all, the second and third batch only add their 8 new ones each
    calls.clear()
    enc.dedupe = "dataset"
    whole, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
    assert np.array_equal(plain, whole)
    assert calls == [12, 8, 8] and len(enc.cycle_cache) == 28
    # grouped (the default: up to 65536 cycles per call): the three batches are one call, and the per-call de-duplication
    # already sees the overlap between them
    enc.group_cycles = 65536
    for mode in (True, "dataset"):
        calls.clear()
        enc.dedupe = mode
        grouped, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
        assert np.array_equal(plain, grouped) and calls == [28]
    enc.group_cycles = 80                   # two loader batches per call
    calls.clear()
    grouped, _ = enc.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=seq_len, has_patch_embed=True, no_labels=True)
    assert np.array_equal(plain, grouped) and calls == [20, 8]


def test_on_the_fly_tokenizer_matches_the_reference_dataset(patch_golden, bulk_golden):
    """f4: windows -> (x, cond, y) in one call against MyLatentAutoregressiveDataset of the UNMODIFIED reference
    (dataloader/base_dataloader.py:74-110) built from the ids its own bulk loop produced (tests/golden/bulk_golden.npz)."""
    from vqb200.dataloader import OnTheFlyTokenizer
    from vqb200 import ops
    case = C.BULK_CASE
    name = case["name"]
    mcase = next(c for c in C.PATCH_CASES if c["name"] == case["model"])
    model = _load(mcase, patch_golden).eval()
    win, labels, _ = C.make_windows(case)
    ref_x, ref_y = bulk_golden[f"{name}/ar_ds_x"], bulk_golden[f"{name}/ar_ds_y"]
    # the reference's start / end tokens follow the largest id present in ITS data set
    max_token = int(bulk_golden[f"{name}/ar_ids"].max())
    tok = OnTheFlyTokenizer(model, window_size=200, device=DEV, max_token=max_token)
    assert tok.num_classes == int(bulk_golden[f"{name}/ar_ds_num_classes"])
    x, cond, y = tok(torch.from_numpy(win), torch.from_numpy(labels), n_cycles=case["seq_len"])
    assert x.dtype == torch.int64 and y.dtype == torch.int64 and cond.dtype == torch.int64
    assert np.array_equal(x.cpu().numpy(), ref_x) and np.array_equal(y.cpu().numpy(), ref_y)
    assert np.array_equal(cond.cpu().numpy(), bulk_golden[f"{name}/ar_ds_cond"])
    # default: num_embeddings + 2 classes, the transformer's own assumption (train_transformer_mtasks.py:146)
    tok2 = OnTheFlyTokenizer(model, window_size=200, device=DEV)
    assert (tok2.start_token, tok2.end_token, tok2.num_classes) == (model.num_embeddings, model.num_embeddings + 1,
                                                                    model.num_embeddings + 2)
    # the kernel alone, ragged sizes
    ids = torch.randint(0, 50, (7, 33), device=DEV)
    xx, yy = ops.ar_pairs(ids, 50, 51)
    assert torch.equal(xx[:, 1:], ids) and torch.equal(yy[:, :-1], ids)
    assert bool((xx[:, 0] == 50).all()) and bool((yy[:, -1] == 51).all())
    with pytest.raises(RuntimeError):
        ops.ar_pairs(ids.cpu(), 50, 51)


def test_row_keys_and_dedupe_kernels_against_the_host_formulas(monkeypatch):
    """vqb_row_keys == the torch evaluation of the same sums (wrapping int64), row widths with and without the 16-byte path,
    -0.0 / NaN payloads kept apart; vqb_dedupe_first == first occurrence by key pair (checked against a Python dict), on
    many duplicates, and the whole dedupe_rows on CUDA == the CPU path, incl. useless hashes (one group, word-for-word
    split) and a key-0-only collision."""
    from vqb200 import ops
    from vqb200.dataloader import latentspace_dataloader as L
    g = torch.Generator().manual_seed(21)
    for n, words in ((1, 4), (37, 400), (1000, 402), (5000, 7), (4096, 400)):
        rows = torch.randn(n, words, generator=g)
        if n > 8:
            rows[3, 0], rows[5, 0] = 0.0, -0.0
            rows[7] = rows[2]
            rows[8, 1] = float("nan")
        want = L.cycle_fingerprints(rows)                                  # CPU: torch ops
        got = L.cycle_fingerprints(rows.to(DEV))                           # CUDA: vqb_row_keys
        assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)
    # first occurrence per key pair
    stream = torch.randn(700, 50, generator=g)
    idx = torch.randint(0, 700, (20000,), generator=g)
    rows = stream[idx].to(DEV)
    keys = L.cycle_fingerprints(rows)
    first = ops.dedupe_first(keys).cpu().tolist()
    seen, want_first = {}, []
    for i, k in enumerate(map(tuple, keys.cpu().tolist())):
        want_first.append(seen.setdefault(k, i))
    assert first == want_first
    assert ops.dedupe_first(keys).cpu().tolist() == first                  # deterministic
    for r in (rows, rows[:1], rows[:0], torch.randn(300, 200, 2, generator=g).to(DEV)):
        rep, inv = L.dedupe_rows(r)
        rep_c, inv_c = L.dedupe_rows(r.cpu())
        assert torch.equal(rep.cpu(), torch.sort(rep_c).values)            # the same representatives (first occurrences)
        if r.shape[0]:
            assert torch.equal(r[rep[inv]].view(torch.int32), r.view(torch.int32))
            assert torch.equal(rep, torch.sort(rep).values) and int(inv.max()) == rep.numel() - 1
    # key 0 equal for every row, key 1 honest: rows fall back to themselves unless both keys agree
    honest = L._hash_weights
    monkeypatch.setattr(L, "_hash_weights", lambda width, device, seed: (
        torch.zeros(width, dtype=torch.int64, device=device) if seed == L._HASH_SEEDS[0] else honest(width, device, seed)))
    L._weight_cache.clear()
    small = torch.randn(64, 12, generator=g)
    small[40] = small[10]
    small[50] = small[0]
    rep, inv = L.dedupe_rows(small.to(DEV))
    # (every row shares key 0 with row 0: only rows whose key 1 equals row 0's -- row 50 -- are grouped; 10 / 40 stay apart)
    assert torch.equal(small.to(DEV)[rep[inv]], small.to(DEV)) and rep.numel() == 63
    # both hashes useless: one group, split again word for word
    monkeypatch.setattr(L, "_hash_weights", lambda width, device, seed: torch.zeros(width, dtype=torch.int64, device=device))
    L._weight_cache.clear()
    rep, inv = L.dedupe_rows(small.to(DEV))
    assert torch.equal(small.to(DEV)[rep[inv]], small.to(DEV)) and rep.numel() == 63      # only row 50 == row 0 is found
    L._weight_cache.clear()


def test_forward_without_the_one_hot_is_the_same_call():
    """VectorQuantizer.forward(z, need_one_hot=False) -- what VQVAEPatch.forward uses, the reference drops min_encodings
    there (model/vq_vae_patch_embedd.py:161) -- returns the same loss, z_q, perplexity and ids, and None for the one-hot."""
    torch.manual_seed(3)
    vq = vqb200.VectorQuantizer(64, 16, 0.25).to(DEV)
    z = torch.randn(7, 16, 16, device=DEV) * 0.02
    full = vq(z)
    lean = vq(z, need_one_hot=False)
    assert lean[3] is None and full[3].shape == (7 * 16, 64)
    for a, b in zip((full[0], full[1], full[2], full[4]), (lean[0], lean[1], lean[2], lean[4])):
        assert torch.equal(a, b)
    assert vqb200.VectorQuantizer(64, 16, 0.25, one_hot="none").to(DEV)(z, need_one_hot=True)[3].shape == (7 * 16, 64)
