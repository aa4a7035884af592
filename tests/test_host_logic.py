"""Host-side logic of the drop-in that needs no GPU: argument checking, layout analysis,
module surface and state-dict contract."""
import numpy as np
import pytest
import torch

import vqb200
from vqb200 import ops


def test_constructor_surface_and_init_range():
    torch.manual_seed(0)
    vq = vqb200.VectorQuantizer(256, 32, 0.25)
    assert (vq.n_e, vq.e_dim, vq.beta) == (256, 32, 0.25)
    assert isinstance(vq.embedding, torch.nn.Embedding)
    w = vq.embedding.weight
    assert w.shape == (256, 32) and w.dtype == torch.float32 and w.requires_grad
    assert w.abs().max().item() <= 1.0 / 256          # model/vector_quantizer.py:74
    assert list(vq.state_dict().keys()) == ["embedding.weight"]   # the only persistent key
    assert [n for n, _ in vq.named_parameters()] == ["embedding.weight"]
    assert hasattr(vq, "device")


def test_same_seed_same_codebook_as_reference_init():
    """uniform_(-1/K, 1/K) right after nn.Embedding's own normal_ init: same RNG consumption
    as the reference constructor, so seeded runs start from the same codebook."""
    torch.manual_seed(123)
    a = vqb200.VectorQuantizer(64, 8, 0.25).embedding.weight.detach().clone()
    torch.manual_seed(123)
    emb = torch.nn.Embedding(64, 8)
    emb.weight.data.uniform_(-1.0 / 64, 1.0 / 64)
    assert torch.equal(a, emb.weight.detach())


def test_cpu_tensors_raise_no_fallback():
    vq = vqb200.VectorQuantizer(16, 8, 0.25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vq(torch.randn(4, 8))


def test_dtype_mismatch_raises_runtime_error():
    vq = vqb200.VectorQuantizer(16, 8, 0.25)
    with pytest.raises(RuntimeError, match="float32"):
        vq(torch.randn(4, 8, dtype=torch.float64))


def test_bad_options():
    with pytest.raises(ValueError):
        vqb200.VectorQuantizer(16, 8, 0.25, one_hot="sparse")
    with pytest.raises(ValueError):
        vqb200.VectorQuantizer(16, 8, 0.25, path="triton")


def test_view_params_contiguous_and_permuted():
    z = torch.randn(6, 16, 32)
    t, n_outer, n_inner, so, si, sd = ops._view_params(z, 32)
    assert (n_outer * n_inner, so, sd) == (96, 32, 1) and t is z
    phys = torch.randn(6, 32, 16)
    view = phys.permute(0, 2, 1)                      # the encoder's layout
    t, n_outer, n_inner, so, si, sd = ops._view_params(view, 32)
    assert t is view and (n_outer, n_inner, so, si, sd) == (6, 16, 512, 1, 16)
    two_d = torch.randn(32, 10).t()                   # (10, 32) column-major
    t, n_outer, n_inner, so, si, sd = ops._view_params(two_d, 32)
    assert t is two_d and (n_outer, n_inner, so, sd) == (10, 1, 1, 10)
    odd = torch.randn(2, 4, 4, 64)[..., ::2]          # 4-D strided: falls back to one copy
    t, n_outer, n_inner, so, si, sd = ops._view_params(odd, 32)
    assert t.is_contiguous() and (n_outer * n_inner, so, sd) == (32, 32, 1)


def test_world_size_helpers_single_process():
    assert vqb200.get_world_size() == 1
    t = torch.arange(4)
    assert vqb200.all_reduce(t) is t


def test_lightning_hook_surface():
    m = vqb200.VQVAEPatch(hidden_dim=16, input_dim=2, num_embeddings=8, embedding_dim=4, n_resblocks=1,
                          learning_rate=1e-3, batch_norm=False)
    for hook in ("training_step", "validation_step", "test_step", "configure_optimizers", "_forward_setp"):
        assert callable(getattr(m, hook))
    opt = m.configure_optimizers()
    assert isinstance(opt, torch.optim.RAdam)
    ids = {id(p) for g in opt.param_groups for p in g["params"]}
    assert id(m.vector_quantization.embedding.weight) in ids
    assert m.enc_out_len == 16 and m.patch_size == 25 and m.embedding_dim == 4 and m.num_embeddings == 8
    hp = vars(m.hparams) if not isinstance(m.hparams, dict) else m.hparams
    assert hp["hidden_dim"] == 16 and hp["patch_size"] == 25 and hp["beta"] == 0.25
    with pytest.raises(NotImplementedError):
        vqb200.VQVAEPatch(hidden_dim=16, input_dim=2, num_embeddings=8, embedding_dim=4, n_resblocks=1,
                          learning_rate=1e-3, use_improved_vq=True)


def test_dedupe_rows_and_encode_unique():
    """Distinct cycles are found on bit patterns: windows with a stride of one cycle."""
    import torch
    from vqb200.dataloader import latentspace_dataloader as L
    g = torch.Generator().manual_seed(5)
    base = torch.randn(40, 200, 2, generator=g)
    base[7] = base[3]                                   # a genuine duplicate among the distinct cycles
    base[9, 0, 0] = 0.0
    base[11] = base[9]
    base[11, 0, 0] = -0.0                               # -0.0 != +0.0 bitwise: kept apart (never merged wrongly)
    windows = torch.stack([base[i:i + 20] for i in range(20)])          # (20 windows, 20 cycles, 200, 2)
    cycles = windows.reshape(-1, 200, 2)
    rep, inv = L.dedupe_rows(cycles)
    assert rep.numel() == 38                            # 39 cycles are touched, one pair is identical
    assert torch.equal(cycles[rep[inv]].view(torch.int32), cycles.view(torch.int32))
    calls = []
    def encode(c):                                      # a per-cycle function with T = 3 "tokens"
        calls.append(c.shape[0])
        return torch.stack([c.sum(dim=(1, 2)), c[:, 0, 0], c[:, -1, 1]], dim=1)
    out = L.encode_unique(encode, cycles)
    assert calls == [38] and torch.equal(out, encode(cycles))
    # no duplicates: one plain call
    calls.clear()
    L.encode_unique(encode, base[[0, 1, 2]])
    assert calls == [3]
    assert L.dedupe_rows(cycles[:0])[0].numel() == 0 and tuple(L.cycle_fingerprints(cycles[:0]).shape) == (0, 2)
    assert tuple(L.dedupe_rows(cycles[:0], return_keys=True)[2].shape) == (0, 2)


def test_dedupe_rows_survives_hash_collisions(monkeypatch):
    """With useless hashes every row lands in one group; the word-for-word check must split it again."""
    import torch
    from vqb200.dataloader import latentspace_dataloader as L
    monkeypatch.setattr(L, "_hash_weights", lambda width, device, seed: torch.zeros(width, dtype=torch.int64, device=device))
    g = torch.Generator().manual_seed(6)
    rows = torch.randn(10, 8, generator=g)
    rows[4] = rows[0]
    rep, inv = L.dedupe_rows(rows)
    assert torch.equal(rows[rep[inv]], rows)
    assert rep.numel() == 9                             # rows 0 and 4 share the first group, the other eight stand alone


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the
    contract's keys, whatever the libraries print."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "vq_encoded_patches_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_encoder_side_ops_refuse_cpu_tensors():
    """The patch-embedding and encoder-layer entry points have no CPU path either."""
    import torch
    from vqb200 import ops
    x = torch.randn(2, 200, 2)
    w = torch.randn(512, 1, 25)
    b = torch.randn(512)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.patch_embed(x, w, b, 25)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.token_bias_gelu(torch.randn(4, 512), b)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.token_linear(torch.randn(256, 512).to(torch.bfloat16), torch.randn(512, 512).to(torch.bfloat16), b)
    # a layout-only helper: on CPU tensors it is torch's own copy, the quantiser behind it still refuses them
    v = torch.randn(3, 32, 16).permute(0, 2, 1)
    assert torch.equal(ops.pack_rows(v, 32), v.contiguous())
    # the fp32-faithful (split) entry points likewise
    pair = torch.randn(256, 1024).to(torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.token_linear_split(pair, torch.randn(512, 1024).to(torch.bfloat16), b)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.token_conv_split(pair, torch.randn(512, 3 * 1024).to(torch.bfloat16), b)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.token_pair(torch.randn(4, 512))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.token_out_proj_pair(pair, torch.randn(5, 512), 0.0, 1)


def test_bf16_pairs_carry_fp32_values_to_2_pow_minus_17():
    """Host side of the fp32-faithful layers: weights are split once per weight version with torch ops (ops.bf16_pair /
    ops.conv_pair); hi + lo reproduces the fp32 value to 2^-17 relative, and the layouts are the ones the kernels index."""
    import torch
    from vqb200 import ops
    g = torch.Generator().manual_seed(5)
    w = torch.randn(64, 128, generator=g) * torch.logspace(-6, 3, 128)
    p = ops.bf16_pair(w)
    assert p.shape == (64, 256) and p.dtype == torch.bfloat16 and p.is_contiguous()
    assert torch.equal(p[:, :128], w.to(torch.bfloat16))
    val = p[:, :128].double() + p[:, 128:].double()
    assert bool(((val - w.double()).abs() <= 2.0 ** -17 * w.double().abs()).all())
    w3 = torch.randn(32, 64, 3, generator=g)
    c = ops.conv_pair(w3)
    assert c.shape == (32, 3 * 128)
    for tap in range(3):
        assert torch.equal(c[:, tap * 128:(tap + 1) * 128], ops.bf16_pair(w3[:, :, tap].contiguous()))


def test_auto_modes_keep_the_pytorch_layers_off_the_gpu_and_in_training():
    """encoder_mode = decoder_mode = "auto" (the default) selects the fused fp32-faithful launches only for calls that qualify;
    CPU tensors, training mode and autograd take the PyTorch layers (same values as the explicit "torch" mode)."""
    import torch
    import vqb200
    torch.manual_seed(3)
    model = vqb200.VQVAEPatch(hidden_dim=256, input_dim=2, num_embeddings=32, embedding_dim=16, n_resblocks=1,
                              learning_rate=1e-3, patch_size=25, batch_norm=False).eval()
    assert model.encoder_mode == "auto" and model.decoder_mode == "auto"
    x = torch.randn(3, 200, 2)
    with torch.no_grad():
        assert not model._fused_ok(x) and not model._fused_decoder_ok(torch.randn(3, 16, 16))
        z_auto = model.encode(x)
        model.encoder_mode = "torch"
        assert torch.equal(model.encode(x), z_auto)
        zq = torch.randn(3, 16, 16)
        model.decoder_mode = "auto"
        y_auto = model.decode(zq)
        model.decoder_mode = "torch"
        assert torch.equal(model.decode(zq), y_auto)
    # the operand cache of the split layers is plain torch code: built on CPU weights too, rebuilt when a weight changes
    w0 = model._split_weights()
    assert model._split_weights() is w0
    with torch.no_grad():
        model.encoder[1].shared_conv.weight.add_(1.0)
    w1 = model._split_weights()
    assert w1 is not w0 and w1["wp"].shape == (256, 512) and w1["layers"][0][0].shape == (256, 512)


def test_cycle_id_cache_encodes_every_distinct_cycle_once_per_dataset():
    """dedupe = "dataset" (SURVEY.md section 8(f) row 2): windows with a stride of one cycle spread over several batches;
    the cache hands every distinct cycle to the encoder once, and the ids equal the plain per-batch encode."""
    from vqb200.dataloader import latentspace_dataloader as L
    g = torch.Generator().manual_seed(11)
    stream = torch.randn(70, 200, 2, generator=g)
    stream[33] = stream[5]                                  # a repeat far apart (different batches)
    windows = torch.stack([stream[i:i + 20] for i in range(50)])        # 50 windows x 20 cycles
    seen = []
    def encode(c):
        seen.append(c.shape[0])
        return torch.stack([(c.sum(dim=(1, 2)) * 1000).long(), (c[:, 0, 0] * 1000).long(), (c[:, -1, 1] * 1000).long()], dim=1)
    for cap in (4 << 30, 70 * 1600 // 2, 0):                # representatives kept / dropped midway / fingerprints only
        cache = L.CycleIdCache("cpu", max_rep_bytes=cap)
        seen.clear()
        outs = [cache.encode(encode, windows[b:b + 10].reshape(-1, 200, 2)) for b in range(0, 50, 10)]
        want = encode(windows.reshape(-1, 200, 2))
        assert torch.equal(torch.cat(outs), want)
        assert sum(seen[:-1]) == 68 and len(cache) == 68    # 69 cycles touched, one pair identical
        assert cache.hits + cache.misses == 1000 and cache.misses == 68
    # a fingerprint that matches while the words differ is caught while the representatives are kept
    cache = L.CycleIdCache("cpu")
    a = torch.zeros(1, 200, 2); b = torch.ones(1, 200, 2)
    cache.insert(a, L.cycle_fingerprints(a), torch.tensor([[1, 2, 3]]))
    assert int(cache.lookup(b, L.cycle_fingerprints(a))[0]) == -1
    assert int(cache.lookup(a, L.cycle_fingerprints(a))[0]) == 0


def test_latent_dataset_pickle_is_the_reference_cache_format(tmp_path):
    """save_latent_dataset writes what BaseDataloader.save_one_pickle_file would (dataloader/base_dataloader.py:236-246):
    one pickle of (train, val, test), each an (x, y) pair, under quality_prediction_data/<dataset_name>/dataset.pickle."""
    import pickle
    from vqb200.dataloader import latentspace_dataloader as L
    name = L.latent_dataset_name("autoregressive_ids", "VQ-VAE-Patch", 20, "abc123")
    assert name == "autoregressive_ids_cycle_20_abc123"
    assert L.latent_dataset_name("classification_ids", "VQ-VAE-Patch", 1, "m") == "asimow_ls_classification_ids_VQ-VAE-Patch_cycle_1_m"
    with pytest.raises(ValueError):
        L.latent_dataset_name("nope", "m", 1, "x")
    splits = [(np.arange(12).reshape(3, 4), np.zeros(3)), (np.arange(4).reshape(1, 4), np.zeros(1)), (np.empty((0, 4), dtype=int), np.empty(0))]
    file = L.save_latent_dataset(splits, str(tmp_path), name)
    assert file == str(tmp_path / "quality_prediction_data" / name / "dataset.pickle")
    with open(file, "rb") as f:
        back = pickle.load(f)
    assert isinstance(back, tuple) and len(back) == 3
    for (x, y), (bx, by) in zip(splits, back):
        assert np.array_equal(x, bx) and np.array_equal(y, by) and bx.dtype == x.dtype


def test_async_host_writer_cpu_path_keeps_order_and_shapes():
    from vqb200.dataloader import latentspace_dataloader as L
    w = L._AsyncHostWriter("cpu")
    parts = [torch.arange(24).view(2, 3, 4) + 100 * i for i in range(4)]
    for t in parts:
        w.put(t)
    out = w.finish()
    assert out.shape == (8, 3, 4) and out.dtype == np.int64
    assert np.array_equal(out, torch.cat(parts).numpy())
    # a row-count hint sizes the array up front; a hint that was too small grows it; batches of changing size; dtype conversion
    for hint in (None, 1, 2, 4, 19, 40):
        w = L._AsyncHostWriter("cpu", rows_hint=hint, out_dtype=np.float64)
        ragged = [torch.full((r, 5), float(r)) for r in (3, 1, 7, 2, 6)]
        for t in ragged:
            w.put(t)
        out = w.finish()
        assert out.dtype == np.float64 and np.array_equal(out, torch.cat(ragged).numpy().astype(np.float64))
    assert L._AsyncHostWriter("cpu").finish() is None


def test_data_set_from_cycle_stream_windows_like_create_sequence_ds():
    """Host logic of create_latent_space_dataset_from_cycles on the CPU with a stand-in per-cycle encoder: window i = cycles
    i .. i + seq_len - 1, n - seq_len windows, label y[i + seq_len] (dataloader/asimow_dataloader.py:185-206), every
    cycle encoded once, and the bulk loops' CPU path (prefetcher and writer degenerate to plain copies)."""
    import types
    from vqb200.dataloader import LatentSpaceEncoder

    class Stub(torch.nn.Module):
        enc_out_len, embedding_dim = 3, 2
        vector_quantization = types.SimpleNamespace(code_counts=None)

    enc = LatentSpaceEncoder(Stub(), window_size=5, device="cpu", encoder_mode=None)
    seen = []

    def fake_ids(x, has_patch_embed=False):           # ids of a cycle = a function of its samples alone
        seen.append(x.shape[0])
        s = x[:, :5, :].sum(dim=(1, 2))
        return torch.stack([(s * (j + 1)).round().long() for j in range(3)], dim=1).view(-1, 1)

    enc.get_latent_space_IDs = fake_ids
    rs = np.random.RandomState(3)
    n, seq_len = 23, 4
    cycles = rs.randint(-9, 9, (n, 5, 2)).astype(np.float64)
    y = rs.randint(0, 2, n).astype(np.float64)
    ids, labels = enc.create_latent_space_dataset_from_cycles(cycles, y, seq_len=seq_len, has_patch_embed=True, batch=10)
    assert seen == [10, 10, 3]
    per_cycle = fake_ids(torch.from_numpy(cycles).float()).view(n, 3).numpy()
    assert ids.shape == (n - seq_len, seq_len, 3) and ids.dtype == np.int64 and ids.flags["C_CONTIGUOUS"]
    for i in range(n - seq_len):
        assert np.array_equal(ids[i], per_cycle[i:i + seq_len])
    assert np.array_equal(labels, y[seq_len:])
    # the same windows through the loop over MATERIALISED windows (what the reference feeds its loop with)
    win = np.stack([cycles[i:i + seq_len].reshape(seq_len * 5, 2) for i in range(n - seq_len)]).astype(np.float32)
    loop_ids, _ = enc.create_latent_space_dataset_VQ_VAE_IDs([torch.from_numpy(win[i:i + 6]) for i in range(0, len(win), 6)],
                                                             seq_len=seq_len, has_patch_embed=True, no_labels=True)
    assert np.array_equal(loop_ids, ids)
    view, _ = enc.create_latent_space_dataset_from_cycles(cycles, y, seq_len=seq_len, has_patch_embed=True, materialize=False)
    assert view.shape == ids.shape and np.array_equal(view, ids) and not view.flags["OWNDATA"] and not view.flags["WRITEABLE"]
    assert np.array_equal(view[7], ids[7]) and np.array_equal(view[[1, 5, 2]], ids[[1, 5, 2]])      # what a Dataset does with it
    flat, zeros = enc.create_latent_space_dataset_from_cycles(cycles, None, seq_len=seq_len, kind="ar_ids")
    assert np.array_equal(flat, ids.reshape(n - seq_len, -1)) and np.array_equal(zeros, np.zeros(n - seq_len))
    one, y1 = enc.create_latent_space_dataset_from_cycles(cycles, y, seq_len=1)
    assert np.array_equal(one[:, 0, :], per_cycle) and np.array_equal(y1, y)
    with pytest.raises(ValueError):
        enc.create_latent_space_dataset_from_cycles(cycles[:, :4, :], y, seq_len=2)


def test_prefetcher_groups_consecutive_loader_batches_in_order():
    """_DevicePrefetcher (CPU path): loader batches are concatenated into groups of at most group_cycles cycles -- never split,
    never reordered, a larger batch is a group of its own -- and the rows of a group are the cycles of its batches in order."""
    from vqb200.dataloader import latentspace_dataloader as L
    rs = np.random.RandomState(0)
    sizes = [3, 7, 1, 16, 5, 2, 2, 9]
    seq_len, window = 2, 4
    loader = [torch.from_numpy(rs.standard_normal((b, seq_len * window + 3, 2)).astype(np.float32)) for b in sizes]
    want = torch.cat([x[:, : seq_len * window, :].reshape(-1, window, 2) for x in loader])
    for group_cycles, expect in ((1, [[3], [7], [1], [16], [5], [2], [2], [9]]),
                                 (20, [[3, 7], [1], [16], [5, 2, 2], [9]]),
                                 (32, [[3, 7, 1], [16], [5, 2, 2], [9]]),
                                 (10 ** 6, [sizes])):
        got, groups = [], []
        for cyc, items in L._DevicePrefetcher(loader, "cpu", seq_len, window, no_labels=True, group_cycles=group_cycles):
            assert cyc.shape == (sum(int(x.shape[0]) for x in items) * seq_len, window, 2)
            got.append(cyc)
            groups.append([int(x.shape[0]) for x in items])
        assert groups == expect and torch.equal(torch.cat(got), want)
    labelled = [(x, torch.full((x.shape[0],), float(i))) for i, x in enumerate(loader)]
    ys = [float(y[0]) for _cyc, items in L._DevicePrefetcher(labelled, "cpu", seq_len, window, group_cycles=20) for _x, y in items]
    assert ys == [float(i) for i in range(len(sizes))]
    assert list(L._DevicePrefetcher([], "cpu", seq_len, window, no_labels=True)) == []
