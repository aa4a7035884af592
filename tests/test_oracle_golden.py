"""Pins the CPU oracle (oracle/vq_oracle.c and the torch port) against outputs of
the UNMODIFIED reference module (tests/golden/vq_golden.npz, produced by
oracle/make_golden.py from model/vector_quantizer.py:76-131).

Index rule: bit-exact, except rows where the reference's own fp32 GEMM rounding
decides between two codes whose fp64 distances are within a few ulp of the
distance magnitude (documented near-ties, SURVEY.md section 0 trap iii); every such
row must be explained, and their rate is bounded per case.
"""
import hashlib

import numpy as np
import pytest

import cases as C
from oracle import vq_oracle as O

NEAR_TIE_ULPS = 8.0
# documented near-tie budget: fraction of rows that may differ (all must be explained)
NEAR_TIE_RATE = {"stress_randn": 2e-3, "k1024_d64": 8e-3, "k8192_d32": 3e-2, "k2_d1": 5e-3}


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", C.CASES, ids=[c["name"] for c in C.CASES])
def test_c_oracle_matches_reference(case, golden, manifest):
    name = case["name"]
    _, z, E = C.make_inputs(case)
    res = O.forward(z, E, case["beta"])
    ref_idx = golden[f"{name}/idx"].astype(np.int64)
    n = ref_idx.shape[0]
    assert res.indices.shape == (n, 1) and res.indices.dtype == np.int64
    got = res.indices.reshape(-1)
    finite_case = case["z"] not in ("nonfinite",) and case["cb"] != "nan_rows"
    if finite_case:
        info = O.explain_mismatches(z, E, got, ref_idx, ulps=NEAR_TIE_ULPS)
        assert info["explained"] == info["mismatch"], info
        assert info["mismatch"] <= NEAR_TIE_RATE.get(name, 0.0) * n, info
    else:
        assert np.array_equal(got, ref_idx)
    same = got == ref_idx
    zq = res.z_q.reshape(n, case["D"])
    head = golden[f"{name}/zq_head"]
    rows = np.nonzero(same[: head.shape[0]])[0]
    assert np.array_equal(zq[: head.shape[0]][rows], head[rows], equal_nan=True)  # z + (e - z), bit-exact
    if same.all() and finite_case:  # NaN payload bits are not part of the contract
        assert _sha(zq) == manifest["vq"][name]["zq_sha256"]
    assert res.counts.sum() == n
    assert np.array_equal(res.counts, np.bincount(got, minlength=case["K"]))
    ref_loss, ref_ppl = golden[f"{name}/loss"], golden[f"{name}/perplexity"]
    if np.isnan(ref_loss):
        assert np.isnan(res.loss)
    else:
        assert res.loss == pytest.approx(ref_loss, rel=1e-5)
    if same.all():
        assert res.perplexity == pytest.approx(ref_ppl, rel=1e-5)


@pytest.mark.parametrize("case", [c for c in C.CASES if c["bwd"]], ids=[c["name"] for c in C.CASES if c["bwd"]])
def test_c_oracle_backward_matches_autograd(case, golden):
    name = case["name"]
    _, z, E = C.make_inputs(case)
    ref_idx = golden[f"{name}/idx"].astype(np.int64)
    w = C.upstream_weights(case)
    gz, gE = O.backward(w, C.G_LOSS, z, ref_idx, E, case["beta"])
    gz = gz.reshape(-1, case["D"])
    head = golden[f"{name}/grad_z_head"]
    np.testing.assert_allclose(gz[: head.shape[0]], head, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(gz.astype(np.float64).sum(0), golden[f"{name}/grad_z_colsum"], rtol=1e-4, atol=1e-4)
    ref_gE = golden[f"{name}/grad_E"]
    scale = np.abs(ref_gE).max() + 1e-30
    np.testing.assert_allclose(gE, ref_gE, rtol=1e-4, atol=1e-5 * scale)


@pytest.mark.parametrize("name", ["default_cfg", "permuted_view", "ties", "nonfinite_z", "k255_d33"])
def test_torch_port_matches_reference(name, golden, manifest):
    """The cpu_baseline port runs the reference's own op sequence; on the
    machine that generated the fixtures it reproduces them exactly, elsewhere
    (different BLAS kernels) only near-ties may move."""
    import torch
    torch.set_float32_matmul_precision("highest")
    case = C.case_by_name(name)
    storage, z, E = C.make_inputs(case)
    zt = torch.from_numpy(storage)
    zt = zt.permute(0, 2, 1) if case["layout"] == "permuted" else zt
    loss, zq, ppl, onehot, idx = O.torch_port_forward(zt, torch.from_numpy(E), case["beta"])
    ref_idx = golden[f"{name}/idx"].astype(np.int64)
    got = idx.numpy().reshape(-1)
    if case["z"] == "nonfinite":
        assert np.array_equal(got, ref_idx)
    else:
        info = O.explain_mismatches(z, E, got, ref_idx, ulps=NEAR_TIE_ULPS)
        assert info["explained"] == info["mismatch"], info
    assert onehot.shape == (ref_idx.shape[0], case["K"]) and idx.shape == (ref_idx.shape[0], 1)
    assert zq.shape == zt.shape and zq.is_contiguous()
    if np.array_equal(got, ref_idx):
        assert _sha(zq.numpy().reshape(-1, case["D"])) == manifest["vq"][name]["zq_sha256"]
    ref_loss = golden[f"{name}/loss"]
    if np.isnan(ref_loss):
        assert torch.isnan(loss)
    else:
        assert loss.item() == pytest.approx(ref_loss, rel=1e-6)
        assert ppl.item() == pytest.approx(golden[f"{name}/perplexity"], rel=1e-5)


def test_gather_restatement(golden):
    case = C.case_by_name("default_cfg")
    _, z, E = C.make_inputs(case)
    idx = golden["default_cfg/idx"].astype(np.int64)
    out = O.gather(idx.reshape(-1, 1), E, target_shape=case["shape"])
    assert out.shape == tuple(case["shape"])
    assert np.array_equal(out.reshape(-1, 32), E[idx])
    with pytest.raises(RuntimeError):
        O.gather(np.array([[256]]), E)


def test_encoder_port_reproduces_the_reference_encoder(patch_golden):
    """oracle.torch_port_encode (the CPU leg of bench.py's bulk-encode object) against z_e of the unmodified reference:
    bit-equal, same permuted-view strides (model/vq_vae_patch_embedd.py:91)."""
    import torch
    import cases as C
    from oracle import vq_oracle as O
    for case in C.PATCH_CASES:
        if case["batch_norm"]:
            continue
        pre = f"{case['name']}/sd/"
        sd = {k[len(pre):]: torch.from_numpy(patch_golden[k]) for k in patch_golden.files if k.startswith(pre)}
        x = torch.from_numpy(C.make_cycles(case))
        with torch.no_grad():
            z = O.torch_port_encode(sd, x, case["patch_size"])
        assert not z.is_contiguous() and z.stride(1) == 1
        assert np.array_equal(z.contiguous().numpy(), patch_golden[f"{case['name']}/z_e"])
        ids = O.torch_port_forward(z, sd["vector_quantization.embedding.weight"], case["beta"])[4]
        assert np.array_equal(ids.numpy().reshape(-1), patch_golden[f"{case['name']}/idx"])


def test_port_ids_of_the_cycle_stream_window_into_the_reference_data_set(patch_golden, bulk_golden):
    """The oracle port on the CYCLES of the bulk fixture's stream, windowed the way the reference's create_sequence_ds does
    (dataloader/asimow_dataloader.py:185-206: n - seq_len windows, window i = cycles i .. i + seq_len - 1, label
    y[i + seq_len]) == what the unmodified reference built from its own windows (bulk_overlap/seq_*): the property
    create_latent_space_dataset_from_cycles rests on (every op before the decoder is per cycle), pinned on the CPU."""
    import torch
    import cases as C
    from oracle import vq_oracle as O
    case = C.BULK_CASE
    mcase = next(c for c in C.PATCH_CASES if c["name"] == case["model"])
    if mcase["batch_norm"]:
        pytest.skip("the encoder port covers the models without BatchNorm")
    pre = f"{mcase['name']}/sd/"
    sd = {k[len(pre):]: torch.from_numpy(patch_golden[k]) for k in patch_golden.files if k.startswith(pre)}
    stream, cycle_labels = C.make_stream(case)
    with torch.no_grad():
        z = O.torch_port_encode(sd, torch.from_numpy(stream), mcase["patch_size"])
        per_cycle = O.torch_port_forward(z, sd["vector_quantization.embedding.weight"], mcase["beta"])[4].numpy().reshape(len(stream), -1)
    seq_len = case["seq_len"]
    want = bulk_golden[f"{case['name']}/seq_ids"]
    n_windows = len(stream) - seq_len
    assert want.shape == (n_windows, seq_len, per_cycle.shape[1])
    got = np.stack([per_cycle[i:i + seq_len] for i in range(n_windows)])
    assert np.array_equal(got, want)
    assert np.array_equal(bulk_golden[f"{case['name']}/seq_labels"], cycle_labels[seq_len:])
    # and the older fixture (the repo's own n - seq_len + 1 windows through the reference's loop) is the same view, one longer
    assert np.array_equal(bulk_golden[f"{case['name']}/ids"][:n_windows], want)
