"""GPU parity tests proper (run with -m gpu on the B200 box): the CUDA path, called through
the C ABI, against (1) the CPU oracle on the same seeded inputs -- indices and z_q bit-exact --
and (2) the committed outputs of the unmodified reference (tests/golden/), where only
documented fp32 near-ties may move.

Tolerances (BASELINE.json north_star): indices bit-exact vs the oracle; z_q bit-exact vs the
oracle; loss, perplexity, grad_z, grad_E within 1e-5 relative (fp32).
"""
import hashlib

import numpy as np
import pytest
import torch

import cases as C
import vqb200
from vqb200 import ops
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu

REL = 1e-5
PATHS = ["fma", "auto"]


def _dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    return torch.device("cuda:0")


def _to_dev(case):
    storage, logical, E = C.make_inputs(case)
    st = torch.from_numpy(storage).to(_dev())
    z = st.permute(0, 2, 1) if case["layout"] == "permuted" else st
    assert tuple(z.shape) == tuple(case["shape"])
    return z, logical, E


@pytest.fixture(autouse=True, params=["auto", "bf16", "tf32"])
def tc_filter(request):
    """Every test of this module runs with both filters of the tcgen05 forward kernel (three bf16 products / one tf32
    product, csrc/vq_fwd_tc.cu) forced, and with the automatic choice: the decision must be the oracle's with either."""
    ops.set_filter(None if request.param == "auto" else request.param)
    yield request.param
    ops.set_filter(None)


def _module(case, path, one_hot="dense"):
    vq = vqb200.VectorQuantizer(case["K"], case["D"], case["beta"], one_hot=one_hot, path=path).to(_dev())
    return vq


def _finite(case):
    return case["z"] != "nonfinite" and case["cb"] != "nan_rows"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("case", C.CASES, ids=[c["name"] for c in C.CASES])
def test_forward_matches_oracle_and_reference(case, path, golden, manifest):
    name = case["name"]
    z, z_np, E = _to_dev(case)
    vq = _module(case, path)
    with torch.no_grad():
        vq.embedding.weight.copy_(torch.from_numpy(E))
        loss, zq, ppl, onehot, idx = vq(z)
    torch.cuda.synchronize()
    n = z.numel() // case["D"]
    # ---- the reference's return contract (model/vector_quantizer.py:118-119) ----
    assert loss.shape == () and ppl.shape == () and loss.dtype == torch.float32
    assert zq.shape == z.shape and zq.is_contiguous() and zq.dtype == torch.float32
    assert idx.shape == (n, 1) and idx.dtype == torch.int64
    assert onehot.shape == (n, case["K"]) and onehot.dtype == torch.float32
    got = idx.cpu().numpy().reshape(-1)
    # ---- vs the oracle: bit-exact ----
    ora = O.forward(z_np, E, case["beta"])
    assert np.array_equal(got, ora.indices.reshape(-1)), "indices differ from the oracle"
    zq_np = zq.cpu().numpy().reshape(n, case["D"])
    assert np.array_equal(zq_np, ora.z_q.reshape(n, case["D"]), equal_nan=True), "z_q differs from the oracle"
    assert np.array_equal(vq.code_counts.cpu().numpy(), ora.counts)
    oh = onehot.cpu().numpy()
    assert np.array_equal(oh.argmax(1), got) and np.array_equal(oh.sum(1), np.ones(n, np.float32))
    if np.isnan(ora.loss):
        assert torch.isnan(loss)
    else:
        assert loss.item() == pytest.approx(float(ora.loss), rel=REL)
    assert ppl.item() == pytest.approx(float(ora.perplexity), rel=REL)
    # ---- vs the reference's stored outputs ----
    ref_idx = golden[f"{name}/idx"].astype(np.int64)
    if _finite(case):
        info = O.explain_mismatches(z_np, E, got, ref_idx, ulps=8.0)
        assert info["explained"] == info["mismatch"], info
        assert info["mismatch"] <= max(1, int(3e-2 * n)) if info["mismatch"] else True
    else:
        assert np.array_equal(got, ref_idx)
    if np.array_equal(got, ref_idx):
        if _finite(case):
            assert hashlib.sha256(zq_np.tobytes()).hexdigest() == manifest["vq"][name]["zq_sha256"]
        ref_loss = golden[f"{name}/loss"]
        if not np.isnan(ref_loss):
            assert loss.item() == pytest.approx(float(ref_loss), rel=REL)
            assert ppl.item() == pytest.approx(float(golden[f"{name}/perplexity"]), rel=REL)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("case", [c for c in C.CASES if c["bwd"]], ids=[c["name"] for c in C.CASES if c["bwd"]])
def test_backward_matches_reference_autograd(case, path, golden):
    name = case["name"]
    z, z_np, E = _to_dev(case)
    vq = _module(case, path, one_hot="none")
    with torch.no_grad():
        vq.embedding.weight.copy_(torch.from_numpy(E))
    z = z.detach().requires_grad_(True)
    loss, zq, ppl, onehot, idx = vq(z)
    assert onehot is None
    assert loss.requires_grad and zq.requires_grad and not ppl.requires_grad and not idx.requires_grad
    w = torch.from_numpy(C.upstream_weights(case)).to(_dev())
    total = C.G_LOSS * loss + (w * zq).sum()
    total.backward()
    torch.cuda.synchronize()
    n = z.numel() // case["D"]
    assert z.grad.shape == z.shape
    gz = z.grad.cpu().numpy().reshape(n, case["D"])
    gE = vq.embedding.weight.grad.cpu().numpy()
    got_idx = idx.cpu().numpy().reshape(-1)
    ref_idx = golden[f"{name}/idx"].astype(np.int64)
    # closed form in fp64 on the indices the kernel chose
    ogz, ogE = O.backward(C.upstream_weights(case), C.G_LOSS, z_np, got_idx, E, case["beta"])
    np.testing.assert_allclose(gz, ogz.reshape(n, case["D"]), rtol=REL, atol=1e-7)
    scale = np.abs(ogE).max() + 1e-30
    np.testing.assert_allclose(gE, ogE, rtol=1e-4, atol=REL * scale)
    if np.array_equal(got_idx, ref_idx):   # and against the reference's autograd
        head = golden[f"{name}/grad_z_head"]
        np.testing.assert_allclose(gz[: head.shape[0]], head, rtol=REL, atol=1e-7)
        rscale = np.abs(golden[f"{name}/grad_E"]).max() + 1e-30
        np.testing.assert_allclose(gE, golden[f"{name}/grad_E"], rtol=1e-4, atol=REL * rscale)


def test_gradient_routing_matches_reference_semantics():
    """g_zq reaches z only (straight-through); the codebook only sees the beta term."""
    dev = _dev()
    torch.manual_seed(3)
    vq = vqb200.VectorQuantizer(32, 8, 0.25, one_hot="none").to(dev)
    z = (0.05 * torch.randn(64, 8, device=dev)).requires_grad_(True)
    loss, zq, *_ = vq(z)
    zq.sum().backward()
    assert torch.equal(z.grad, torch.ones_like(z))
    assert torch.count_nonzero(vq.embedding.weight.grad) == 0
    z.grad = None
    vq.zero_grad()
    loss, zq, *_ = vq(z)
    loss.backward()
    assert torch.count_nonzero(vq.embedding.weight.grad) > 0
    with torch.no_grad():
        out = vq(z)
    assert not out[0].requires_grad and not out[1].requires_grad


def test_gather_and_bad_index():
    dev = _dev()
    vq = vqb200.VectorQuantizer(16, 8, 0.25).to(dev)
    idx = torch.tensor([[3], [0], [15], [3]], device=dev)
    out = vq.get_embedding_from_one_hot(idx, (2, 2, 8))
    assert out.shape == (2, 2, 8) and out.is_contiguous()
    assert torch.equal(out.view(4, 8), vq.embedding.weight.detach()[idx.view(-1)])
    bad = ops.gather(torch.tensor([[16]], device=dev), vq.embedding.weight)
    assert torch.isnan(bad).all()


@pytest.mark.parametrize("path", PATHS)
def test_encode_indices_and_empty_input(path):
    dev = _dev()
    vq = vqb200.VectorQuantizer(256, 32, 0.25, path=path).to(dev)
    z = 0.1 * torch.randn(1000, 32, device=dev)
    with torch.no_grad():
        full = vq(z)[4]
    ids = vq.encode_indices(z)
    assert torch.equal(ids, full)
    loss, zq, ppl, oh, idx = vq(torch.empty(0, 32, device=dev))
    assert idx.shape == (0, 1) and zq.shape == (0, 32) and oh.shape == (0, 256)
    assert torch.isnan(loss)


@pytest.mark.parametrize("path", PATHS)
def test_random_shapes_against_oracle(path):
    """Property test: random (K, D, N, layout) against the oracle, bit-exact."""
    dev = _dev()
    rs = np.random.RandomState(2024)
    for trial in range(24):
        K = int(rs.choice([1, 2, 3, 17, 64, 100, 256, 300, 1000]))
        D = int(rs.choice([1, 2, 4, 7, 8, 16, 24, 32, 48, 64, 128]))
        n = int(rs.choice([1, 31, 128, 129, 257, 1000, 4096]))
        scale = float(rs.choice([0.1, 1.0]))
        E = (rs.standard_normal((K, D)) * float(rs.choice([1.0 / K, 0.1]))).astype(np.float32)
        z_np = (scale * rs.standard_normal((n, D))).astype(np.float32)
        layout = rs.choice(["contig", "colmajor", "permuted"])
        zt = torch.from_numpy(z_np).to(dev)
        if layout == "colmajor":
            zt = zt.t().contiguous().t()
        elif layout == "permuted" and n % 4 == 0:
            zt = zt.view(n // 4, 4, D).permute(0, 2, 1).contiguous().permute(0, 2, 1)
        vq = vqb200.VectorQuantizer(K, D, 0.25, one_hot="none", path=path).to(dev)
        with torch.no_grad():
            vq.embedding.weight.copy_(torch.from_numpy(E))
            loss, zq, ppl, _, idx = vq(zt)
        ora = O.forward(z_np, E, 0.25)
        tag = f"trial {trial}: K={K} D={D} n={n} {layout}"
        assert np.array_equal(idx.cpu().numpy().reshape(-1), ora.indices.reshape(-1)), tag
        assert np.array_equal(zq.cpu().numpy().reshape(n, D), ora.z_q.reshape(n, D)), tag
        assert loss.item() == pytest.approx(float(ora.loss), rel=REL), tag
        assert ppl.item() == pytest.approx(float(ora.perplexity), rel=REL), tag


@pytest.mark.parametrize("path", PATHS)
def test_full_size_properties(path):
    """BASELINE config 2 at full size (2^24 vectors, K=256, D=32): size-independent properties
    plus an oracle check on a 2^16-row sample."""
    dev = _dev()
    n, K, D = 1 << 24, 256, 32
    g = torch.Generator(device=dev).manual_seed(1234)
    z = 0.1 * torch.randn(n, D, device=dev, generator=g)
    torch.manual_seed(0)
    vq = vqb200.VectorQuantizer(K, D, 0.25, one_hot="none", path=path).to(dev)
    with torch.no_grad():
        loss, zq, ppl, _, idx = vq(z)
    torch.cuda.synchronize()
    E = vq.embedding.weight.detach()
    flat = idx.view(-1)
    assert int(flat.min()) >= 0 and int(flat.max()) < K
    counts = vq.code_counts
    assert int(counts.sum()) == n
    assert torch.equal(counts, torch.bincount(flat, minlength=K))
    # z_q is exactly z + (E[idx] - z); the loss is the mean of the squared residual
    e = E[flat]
    assert torch.equal(zq, z + (e - z))
    m = ((e - z).double() ** 2).mean()
    assert loss.item() == pytest.approx(float(m * 1.25), rel=REL)
    p = counts.double() / n
    assert ppl.item() == pytest.approx(float(torch.exp(-(p * torch.log(p + 1e-10)).sum())), rel=REL)
    # optimality: no code is closer than the chosen one (fp64 check on a strided sample)
    sample = torch.arange(0, n, n // 4096, device=dev)[:4096]
    d64 = torch.cdist(z[sample].double(), E.double())
    chosen = d64.gather(1, flat[sample].view(-1, 1)).view(-1)
    assert torch.all(chosen <= d64.min(1).values + 1e-6)
    # oracle on the first 2^16 rows: bit-exact
    head = 1 << 16
    ora = O.forward(z[:head].cpu().numpy(), E.cpu().numpy(), 0.25)
    assert np.array_equal(flat[:head].cpu().numpy(), ora.indices.reshape(-1))
    # idempotence: quantising the quantised vectors' code rows returns codes with identical rows
    with torch.no_grad():
        idx2 = vq.encode_indices(e[: 1 << 20].contiguous()).view(-1)
    assert torch.equal(E[idx2], e[: 1 << 20])


@pytest.mark.parametrize("K,D,n", [(256, 32, 300_000), (600, 32, 150_000), (256, 64, 150_000), (100, 16, 150_000)])
def test_host_buffer_path_matches_oracle(K, D, n):
    """vqb_encode_host (pinned host buffers, pipelined copies) on the plain, chunked-K, wide-D and narrow-D shapes."""
    rs = np.random.RandomState(7)
    E = rs.uniform(-1.0 / K, 1.0 / K, (K, D)).astype(np.float32)
    z = torch.from_numpy((0.1 * rs.standard_normal((n, D))).astype(np.float32)).pin_memory()
    zq = torch.empty_like(z).pin_memory()
    idx = torch.empty(n, dtype=torch.int64).pin_memory()
    counts = np.zeros(K, np.uint64)
    enc = ops.HostEncoder(E, device=0, chunk_rows=65536, depth=3)
    loss, ppl = enc.encode(z, 0.25, zq_out=zq, idx_out=idx, counts_out=counts)
    assert enc.last_launches >= 2 + (n + 65535) // 65536
    ora = O.forward(z.numpy(), E, 0.25)
    assert np.array_equal(idx.numpy(), ora.indices.reshape(-1))
    assert np.array_equal(zq.numpy(), ora.z_q)
    assert np.array_equal(counts.astype(np.int64), ora.counts)
    assert loss == pytest.approx(float(ora.loss), rel=REL)
    assert ppl == pytest.approx(float(ora.perplexity), rel=REL)
    # ids-only mode (what the latent-dataset builder consumes); the second call re-uses the prepared codebook
    idx2 = torch.empty(n, dtype=torch.int64).pin_memory()
    enc.encode(z, 0.25, idx_out=idx2)
    assert torch.equal(idx, idx2)
    # compact ids: the same ids in 1 (K <= 256) or 2 bytes each
    small = torch.empty(n, dtype=torch.uint8 if K <= 256 else torch.uint16).pin_memory()
    enc.encode(z, 0.25, idx_out=small)
    assert np.array_equal(small.numpy().astype(np.int64), idx.numpy())
    if K > 256:
        with pytest.raises(RuntimeError):
            enc.encode(z, 0.25, idx_out=torch.empty(n, dtype=torch.uint8))
    enc.close()


def test_host_buffer_path_validates_its_arguments():
    """Raw host pointers cross the C ABI: wrong dtypes, strided views, short buffers and device tensors raise instead of
    being misread or overrun (round-1 advisor finding)."""
    rs = np.random.RandomState(3)
    E = rs.uniform(-0.01, 0.01, (64, 32)).astype(np.float32)
    enc = ops.HostEncoder(E, device=0, chunk_rows=4096, depth=2)
    z = (0.1 * rs.standard_normal((1000, 32))).astype(np.float32)
    idx = np.empty(1000, np.int64)
    enc.encode(z, idx_out=idx)                                        # baseline: fine
    enc.encode(torch.from_numpy(z).reshape(10, 100, 32), idx_out=idx)  # any shape with numel % d == 0
    for bad_z in (z.astype(np.float64), z[:, ::2], z.reshape(-1)[:31999], torch.from_numpy(z).cuda(), list(z[:2])):
        with pytest.raises((RuntimeError, TypeError)):
            enc.encode(bad_z, idx_out=idx)
    for kw in (dict(idx_out=np.empty(1000, np.int32)), dict(idx_out=np.empty(999, np.int64)),
               dict(zq_out=np.empty((1000, 16), np.float32)), dict(counts_out=np.zeros(63, np.uint64)),
               dict(counts_out=np.zeros(64, np.float64)), dict(idx_out=np.empty(2000, np.int64)[::2])):
        with pytest.raises(RuntimeError):
            enc.encode(z, **kw)
    enc.close()


def test_codebook_cache_follows_the_weight_version():
    """The binding re-uses the prepared codebook (norms, census, tcgen05 image) while (data_ptr, _version) of the weight
    stay the same, and re-prepares after an in-place update; a write through .data needs clear_workspaces()."""
    dev = _dev()
    rs = np.random.RandomState(11)
    z_np = (0.1 * rs.standard_normal((5000, 32))).astype(np.float32)
    E = rs.uniform(-1 / 256, 1 / 256, (256, 32)).astype(np.float32)
    z = torch.from_numpy(z_np).to(dev)
    w = torch.nn.Parameter(torch.from_numpy(E).to(dev))
    for path in ("tc", "fma", "auto"):
        a = ops.forward(z, w, 0.25, path=path)
        b = ops.forward(z, w, 0.25, path=path)                        # cached
        ora = O.forward(z_np, E, 0.25)
        for out in (a, b):
            assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora.indices.reshape(-1))
            assert np.array_equal(out[1].cpu().numpy(), ora.z_q)
            assert out[0].item() == pytest.approx(float(ora.loss), rel=REL)
    with torch.no_grad():
        w.mul_(-3.0)                                                  # in place: bumps _version
    E2 = w.detach().cpu().numpy()
    ora2 = O.forward(z_np, E2, 0.25)
    for path in ("auto", "fma"):
        out = ops.forward(z, w, 0.25, path=path)
        assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora2.indices.reshape(-1))
    w.data.add_(0.001)                                                # invisible to _version
    ops.clear_workspaces()
    ora3 = O.forward(z_np, w.detach().cpu().numpy(), 0.25)
    out = ops.forward(z, w, 0.25)
    assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora3.indices.reshape(-1))
    # a small call (FMA path) followed by a large one (tcgen05 path) on the same workspace: the image is built then
    ops.clear_workspaces()
    ops.forward(z[:64], w, 0.25)
    out = ops.forward(z, w, 0.25)
    assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora3.indices.reshape(-1))


# ---------------------------------------------------------------------------------------
# tcgen05 path specifics (D = 32, K <= 256, contiguous rows): the cases the filter must hand
# to the exact scan -- exact ties, near-ties, non-finite vectors and codebooks, pad codes
# ---------------------------------------------------------------------------------------
def _tc_vs_oracle(z_np, E, beta=0.25, path="tc"):
    dev = _dev()
    z = torch.from_numpy(z_np).to(dev)
    out = ops.forward(z, torch.from_numpy(E).to(dev), beta, path=path, want_stats=True)
    torch.cuda.synchronize()
    ora = O.forward(z_np, E, beta)
    n, d = z_np.shape
    assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora.indices.reshape(-1))
    assert np.array_equal(out[1].cpu().numpy().reshape(n, d), ora.z_q.reshape(n, d), equal_nan=True)
    assert np.array_equal(out[4].cpu().numpy(), ora.counts)
    if np.isnan(ora.loss):
        assert torch.isnan(out[0])
    else:
        assert out[0].item() == pytest.approx(float(ora.loss), rel=REL)
    assert out[2].item() == pytest.approx(float(ora.perplexity), rel=REL)
    return out[5].cpu().numpy()


def test_tc_exact_ties_take_lowest_index():
    rs = np.random.RandomState(31)
    K, D, n = 256, 32, 1000
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    E[200] = E[7]; E[201] = E[7]; E[16] = E[3]; E[255] = E[0]      # duplicates across A- and B-groups
    pick = rs.randint(0, K, n)
    z = E[pick].copy()
    z[n // 2:] += (0.01 * rs.standard_normal((n - n // 2, D))).astype(np.float32)
    z[:8] = (E[1] + E[2]) / 2                                       # equidistant from two codes
    stats = _tc_vs_oracle(z, E)
    assert stats[1] > 0                                             # ties went through the exact scan


def test_tc_nonfinite_vectors_and_codebook():
    rs = np.random.RandomState(32)
    K, D, n = 256, 32, 640
    E = rs.uniform(-1.0 / K, 1.0 / K, (K, D)).astype(np.float32)
    z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
    z[1, 2] = np.nan; z[2, 0] = np.inf; z[3, 1] = -np.inf; z[4, :] = np.nan
    z[5, 0] = np.inf; z[5, 1] = -np.inf; z[6, :] = 0.0; z[7, 3] = 3.0e38; z[130, 31] = 1.0e30; z[639, 0] = np.nan
    _tc_vs_oracle(z, E)
    E2 = E.copy(); E2[5, 1] = np.nan; E2[2, 3] = np.inf; E2[77, 30] = -np.inf
    _tc_vs_oracle((0.1 * rs.standard_normal((n, D))).astype(np.float32), E2)
    E3 = E.copy(); E3[9] *= 1.0e20                                  # finite entries, ee overflows to inf
    _tc_vs_oracle((0.1 * rs.standard_normal((n, D))).astype(np.float32), E3)


@pytest.mark.parametrize("K", [1, 2, 15, 16, 17, 31, 33, 100, 255, 256])
@pytest.mark.parametrize("n", [128, 129, 1000])
def test_tc_pad_codes_and_ragged_tiles(K, n):
    rs = np.random.RandomState(K * 1000 + n)
    E = (0.1 * rs.standard_normal((K, 32))).astype(np.float32)
    z = (0.1 * rs.standard_normal((n, 32))).astype(np.float32)
    _tc_vs_oracle(z, E)


def test_tc_scale_extremes():
    """Tiny and huge magnitudes, and a codebook much larger than the vectors (the filter radius is
    dominated by ee there: many vectors take the exact scan, results must not change)."""
    rs = np.random.RandomState(33)
    for zs, es in ((1e-20, 1e-20), (1e15, 1e15), (0.01, 10.0), (10.0, 0.001), (1e-3, 1.0)):
        E = (es * rs.standard_normal((256, 32))).astype(np.float32)
        z = (zs * rs.standard_normal((2000, 32))).astype(np.float32)
        _tc_vs_oracle(z, E)


def test_tc_matches_fma_on_4m_vectors():
    """Bit-exact agreement of the two kernel paths on 2^22 vectors, trained-like and default codebooks."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(99)
    for scale_e in (None, 0.1):
        z = 0.1 * torch.randn(1 << 22, 32, device=dev, generator=g)
        w = (torch.rand(256, 32, device=dev, generator=g) * 2 - 1) / 256 if scale_e is None else \
            scale_e * torch.randn(256, 32, device=dev, generator=g)
        a = ops.forward(z, w, 0.25, path="fma")
        b = ops.forward(z, w, 0.25, path="tc")
        assert torch.equal(a[3], b[3]) and torch.equal(a[1], b[1]) and torch.equal(a[4], b[4])
        assert b[0].item() == pytest.approx(a[0].item(), rel=1e-6)


def test_tc_crowded_candidates_and_queue_overflow():
    """The fix-up queue takes vectors with at most 32 candidate codes and 2048 vectors per CTA; everything
    beyond that is decided by the warp-wide exact scan inside the main kernel.  Both routes must agree with
    the oracle: (a) a codebook of 256 identical rows (every code is a candidate, lowest index wins),
    (b) a codebook of 8 distinct rows repeated 32 times (cross product of several A- and B-groups),
    (c) 2^19 vectors that are ALL uncertain (each one the midpoint of two codes), which overflows every queue."""
    rs = np.random.RandomState(34)
    K, D = 256, 32
    z = (0.1 * rs.standard_normal((1500, D))).astype(np.float32)
    E = np.tile((0.1 * rs.standard_normal((1, D))).astype(np.float32), (K, 1))
    stats = _tc_vs_oracle(z, E)
    assert stats[1] == 1500
    E = np.tile((0.1 * rs.standard_normal((8, D))).astype(np.float32), (32, 1))
    stats = _tc_vs_oracle(z, E)
    assert stats[1] == 1500
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    a, b = rs.randint(0, K, 1 << 19), rs.randint(0, K, 1 << 19)
    z = ((E[a] + E[b]) * 0.5).astype(np.float32)
    stats = _tc_vs_oracle(z, E)
    assert stats[1] > (1 << 18)


def test_tc_subnormal_inputs():
    """Sub-normal components are lost by the bf16 split (and may be flushed by the tensor core): the
    absolute floor of the filter radius must send such near-ties to the exact path."""
    rs = np.random.RandomState(35)
    E = (rs.standard_normal((256, 32)) * 1e-38).astype(np.float32)
    z = (rs.standard_normal((2000, 32)) * 1e-38).astype(np.float32)
    _tc_vs_oracle(z, E)
    E = (rs.standard_normal((256, 32))).astype(np.float32)
    z = (rs.standard_normal((2000, 32)) * 1e-39).astype(np.float32)
    z[:, 0] = 0.0
    _tc_vs_oracle(z, E)


@pytest.mark.parametrize("path", PATHS)
def test_ids_only_mode_skips_nothing_that_matters(path):
    """want_zq=False, want_loss=False (encode_indices): ids and histogram equal those of the full call."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(5)
    z = 0.1 * torch.randn(70000, 32, device=dev, generator=g)
    w = (torch.rand(256, 32, device=dev, generator=g) * 2 - 1) / 256
    full = ops.forward(z, w, 0.25, path=path)
    ids = ops.forward(z, w, 0.25, path=path, want_zq=False, want_loss=False)
    assert ids[0] is None and ids[1] is None
    assert torch.equal(full[3], ids[3]) and torch.equal(full[4], ids[4])
    assert ids[2].item() == pytest.approx(full[2].item(), rel=1e-6)


@pytest.mark.parametrize("D", [4, 8, 12, 16, 20, 28])
@pytest.mark.parametrize("K", [5, 64, 256])
def test_tc_narrow_vectors(D, K):
    """D < 32 (multiples of 4) runs on the tcgen05 path as D = 32 with zero columns: the TMA boxes
    zero-fill on load and clip on store; results stay bit-identical to the oracle."""
    rs = np.random.RandomState(D * 1000 + K)
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    if K > 2:
        E[K - 1] = E[0]                                   # an exact tie across the codebook
    for n in (128, 777):
        z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
        z[0] = E[0]
        dev = _dev()
        assert ops._tc_eligible(n, K, D)
        out = ops.forward(torch.from_numpy(z).to(dev), torch.from_numpy(E).to(dev), 0.25, path="tc", want_stats=True)
        ora = O.forward(z, E, 0.25)
        assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora.indices.reshape(-1))
        assert np.array_equal(out[1].cpu().numpy(), ora.z_q.reshape(n, D))
        assert np.array_equal(out[4].cpu().numpy(), ora.counts)
        assert out[0].item() == pytest.approx(float(ora.loss), rel=REL)


@pytest.mark.parametrize("K,D", [(257, 32), (300, 32), (512, 32), (1000, 16), (1024, 32), (2048, 8), (4100, 32)])
def test_tc_chunked_large_codebooks(K, D):
    """K > 256 on the tcgen05 path: one pass per 256-code chunk, the chunk winners compared by their
    oracle-order distances (first NaN wins, ties keep the lower index), then the finish kernel.  Covers a
    ragged last chunk, duplicated rows across chunks, NaN rows and a non-finite code in a late chunk."""
    rs = np.random.RandomState(K + D)
    n = 3000
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    E[K - 1] = E[3]                                       # tie between the first and the last chunk
    E[260 % K] = E[10]
    z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
    z[0] = E[3]; z[1] = E[10]; z[2] = (E[5] + E[K - 2]) / 2
    z[7, 1] = np.nan; z[8, 0] = np.inf
    dev = _dev()
    assert ops._tc_eligible(n, K, D)
    for Eb in (E, None):
        if Eb is None:
            Eb = E.copy(); Eb[K - 5, D - 1] = np.nan; Eb[2, 0] = np.inf     # poisoned columns
        out = ops.forward(torch.from_numpy(z).to(dev), torch.from_numpy(Eb).to(dev), 0.25, path="tc")
        ora = O.forward(z, Eb, 0.25)
        assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora.indices.reshape(-1))
        assert np.array_equal(out[1].cpu().numpy(), ora.z_q.reshape(n, D), equal_nan=True)
        assert np.array_equal(out[4].cpu().numpy(), ora.counts)
        if np.isnan(ora.loss):
            assert torch.isnan(out[0])
        else:
            assert out[0].item() == pytest.approx(float(ora.loss), rel=REL)


@pytest.mark.parametrize("D", [36, 48, 64])
@pytest.mark.parametrize("K", [7, 256, 700])
def test_tc_wide_vectors(D, K):
    """32 < D <= 64 on the tcgen05 path: two 32-component D-chunks per tile accumulated in one TMEM buffer,
    fp32 codebook rows from global memory; also through the chunked large-K passes (K = 700)."""
    rs = np.random.RandomState(D * 7 + K)
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    E[K - 1] = E[0]
    dev = _dev()
    for n in (128, 1500):
        z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
        z[0] = E[0]; z[1] = (E[1] + E[2]) / 2; z[5, D - 1] = np.nan; z[6, 33] = np.inf
        assert ops._tc_eligible(n, K, D)
        for Eb in (E, None):
            if Eb is None:
                Eb = E.copy(); Eb[min(3, K - 1), D - 2] = np.nan          # poisoned column in the second chunk
            out = ops.forward(torch.from_numpy(z).to(dev), torch.from_numpy(Eb).to(dev), 0.25, path="tc", want_stats=True)
            ora = O.forward(z, Eb, 0.25)
            assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora.indices.reshape(-1))
            assert np.array_equal(out[1].cpu().numpy(), ora.z_q.reshape(n, D), equal_nan=True)
            assert np.array_equal(out[4].cpu().numpy(), ora.counts)
            if np.isnan(ora.loss):
                assert torch.isnan(out[0])
            else:
                assert out[0].item() == pytest.approx(float(ora.loss), rel=REL)


def _check_against_oracle(z, E, out, n, D):
    ora = O.forward(z, E, 0.25)
    assert np.array_equal(out[3].cpu().numpy().reshape(-1), ora.indices.reshape(-1))
    assert np.array_equal(out[1].cpu().numpy(), ora.z_q.reshape(n, D), equal_nan=True)
    assert np.array_equal(out[4].cpu().numpy(), ora.counts)
    if np.isnan(ora.loss):
        assert torch.isnan(out[0])
    else:
        assert out[0].item() == pytest.approx(float(ora.loss), rel=REL)
    assert out[2].item() == pytest.approx(float(ora.perplexity), rel=REL)


@pytest.mark.parametrize("K,D", [(7, 68), (256, 96), (256, 128), (300, 100), (700, 128), (1024, 64), (2100, 32), (4096, 128)])
def test_tcs_wide_vectors_and_large_codebooks(K, D):
    """The tile-stationary kernel (csrc/vq_fwd_tcs.cu): 64 < D <= 128 (three / four D-chunks per tile, operand blocks
    streamed through the ring) and K > 256 with D > 32, histogram in shared memory (K <= 2048) and in global memory;
    ties across chunks and column halves, NaN / inf rows, poisoned codebook columns."""
    rs = np.random.RandomState(K * 3 + D)
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    E[K - 1] = E[0]                                           # exact tie, first against last code
    if K > 200:
        E[130] = E[5]                                         # ... and across the two column halves of a chunk
    dev = _dev()
    for n in (128, 1500, 128 * 148 + 300):
        z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
        z[0] = E[0]; z[1] = (E[1] + E[2]) / 2; z[2] = E[5 % K]; z[5, D - 1] = np.nan; z[6, D // 2 + 1] = np.inf
        assert ops._tc_eligible(n, K, D)
        for Eb in (E, None):
            if Eb is None:
                if n != 1500:
                    continue
                Eb = E.copy(); Eb[min(3, K - 1), D - 2] = np.nan; Eb[K // 2, 0] = np.inf
            out = ops.forward(torch.from_numpy(z).to(dev), torch.from_numpy(Eb).to(dev), 0.25, path="tc")
            _check_against_oracle(z, Eb, out, n, D)


@pytest.mark.parametrize("K,D,tiles_per_cta", [(512, 32, 5), (1024, 32, 4), (768, 16, 7), (300, 64, 3), (256, 128, 5)])
def test_tcs_many_tiles_per_cta(K, D, tiles_per_cta):
    """Several tiles (odd and even counts: pairs and a lone last tile) per CTA, so every ring wraps and both TMEM
    buffers / epilogue groups see many chunk items; ids-only calls take no finish pass."""
    rs = np.random.RandomState(K + 7 * D + tiles_per_cta)
    n = 128 * 148 * tiles_per_cta - 61
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
    dev = _dev()
    zt, Et = torch.from_numpy(z).to(dev), torch.from_numpy(E).to(dev)
    out = ops.forward(zt, Et, 0.25, path="tc")
    _check_against_oracle(z, E, out, n, D)
    ids = ops.forward(zt, Et, 0.25, path="tc", want_zq=False, want_loss=False)
    assert np.array_equal(ids[3].cpu().numpy(), out[3].cpu().numpy())
    assert np.array_equal(ids[4].cpu().numpy(), out[4].cpu().numpy())


def test_tcs_degenerate_codebook_overflows_the_queue():
    """Every code duplicated: no vector can be certified, the per-CTA queues (2048 entries) overflow and the rest is
    decided by the exact scan inside the main loop -- slow, but still the oracle's ids (lowest index of each pair)."""
    rs = np.random.RandomState(11)
    K, D = 512, 32
    n = 148 * 2048 + 148 * 128 * 2
    E = (0.1 * rs.standard_normal((K, D))).astype(np.float32)
    E[256:] = E[:256]
    z = (0.1 * rs.standard_normal((n, D))).astype(np.float32)
    dev = _dev()
    out = ops.forward(torch.from_numpy(z).to(dev), torch.from_numpy(E).to(dev), 0.25, path="tc", want_stats=True)
    _check_against_oracle(z, E, out, n, D)
    assert int(out[5][1].item()) == n                         # every vector went through an exact scan
    assert int(out[3].max().item()) < 256


@pytest.mark.parametrize("K,D,zscale", [(256, 32, 0.1), (256, 32, 30.0), (200, 24, 1.0), (512, 32, 0.1), (256, 64, 0.1)])
def test_tc_twin_codes_at_every_scale(K, D, zscale):
    """The certificate's boundary, probed at every scale at once: every code has a twin at a relative distance drawn
    log-uniformly from 1e-8 .. 1e-1 (and a few exact twins), so that for every vector the runner-up is closer to the
    winner than, near, or beyond the filter radius of either filter.  A radius that is too small for the tensor cores'
    actual arithmetic shows up here as ids that differ from the exact CUDA-core kernel."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(K * 7 + D)
    n = 1 << 20
    base = 0.1 * torch.randn(K // 2, D, device=dev, generator=g)
    rel = 10.0 ** (-8.0 + 7.0 * torch.rand(K // 2, 1, device=dev, generator=g))
    twin = base * (1.0 + rel * torch.randn(K // 2, D, device=dev, generator=g))
    twin[:4] = base[:4]                                          # exact twins: the lower index must win
    E = torch.stack([base, twin], dim=1).reshape(-1, D).contiguous()
    if E.shape[0] < K:
        E = torch.cat([E, 0.1 * torch.randn(K - E.shape[0], D, device=dev, generator=g)])
    z = zscale * torch.randn(n, D, device=dev, generator=g)
    a = ops.forward(z, E, 0.25, path="tc", want_stats=True)
    b = ops.forward(z, E, 0.25, path="fma")
    assert int((a[3] != b[3]).sum().item()) == 0
    assert torch.equal(a[1], b[1]) and torch.equal(a[4], b[4])
    assert a[0].item() == pytest.approx(b[0].item(), rel=REL)
    assert int(a[5][1].item()) > n // 50                          # the case does exercise the exact paths


@pytest.mark.parametrize("K", [256, 100, 1])
@pytest.mark.parametrize("n", [128, 129, 255, 1000, 128 * 148 * 2 + 77])
def test_backward_tma_ring_kernel(K, n):
    """D = 32, K <= 256, contiguous rows, both gradients wanted: the TMA-ring kernel (csrc/vq_bwd.cu).  Ragged and
    odd-length last tiles (the ids travel as 16-byte granules; an odd last row is fetched separately), fewer codes
    than the shared-memory accumulator holds, and the same call on a strided view of z (generic kernel) as a
    cross-check.  fp64 closed form of the oracle; grad_z 1e-5 relative, grad_E 1e-5 of its largest entry."""
    dev = _dev()
    rng = np.random.default_rng(n * 1000 + K)
    D = 32
    z_np = (0.1 * rng.standard_normal((n, D))).astype(np.float32)
    E = (0.1 * rng.standard_normal((K, D))).astype(np.float32)
    g_np = rng.standard_normal((n, D)).astype(np.float32)
    z = torch.from_numpy(z_np).to(dev)
    w = torch.from_numpy(E).to(dev)
    g = torch.from_numpy(g_np).to(dev)
    gl = torch.tensor(1.7, device=dev)
    idx = ops.forward(z, w, 0.25, path="fma")[3]
    gz, gE = ops.backward(g, gl, z, idx, w, 0.25)
    torch.cuda.synchronize()
    ogz, ogE = O.backward(g_np, 1.7, z_np, idx.cpu().numpy().reshape(-1), E, 0.25)
    np.testing.assert_allclose(gz.cpu().numpy(), ogz, rtol=REL, atol=1e-7)
    scale = np.abs(ogE).max() + 1e-30
    np.testing.assert_allclose(gE.cpu().numpy(), ogE, rtol=1e-4, atol=REL * scale)
    # the generic kernel on a strided view of the same vectors
    wide = torch.zeros(n, 2 * D, device=dev)
    wide[:, :D] = z
    gz2, gE2 = ops.backward(g, gl, wide[:, :D], idx, w, 0.25)
    assert torch.equal(gz, gz2)
    np.testing.assert_allclose(gE2.cpu().numpy(), gE.cpu().numpy(), rtol=1e-4, atol=REL * scale)


def test_backward_tma_ring_ignores_out_of_range_ids():
    """Ids outside [0, K) contribute a zero residual (grad_z = g_zq, nothing added to grad_E) in every backward
    kernel; the TMA-ring kernel and the generic one agree on it."""
    dev = _dev()
    torch.manual_seed(11)
    n, D, K = 1000, 32, 64
    z = 0.1 * torch.randn(n, D, device=dev)
    w = 0.1 * torch.randn(K, D, device=dev)
    g = torch.randn(n, D, device=dev)
    gl = torch.tensor(0.9, device=dev)
    idx = ops.forward(z, w, 0.25)[3].clone()
    idx[5] = -3
    idx[77] = K
    idx[999] = 1 << 40
    gz, gE = ops.backward(g, gl, z, idx, w, 0.25)
    wide = torch.zeros(n, 2 * D, device=dev)
    wide[:, :D] = z
    gz2, gE2 = ops.backward(g, gl, wide[:, :D], idx, w, 0.25)
    assert torch.equal(gz, gz2)
    for r in (5, 77, 999):
        assert torch.equal(gz[r], g[r])
    torch.testing.assert_close(gE, gE2, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("B,T,D", [(1000, 16, 32), (37, 16, 32), (5, 8, 64), (300, 3, 8), (64, 64, 128)])
def test_pack_rows_kernel(B, T, D):
    """vqb_pack_rows: the reference encoder's permuted view (logical (B, T, D) over physical (B, D, T)) -> contiguous rows,
    bit-equal to torch's strided copy; other layouts fall back to that copy."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(B + T + D)
    phys = torch.randn(B, D, T, device=dev, generator=g)
    view = phys.permute(0, 2, 1)
    packed = ops.pack_rows(view, D)
    assert packed.is_contiguous() and torch.equal(packed, view.contiguous())
    wide = torch.randn(B, D + 3, T, device=dev, generator=g)[:, :D, :].permute(0, 2, 1)   # gap between the blocks
    assert torch.equal(ops.pack_rows(wide, D), wide.contiguous())
    other = torch.randn(B, T, 2 * D, device=dev, generator=g)[:, :, :D]                   # not the encoder's layout
    assert torch.equal(ops.pack_rows(other, D), other.contiguous())
