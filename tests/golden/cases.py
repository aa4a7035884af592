"""Seeded input synthesis shared by oracle/make_golden.py (which runs the
unmodified reference on these inputs) and the tests (which re-create the same
inputs and compare against the stored reference outputs).

Inputs come from numpy's frozen legacy ``RandomState`` stream so they are
bit-reproducible on any machine; only reference OUTPUTS are stored in
``vq_golden.npz``.
"""
from __future__ import annotations

import numpy as np

# name, K, D, logical z shape, layout, z kind, codebook kind, beta, seed, backward?
CASES = [
    dict(name="default_cfg", K=256, D=32, shape=(2048, 16, 32), layout="contig", z="randn0.1", cb="default", beta=0.25, seed=11, bwd=True),
    dict(name="trained_cb", K=256, D=32, shape=(256, 16, 32), layout="contig", z="clustered", cb="randn0.1", beta=0.25, seed=12, bwd=True),
    dict(name="stress_randn", K=256, D=32, shape=(1024, 16, 32), layout="contig", z="randn", cb="default", beta=0.25, seed=13, bwd=False),
    dict(name="permuted_view", K=256, D=32, shape=(64, 16, 32), layout="permuted", z="randn0.1", cb="default", beta=0.25, seed=14, bwd=True),
    dict(name="four_d", K=64, D=32, shape=(2, 4, 4, 32), layout="contig", z="randn0.1", cb="default", beta=0.5, seed=15, bwd=True),
    dict(name="ties", K=16, D=8, shape=(300, 8), layout="contig", z="on_codes", cb="dup_rows", beta=0.25, seed=16, bwd=True),
    dict(name="nonfinite_z", K=8, D=4, shape=(64, 4), layout="contig", z="nonfinite", cb="default", beta=0.25, seed=17, bwd=False),
    dict(name="nonfinite_cb", K=8, D=4, shape=(64, 4), layout="contig", z="randn", cb="nan_rows", beta=0.25, seed=18, bwd=False),
    dict(name="k1", K=1, D=8, shape=(129, 8), layout="contig", z="randn0.1", cb="default", beta=0.25, seed=19, bwd=True),
    dict(name="k2_d1", K=2, D=1, shape=(1000, 1), layout="contig", z="randn", cb="randn0.1", beta=0.25, seed=20, bwd=True),
    dict(name="k255_d33", K=255, D=33, shape=(1000, 33), layout="contig", z="randn0.1", cb="randn0.1", beta=0.25, seed=21, bwd=True),
    dict(name="k257_d64", K=257, D=64, shape=(513, 64), layout="contig", z="randn0.1", cb="default", beta=1.0, seed=22, bwd=True),
    dict(name="k64_d8", K=64, D=8, shape=(4097, 8), layout="contig", z="randn0.1", cb="default", beta=0.25, seed=23, bwd=False),
    dict(name="k128_d16", K=128, D=16, shape=(4096, 16), layout="contig", z="clustered", cb="randn0.1", beta=0.25, seed=24, bwd=False),
    dict(name="k1024_d64", K=1024, D=64, shape=(512, 64), layout="contig", z="randn", cb="default", beta=0.25, seed=25, bwd=False),
    dict(name="k8192_d32", K=8192, D=32, shape=(256, 32), layout="contig", z="randn", cb="default", beta=0.25, seed=26, bwd=False),
    dict(name="k512_d128", K=512, D=128, shape=(384, 128), layout="contig", z="randn0.1", cb="randn0.1", beta=0.25, seed=27, bwd=True),
    dict(name="single_row", K=256, D=32, shape=(1, 32), layout="contig", z="randn0.1", cb="default", beta=0.25, seed=28, bwd=True),
    dict(name="permuted_ragged", K=256, D=32, shape=(37, 16, 32), layout="permuted", z="clustered", cb="randn0.1", beta=0.25, seed=29, bwd=True),
]

GRAD_ROWS = 256  # number of leading grad_z / z_q rows stored verbatim


def case_by_name(name: str) -> dict:
    for c in CASES:
        if c["name"] == name:
            return c
    raise KeyError(name)


def make_codebook(case: dict, rs: np.random.RandomState) -> np.ndarray:
    K, D, kind = case["K"], case["D"], case["cb"]
    if kind == "default":       # reference init, model/vector_quantizer.py:74
        E = rs.uniform(-1.0 / K, 1.0 / K, size=(K, D))
    elif kind == "randn0.1":    # trained-scale codebook
        E = 0.1 * rs.standard_normal((K, D))
    elif kind == "dup_rows":    # exact duplicates: the lowest index must win
        E = 0.1 * rs.standard_normal((K, D))
        E[7] = E[3]
        E[11] = E[3]
        E[15] = E[0]
    elif kind == "nan_rows":
        E = 0.1 * rs.standard_normal((K, D))
        E[5, 1] = np.nan
        E[2, 3] = np.inf
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(E, dtype=np.float32)


def make_inputs(case: dict):
    """Returns (z_storage, z_logical_view, codebook).  ``z_logical_view`` is a
    numpy view of ``z_storage`` with the logical shape ``case['shape']``; for
    layout "permuted" the storage is (B, D, T) and the view is its
    transpose(0, 2, 1) -- the strides the reference encoder really produces
    (model/vq_vae_patch_embedd.py:91)."""
    rs = np.random.RandomState(case["seed"])
    E = make_codebook(case, rs)
    shape = tuple(case["shape"])
    D = case["D"]
    n = int(np.prod(shape)) // D
    kind = case["z"]
    if kind == "randn0.1":
        flat = 0.1 * rs.standard_normal((n, D))
    elif kind == "randn":
        flat = rs.standard_normal((n, D))
    elif kind == "clustered":
        pick = rs.randint(0, case["K"], size=n)
        clean = np.where(np.isfinite(E), E, 0.0).astype(np.float64)
        flat = clean[pick] + 0.03 * rs.standard_normal((n, D))
    elif kind == "on_codes":    # some rows sit exactly on (duplicated) codes, some between two codes
        pick = rs.randint(0, case["K"], size=n)
        flat = E[pick].astype(np.float64)
        half = n // 2
        flat[half:] += 0.05 * rs.standard_normal((n - half, D))
        mid = (E[1].astype(np.float64) + E[2].astype(np.float64)) / 2
        flat[:8] = mid
    elif kind == "nonfinite":
        flat = rs.standard_normal((n, D))
        flat[1, 2] = np.nan
        flat[2, 0] = np.inf
        flat[3, 1] = -np.inf
        flat[4, :] = np.nan
        flat[5, 0] = np.inf
        flat[5, 1] = -np.inf
        flat[6, :] = 0.0
        flat[7, 3] = 3.0e38
    else:
        raise ValueError(kind)
    flat = np.ascontiguousarray(flat, dtype=np.float32)
    logical = flat.reshape(shape)
    if case["layout"] == "permuted":
        assert len(shape) == 3
        storage = np.ascontiguousarray(logical.transpose(0, 2, 1))  # (B, D, T)
        return storage, storage.transpose(0, 2, 1), E
    return logical, logical, E


def upstream_weights(case: dict) -> np.ndarray:
    """Weights w of the scalar L = g_loss * loss + sum(w * z_q) used to drive
    the backward pass (so g_zq = w)."""
    rs = np.random.RandomState(case["seed"] + 1000)
    return np.ascontiguousarray(rs.standard_normal(tuple(case["shape"])), dtype=np.float32)


G_LOSS = 1.7


# ---- small VQ-VAE-Patch configs whose weights are stored with the fixture ----
PATCH_CASES = [
    dict(name="patch_small", hidden_dim=32, input_dim=2, num_embeddings=16, embedding_dim=8,
         n_resblocks=2, patch_size=25, seq_len=200, batch_norm=False, beta=0.25, batch=8, seed=101),
    dict(name="patch_small_bn", hidden_dim=32, input_dim=2, num_embeddings=32, embedding_dim=8,
         n_resblocks=1, patch_size=25, seq_len=200, batch_norm=True, beta=0.25, batch=8, seed=102),
    dict(name="patch_small_p10", hidden_dim=16, input_dim=2, num_embeddings=16, embedding_dim=4,
         n_resblocks=1, patch_size=10, seq_len=200, batch_norm=False, beta=0.25, batch=4, seed=103),
]


def make_cycles(case: dict) -> np.ndarray:
    rs = np.random.RandomState(case["seed"])
    return np.ascontiguousarray(rs.standard_normal((case["batch"], case["seq_len"], case["input_dim"])),
                                dtype=np.float32)


# ---- a patch model wide enough for the fused tcgen05 encoder layers (hidden_dim a multiple of 256) ----
# Only the tensors the ENCODE path reads are stored with the fixture (patch projection, centre taps of the residual
# blocks' k=3 convolutions -- the outer taps only ever meet zero padding, SURVEY.md appendix A.7 --, the 1x1
# projection and the codebook), together with the reference's z_e and ids.
PATCH_WIDE_CASE = dict(name="patch_wide", hidden_dim=256, input_dim=2, num_embeddings=64, embedding_dim=32,
                       n_resblocks=2, patch_size=25, seq_len=200, batch_norm=False, beta=0.25, batch=256, seed=104)


# ---- the reference's bulk loops over overlapping windows (dataloader/latentspace_dataloader.py:171-263) ----
# A stream of `n_stream` cycles is cut into windows of `seq_len` cycles with a stride of one cycle (the layout
# dataloader/asimow_dataloader.py:185-206 produces), batched like a DataLoader would; the model is PATCH_CASES[0].
BULK_CASE = dict(name="bulk_overlap", model="patch_small", seq_len=4, n_stream=15, batch=4, seed=105)


def make_stream(case: dict):
    """-> (cycles (n_stream, 200, 2) float32 -- the stream make_windows slides over --, one label per CYCLE (n_stream,) float64)."""
    rs = np.random.RandomState(case["seed"])
    stream = rs.standard_normal((case["n_stream"], 200, 2)).astype(np.float32)
    cycle_labels = np.random.RandomState(case["seed"] + 1).randint(0, 2, case["n_stream"]).astype(np.float64)
    return stream, cycle_labels


def make_windows(case: dict):
    """-> (windows (n, seq_len*200, 2) float32, labels (n,) float32, list of (lo, hi) batch slices)."""
    rs = np.random.RandomState(case["seed"])
    stream = rs.standard_normal((case["n_stream"], 200, 2)).astype(np.float32)
    n = case["n_stream"] - case["seq_len"] + 1
    win = np.stack([stream[i:i + case["seq_len"]].reshape(case["seq_len"] * 200, 2) for i in range(n)])
    labels = rs.randint(0, 2, n).astype(np.float32)
    slices = [(lo, min(lo + case["batch"], n)) for lo in range(0, n, case["batch"])]
    return np.ascontiguousarray(win), labels, slices


def unexplained_mismatches(z_e: np.ndarray, codebook: np.ndarray, got: np.ndarray, ref: np.ndarray, rel: float) -> int:
    """Ids may differ from the reference's only where the reference's own z_e puts the two codes at (nearly) the same
    distance: |d(z, E[got]) - d(z, E[ref])| <= rel * (|z|^2 + |e|^2), evaluated in float64 on the REFERENCE z_e.
    `rel` is the relative accuracy of the encoder that produced `got` (about 1e-5 for an fp32 encoder with another
    summation order; 2^-8-ish for bf16 operands).  Returns the number of mismatches this does not explain."""
    z = np.asarray(z_e, dtype=np.float64).reshape(-1, codebook.shape[1])
    E = np.asarray(codebook, dtype=np.float64)
    got = np.asarray(got).reshape(-1)
    ref = np.asarray(ref).reshape(-1)
    bad = np.nonzero(got != ref)[0]
    if bad.size == 0:
        return 0
    zb = z[bad]
    dg = ((zb - E[got[bad]]) ** 2).sum(1)
    dr = ((zb - E[ref[bad]]) ** 2).sum(1)
    scale = (zb ** 2).sum(1) + np.maximum((E[got[bad]] ** 2).sum(1), (E[ref[bad]] ** 2).sum(1))
    return int((np.abs(dg - dr) > rel * scale).sum())
