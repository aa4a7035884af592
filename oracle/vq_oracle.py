"""Python face of the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package never
does; it fails loudly when its CUDA library is missing.

Two restatements of the reference live here:

* ``forward`` / ``backward`` / ``distances`` / ``gather`` -- ctypes bindings to
  ``vq_oracle.c`` (fixed fp32 "oracle order", see that file's header).  This
  is the checker for the CUDA kernels: indices must match bit for bit.
* ``torch_port_forward`` -- the reference's own op sequence
  (model/vector_quantizer.py:88-119: GEMM distances, one-hot, GEMM gather) on
  torch CPU tensors.  It is what the reference costs on host cores, so it is
  the ``cpu_baseline`` (kind "port") that bench.py times.

Parity pin: no golden vectors ship with the reference; both restatements are
checked against outputs of the unmodified reference module
(oracle/make_golden.py -> tests/golden/*.npz, tests/test_oracle_golden.py).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import NamedTuple, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvq_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile vq_oracle.c with the committed Makefile (building the checker is
    not using it)."""
    src = os.path.join(_HERE, "vq_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libvq_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    lib = ctypes.CDLL(_LIB_PATH)
    f32p = ctypes.POINTER(ctypes.c_float)
    f64p = ctypes.POINTER(ctypes.c_double)
    i64p = ctypes.POINTER(ctypes.c_int64)
    lib.vq_oracle_forward.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int,
                                      ctypes.c_float, i64p, f32p, i64p, f32p, f32p, ctypes.c_int]
    lib.vq_oracle_forward.restype = ctypes.c_int
    lib.vq_oracle_distances.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int,
                                        f32p, f64p]
    lib.vq_oracle_distances.restype = ctypes.c_int
    lib.vq_oracle_gather.argtypes = [i64p, ctypes.c_int64, f32p, ctypes.c_int, ctypes.c_int, f32p]
    lib.vq_oracle_gather.restype = ctypes.c_int
    lib.vq_oracle_backward.argtypes = [f32p, ctypes.c_float, f32p, i64p, f32p, ctypes.c_int64,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_float, f32p, f32p]
    lib.vq_oracle_backward.restype = ctypes.c_int
    lib.vq_oracle_num_threads.restype = ctypes.c_int
    _lib = lib
    return lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty):
    return a.ctypes.data_as(ctypes.POINTER(ty)) if a is not None else None


class ForwardResult(NamedTuple):
    loss: np.float32
    z_q: np.ndarray          # z.shape, fp32
    perplexity: np.float32
    indices: np.ndarray      # (N, 1) int64
    counts: np.ndarray       # (K,) int64


def num_threads() -> int:
    return int(_load().vq_oracle_num_threads())


def forward(z, codebook, beta: float, nthreads: int = 0, want_zq: bool = True) -> ForwardResult:
    """Oracle-order restatement of model/vector_quantizer.py:88-119 (the (N,K)
    one-hot is implied by ``indices``)."""
    lib = _load()
    z = _f32(z)
    E = _f32(codebook)
    k, d = E.shape
    if z.size % d:
        raise ValueError(f"numel {z.size} not divisible by e_dim {d}")
    flat = z.reshape(-1, d)
    n = flat.shape[0]
    idx = np.empty((n,), np.int64)
    zq = np.empty_like(flat) if want_zq else None
    counts = np.zeros((k,), np.int64)
    loss = np.zeros((1,), np.float32)
    ppl = np.zeros((1,), np.float32)
    rc = lib.vq_oracle_forward(_p(flat, ctypes.c_float), n, d, _p(E, ctypes.c_float), k,
                               ctypes.c_float(beta), _p(idx, ctypes.c_int64),
                               _p(zq, ctypes.c_float), _p(counts, ctypes.c_int64),
                               _p(loss, ctypes.c_float), _p(ppl, ctypes.c_float), nthreads)
    if rc != 0:
        raise RuntimeError(f"vq_oracle_forward failed: {rc}")
    return ForwardResult(loss[0], zq.reshape(z.shape) if want_zq else None, ppl[0],
                         idx.reshape(n, 1), counts)


def distances(z, codebook, want32: bool = True, want64: bool = True):
    """(dist32, dist64): oracle-order fp32 distances and their fp64 counterpart."""
    lib = _load()
    E = _f32(codebook)
    k, d = E.shape
    flat = _f32(z).reshape(-1, d)
    n = flat.shape[0]
    d32 = np.empty((n, k), np.float32) if want32 else None
    d64 = np.empty((n, k), np.float64) if want64 else None
    rc = lib.vq_oracle_distances(_p(flat, ctypes.c_float), n, d, _p(E, ctypes.c_float), k,
                                 _p(d32, ctypes.c_float), _p(d64, ctypes.c_double))
    if rc != 0:
        raise RuntimeError(f"vq_oracle_distances failed: {rc}")
    return d32, d64


def gather(indices, codebook, target_shape=None) -> np.ndarray:
    """model/vector_quantizer.py:121-131."""
    lib = _load()
    E = _f32(codebook)
    k, d = E.shape
    idx = np.ascontiguousarray(indices, dtype=np.int64).reshape(-1)
    out = np.empty((idx.shape[0], d), np.float32)
    rc = lib.vq_oracle_gather(_p(idx, ctypes.c_int64), idx.shape[0], _p(E, ctypes.c_float), k, d,
                              _p(out, ctypes.c_float))
    if rc != 0:
        raise RuntimeError(f"vq_oracle_gather failed: {rc}")
    return out.reshape(target_shape) if target_shape is not None else out


def backward(g_zq: Optional[np.ndarray], g_loss: float, z, indices, codebook, beta: float):
    """Closed-form autograd of the forward, fp64 inside.
    Returns (grad_z with z's shape, grad_E (K, D))."""
    lib = _load()
    E = _f32(codebook)
    k, d = E.shape
    zc = _f32(z)
    flat = zc.reshape(-1, d)
    n = flat.shape[0]
    idx = np.ascontiguousarray(indices, dtype=np.int64).reshape(-1)
    g = _f32(g_zq).reshape(-1, d) if g_zq is not None else None
    gz = np.empty_like(flat)
    gE = np.empty_like(E)
    rc = lib.vq_oracle_backward(_p(g, ctypes.c_float), ctypes.c_float(g_loss),
                                _p(flat, ctypes.c_float), _p(idx, ctypes.c_int64),
                                _p(E, ctypes.c_float), n, d, k, ctypes.c_float(beta),
                                _p(gz, ctypes.c_float), _p(gE, ctypes.c_float))
    if rc != 0:
        raise RuntimeError(f"vq_oracle_backward failed: {rc}")
    return gz.reshape(zc.shape), gE


def explain_mismatches(z, codebook, idx_a, idx_b, ulps: float = 8.0):
    """For rows where two index vectors disagree, decide whether the
    disagreement is an fp32 near-tie: the two chosen codes' fp64 distances
    differ by at most ``ulps`` units in the last place of the fp32 distance
    magnitude (which is rounded at ulp(zz + ee), SURVEY.md section 0 trap iii).

    Returns dict(mismatch=int, explained=int, worst_ulps=float, rows=list).
    """
    E = _f32(codebook)
    d = E.shape[1]
    flat = _f32(z).reshape(-1, d)
    a = np.asarray(idx_a).reshape(-1)
    b = np.asarray(idx_b).reshape(-1)
    rows = np.nonzero(a != b)[0]
    if rows.size == 0:
        return dict(mismatch=0, explained=0, worst_ulps=0.0, rows=[])
    _, d64 = distances(flat[rows], E, want32=False, want64=True)
    da = d64[np.arange(rows.size), a[rows]]
    db = d64[np.arange(rows.size), b[rows]]
    zz = (flat[rows].astype(np.float64) ** 2).sum(1)
    ee = np.maximum((E[a[rows]].astype(np.float64) ** 2).sum(1),
                    (E[b[rows]].astype(np.float64) ** 2).sum(1))
    scale = np.maximum(zz + ee, np.finfo(np.float32).tiny)
    ulp = np.spacing(scale.astype(np.float32)).astype(np.float64)
    gap = np.abs(da - db) / ulp
    nonfinite = ~np.isfinite(gap)
    ok = (gap <= ulps) & ~nonfinite
    return dict(mismatch=int(rows.size), explained=int(ok.sum()),
                worst_ulps=float(np.nanmax(np.where(nonfinite, np.nan, gap))) if (~nonfinite).any() else float("nan"),
                rows=rows.tolist())


# --------------------------------------------------------------------------------------
# The reference's own op sequence on torch CPU tensors (the cpu_baseline "port").
# --------------------------------------------------------------------------------------
def torch_port_forward(z, weight, beta: float):
    """Follows model/vector_quantizer.py:88-119 op by op on whatever device the
    tensors live on (used on CPU).  Returns the reference's 5-tuple."""
    import torch

    k, d = weight.shape
    rows = z.reshape(-1, d)                                            # :88
    sq_z = (rows ** 2).sum(dim=1, keepdim=True)                        # :91
    sq_e = (weight ** 2).sum(dim=1)                                    # :92
    dist = sq_z + sq_e - 2 * torch.matmul(rows, weight.t())            # :91-93
    nearest = dist.argmin(dim=1).unsqueeze(1)                          # :96
    onehot = torch.zeros(nearest.shape[0], k, dtype=rows.dtype, device=rows.device)
    onehot.scatter_(1, nearest, 1)                                     # :98-100
    picked = torch.matmul(onehot, weight).view(z.shape)                # :103
    loss = ((picked.detach() - z) ** 2).mean() + beta * ((picked - z.detach()) ** 2).mean()  # :107-108
    out = z + (picked - z).detach()                                    # :111
    usage = onehot.mean(dim=0)                                         # :114
    perplexity = torch.exp(-(usage * torch.log(usage + 1e-10)).sum())  # :115
    return loss, out.contiguous(), perplexity, onehot, nearest         # :118-119


def torch_port_encode(sd: dict, x, patch_size: int = 25):
    """The reference's encode call on torch CPU tensors, with its own loop structure: PatchEmbedding.forward
    (model/vq_vae_patch_embedd.py:13-17), CNNBlock.forward with seperate=True (:103-111: a Python loop over the token
    positions, every ResBlock (:60-74) applied to a LENGTH-1 slice through Conv1d(k=3, pad=1)), SepCNNBlock.forward
    (:83-91: one 1x1 conv per position, cat, permute).  `sd` is a VQVAEPatch state dict (batch_norm=False, dropout in
    eval mode = identity).  Returns z_e as the reference does: logical (B, T, D), a permuted view of (B, D, T)."""
    import torch
    import torch.nn.functional as F

    b = x.shape[0]
    t = x.permute(0, 2, 1).reshape(b, 1, -1)                                                    # :14-15
    t = F.conv1d(t, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=patch_size)   # :16  (B, H, T)
    n_blocks = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("encoder.0.shared_conv."))
    cols = []
    for i in range(t.shape[2]):                                                                 # :106
        h = t[:, :, i].unsqueeze(2)                                                             # :107
        for j in range(n_blocks):                                                               # :109 (nn.Sequential of ResBlocks)
            pre = f"encoder.0.shared_conv.{j}.block."
            u = F.conv1d(F.gelu(h), sd[pre + "1.weight"], sd[pre + "1.bias"], padding=1)        # :64-65
            u = F.conv1d(F.gelu(u), sd[pre + "4.weight"], sd[pre + "4.bias"], padding=1)        # :67-68
            h = h + u                                                                           # :74
        cols.append(h)
    t = torch.cat(cols, dim=2)                                                                  # :111
    outs = [F.conv1d(t[:, :, i].unsqueeze(2), sd["encoder.1.shared_conv.weight"], sd["encoder.1.shared_conv.bias"])
            for i in range(t.shape[2])]                                                         # :86-89
    return torch.cat(outs, dim=2).permute(0, 2, 1)                                              # :90-91
