"""Tensor-level entry points: torch tensors in, C-ABI calls out.

PyTorch is plumbing here (device memory from the caching allocator, the current
stream, autograd bookkeeping); all arithmetic happens inside libvqb200.so.
"""
from __future__ import annotations

import collections
import ctypes
import weakref
from typing import Optional, Tuple

import torch

from . import _lib

PATHS = {"auto": _lib.PATH_AUTO, "fma": _lib.PATH_FMA, "tc": _lib.PATH_TC}
KEEP_CODEBOOK, KEEP_TC_IMAGE = 4, 8            # VQB_KEEP_* (include/vqb200.h)
IDS_FLAGS = {"int64": 0, "uint8": 16, "uint16": 32}   # VQB_IDS_*

#: reuse what the workspace holds about a codebook (code norms, census, tcgen05 operand image) while the caller passes the
#: SAME tensor object with the same (data_ptr, _version) -- in inference the codebook kernels then run once, not once per call.  In-place
#: updates through autograd-visible ops (optimizers, copy_, load_state_dict) bump _version; writing through `.data`
#: does not: call clear_workspaces() after such a write, or switch this off.
CACHE_CODEBOOK = True

# (device, k, d, stream handle) -> workspace tensor; least recently used entries are dropped beyond _WS_MAX (one entry
# holds ~3.2 MB, mostly the fix-up queues), clear_workspaces() empties it
_workspaces: "collections.OrderedDict" = collections.OrderedDict()
_ws_state = {}          # id(workspace tensor) -> (weakref to the weight tensor, data_ptr, _version, tcgen05 image valid)
_WS_MAX = 16


def set_filter(mode: Optional[str]) -> None:
    """Debug / A-B: the filter of the tcgen05 forward kernel -- None (automatic), "bf16" (three bf16 products) or "tf32"
    (one tf32 product).  Cached operand images belong to a filter: the workspaces are dropped."""
    code = {None: -1, "auto": -1, "bf16": 0, "tf32": 1}[mode]
    _lib.check(_lib.load().vqb_debug_set_filter(code), "vqb_debug_set_filter")
    clear_workspaces()


def clear_workspaces() -> None:
    """Drop every cached kernel workspace (their memory returns to the caching allocator)."""
    _workspaces.clear()
    _ws_state.clear()


def _require_cuda_fp32(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if t.dtype != torch.float32:
        # the reference module is fp32-only as well
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} lives on {t.device}: the B200 vector quantiser has no CPU fallback; "
            "move the module and its input to a CUDA device")


def _workspace(device: torch.device, k: int, d: int) -> torch.Tensor:
    stream = torch.cuda.current_stream(device).cuda_stream
    key = (device.index, k, d, stream)
    ws = _workspaces.get(key)
    if ws is None:
        nbytes = _lib.load().vqb_workspace_bytes(k, d)
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
        while len(_workspaces) > _WS_MAX:
            _, old = _workspaces.popitem(last=False)
            _ws_state.pop(id(old), None)
    else:
        _workspaces.move_to_end(key)
    return ws


def _ws_ptr(ws: torch.Tensor) -> Tuple[int, int]:
    p = ws.data_ptr()
    a = (p + 255) // 256 * 256
    return a, ws.numel() - (a - p)


def _tc_eligible(n: int, k: int, d: int) -> bool:
    """Shapes the tcgen05 kernels take (vq_fwd_tc.cu: tc_shape_supported; vq_fwd_tcs.cu: tcs_shape_supported) once rows are
    contiguous."""
    return 4 <= d <= 128 and d % 4 == 0 and 1 <= k <= 16384 and n >= 128


def pack_rows(z: torch.Tensor, d: int) -> torch.Tensor:
    """Contiguous (.., d) rows of a strided view.  The reference encoder's layout -- logical (B, T, d) over physical
    (B, d, T) -- goes through vqb_pack_rows (a per-cycle transpose with fully coalesced accesses); any other layout takes
    torch's generic strided copy, like the reference's own reshape (model/vector_quantizer.py:88)."""
    if z.is_contiguous():
        return z
    if (z.is_cuda and z.dtype == torch.float32 and z.dim() == 3 and z.shape[2] == d and z.stride(1) == 1
            and z.stride(2) == z.shape[1] and z.shape[1] <= 64 and d <= 128 and z.stride(0) >= d * z.shape[1]):
        out = torch.empty(z.shape, dtype=torch.float32, device=z.device)
        with torch.cuda.device(z.device):
            rc = _lib.load().vqb_pack_rows(z.device.index, z.data_ptr(), z.shape[0], z.shape[1], d, z.stride(0), 1,
                                           z.stride(2), out.data_ptr(), torch.cuda.current_stream(z.device).cuda_stream)
        if rc == 0:
            return out
        if rc != _lib.E_UNSUPPORTED:
            _lib.check(rc, "vqb_pack_rows")
    return z.contiguous()


def _view_params(z: torch.Tensor, d: int, k: int = 0, path: str = "fma"):
    """Express z (logical row-major order of reshape(-1, d)) as (n_outer, n_inner, d) with
    element strides, without copying when the layout allows it.  Returns
    (tensor_to_keep_alive, n_outer, n_inner, s_outer, s_inner, s_d).

    A strided view (e.g. the encoder's permuted output) is read in place by the FMA kernel.  When
    the shape qualifies for the tcgen05 kernel, which needs contiguous rows for its TMA tiles, one
    packing copy (8*d bytes per vector) buys a ~4x faster quantiser, so it is made here."""
    n = z.numel() // d
    if z.is_contiguous():
        return z, n, 1, d, d, 1
    if path in ("auto", "tc") and _tc_eligible(n, k, d):
        return pack_rows(z, d), n, 1, d, d, 1
    if z.dim() >= 2 and z.shape[-1] == d:
        if z.dim() == 2:
            return z, z.shape[0], 1, z.stride(0), 0, z.stride(1)
        if z.dim() == 3:  # e.g. the encoder's permuted view, strides (d*T, 1, T)
            return z, z.shape[0], z.shape[1], z.stride(0), z.stride(1), z.stride(2)
    zc = z.contiguous()  # exotic layout: same copy the reference's reshape makes (:88)
    return zc, n, 1, d, d, 1


def forward(z: torch.Tensor, weight: torch.Tensor, beta: float, path: str = "auto",
            want_zq: bool = True, want_stats: bool = False, want_loss: bool = True):
    """Fused nearest-code search.  Returns (loss, z_q, perplexity, indices (N,1) int64,
    counts (K,) int64[, stats (4,) int64]).  want_zq=False / want_loss=False: that output is None and
    the kernel skips the work behind it (ids-only encoding, dataloader/latentspace_dataloader.py:160-161)."""
    _require_cuda_fp32(z, "z")
    _require_cuda_fp32(weight, "embedding.weight")
    if weight.dim() != 2:
        raise RuntimeError("embedding.weight must be (n_e, e_dim)")
    if z.device != weight.device:
        raise RuntimeError(f"z is on {z.device} but the codebook is on {weight.device}")
    k, d = weight.shape
    if z.numel() % d:
        raise RuntimeError(f"shape '[-1, {d}]' is invalid for input of size {z.numel()}")
    lib = _lib.load()
    dev = z.device
    w = weight.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    zsrc, n_outer, n_inner, s_outer, s_inner, s_d = _view_params(z.detach(), d, k, path)
    n = n_outer * n_inner
    with torch.cuda.device(dev):
        zq = torch.empty(z.shape, dtype=torch.float32, device=dev) if want_zq else None
        idx = torch.empty((n, 1), dtype=torch.int64, device=dev)
        scal = torch.empty(2, dtype=torch.float32, device=dev)
        counts = torch.empty(k, dtype=torch.int64, device=dev)
        stats = torch.empty(4, dtype=torch.int64, device=dev) if want_stats else None
        ws = _workspace(dev, k, d)
        ws_ptr, ws_bytes = _ws_ptr(ws)
        stream = torch.cuda.current_stream(dev).cuda_stream
        # what this workspace already holds about this very codebook (see CACHE_CODEBOOK)
        flags = PATHS[path]
        # (identity of the tensor OBJECT, not only of its address: the caching allocator hands a freed block to the next
        # tensor of the same size, which would then look like the old codebook at version 0)
        tag = (weight.data_ptr(), weight._version)
        state = _ws_state.get(id(ws)) if CACHE_CODEBOOK else None
        if state is not None and state[0]() is weight and state[1:3] == tag and w.data_ptr() == tag[0]:
            flags |= KEEP_CODEBOOK | (KEEP_TC_IMAGE if state[3] else 0)
        contiguous_rows = s_d == 1 and n_inner == 1 and s_outer == d
        takes_tc = (path != "fma" and contiguous_rows and k <= 256 and _tc_eligible(n, k, d) and n < (1 << 31)
                    and zsrc.data_ptr() % 16 == 0)
        _ws_state.pop(id(ws), None)            # nothing is vouched for if the call below fails
        rc = lib.vqb_forward(
            dev.index, zsrc.data_ptr(), n_outer, n_inner, d, s_outer, s_inner, s_d,
            w.data_ptr(), k, float(beta),
            zq.data_ptr() if zq is not None else None, idx.data_ptr(),
            scal.data_ptr() if want_loss else None, scal.data_ptr() + 4, counts.data_ptr(),
            stats.data_ptr() if stats is not None else None,
            ws_ptr, ws_bytes, flags, stream)
    _lib.check(rc, "vqb_forward")
    if CACHE_CODEBOOK:
        _ws_state[id(ws)] = (weakref.ref(weight),) + tag + (takes_tc or bool(flags & KEEP_TC_IMAGE),)
    out = (scal[0] if want_loss else None, zq, scal[1], idx, counts)
    return out + (stats,) if want_stats else out


def backward(g_zq: Optional[torch.Tensor], g_loss: Optional[torch.Tensor], z: torch.Tensor,
             idx: torch.Tensor, weight: torch.Tensor, beta: float,
             need_z: bool = True, need_e: bool = True):
    """Fused straight-through backward.  Returns (grad_z or None, grad_E or None)."""
    lib = _lib.load()
    dev = z.device
    k, d = weight.shape
    w = weight.detach()
    if not w.is_contiguous():
        w = w.contiguous()
    zsrc, n_outer, n_inner, s_outer, s_inner, s_d = _view_params(z.detach(), d)
    if g_zq is not None:
        _require_cuda_fp32(g_zq, "grad of z_q")
        g_zq = g_zq.contiguous()
    if g_loss is not None:
        _require_cuda_fp32(g_loss, "grad of loss")
        g_loss = g_loss.reshape(1).contiguous()
    with torch.cuda.device(dev):
        grad_z = torch.empty(z.shape, dtype=torch.float32, device=dev) if need_z else None
        grad_e = torch.empty((k, d), dtype=torch.float32, device=dev) if need_e else None
        ws = _workspace(dev, k, d)
        ws_ptr, ws_bytes = _ws_ptr(ws)
        stream = torch.cuda.current_stream(dev).cuda_stream
        rc = lib.vqb_backward(
            dev.index, g_zq.data_ptr() if g_zq is not None else None,
            g_loss.data_ptr() if g_loss is not None else None,
            zsrc.data_ptr(), n_outer, n_inner, d, s_outer, s_inner, s_d,
            idx.data_ptr(), w.data_ptr(), k, float(beta),
            grad_z.data_ptr() if grad_z is not None else None,
            grad_e.data_ptr() if grad_e is not None else None,
            ws_ptr, ws_bytes, stream)
    _lib.check(rc, "vqb_backward")
    return grad_z, grad_e


def gather(indices: torch.Tensor, weight: torch.Tensor, target_shape=None, check_range: bool = False) -> torch.Tensor:
    """E[indices] (model/vector_quantizer.py:121-131).  Out-of-range ids give NaN rows and set the kernel's flag;
    check_range=True reads that flag (one device synchronisation) and raises like the reference's scatter_ does."""
    _require_cuda_fp32(weight, "embedding.weight")
    if indices.dtype != torch.int64:
        raise RuntimeError(f"indices must be int64, got {indices.dtype}")
    if indices.device != weight.device:
        raise RuntimeError("indices and codebook must be on the same device")
    lib = _lib.load()
    dev = weight.device
    k, d = weight.shape
    flat = indices.reshape(-1).contiguous()
    n = flat.numel()
    w = weight.detach().contiguous()
    with torch.cuda.device(dev):
        out = torch.empty((n, d), dtype=torch.float32, device=dev)
        bad = torch.zeros(1, dtype=torch.int32, device=dev)
        rc = lib.vqb_gather(dev.index, flat.data_ptr(), n, w.data_ptr(), k, d, out.data_ptr(),
                            bad.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "vqb_gather")
    if check_range and int(bad.item()) != 0:
        raise IndexError(f"gather: an index lies outside [0, {k})")
    return out.view(target_shape) if target_shape is not None else out


def ar_pairs(ids: torch.Tensor, start_token: int, end_token: int):
    """(x, y) of the transformer's next-token task (dataloader/base_dataloader.py:85-95): ids (W, n) int64 CUDA ->
    x = [start, ids...], y = [ids..., end], both (W, n + 1) int64 -- vqb_ar_pairs, one pass."""
    if not isinstance(ids, torch.Tensor) or ids.dtype != torch.int64 or ids.dim() != 2:
        raise RuntimeError("ar_pairs: ids must be a 2-D int64 tensor")
    if not ids.is_cuda:
        raise RuntimeError(f"ar_pairs: ids live on {ids.device}; there is no CPU fallback")
    ids = ids.contiguous()
    w, n = ids.shape
    lib = _lib.load()
    with torch.cuda.device(ids.device):
        x = torch.empty((w, n + 1), dtype=torch.int64, device=ids.device)
        y = torch.empty((w, n + 1), dtype=torch.int64, device=ids.device)
        rc = lib.vqb_ar_pairs(ids.device.index, ids.data_ptr(), w, n, int(start_token), int(end_token), x.data_ptr(),
                              y.data_ptr(), torch.cuda.current_stream(ids.device).cuda_stream)
    _lib.check(rc, "vqb_ar_pairs")
    return x, y


def row_keys(rows: torch.Tensor, mult: torch.Tensor) -> torch.Tensor:
    """(n, 2) int64 fingerprints of the bit patterns of `rows` ((n, ...) of a 4-byte dtype, CUDA): keys[i, s] =
    sum_j uint32(word_ij) * mult[s, j] (wrapping, words zero-extended) with mult (2, words) int64 -- vqb_row_keys, one pass over the rows."""
    if not isinstance(rows, torch.Tensor) or not rows.is_cuda:
        raise RuntimeError("row_keys: rows must be a CUDA tensor; there is no CPU fallback")
    if rows.element_size() != 4:
        raise RuntimeError(f"row_keys: rows must have a 4-byte dtype, got {rows.dtype}")
    n = rows.shape[0]
    if n == 0:
        return torch.empty((0, 2), dtype=torch.int64, device=rows.device)
    flat = rows.reshape(n, -1).contiguous()
    words = flat.shape[1]
    if mult.dtype != torch.int64 or tuple(mult.shape) != (2, words) or mult.device != rows.device or not mult.is_contiguous():
        raise RuntimeError(f"row_keys: mult must be a contiguous (2, {words}) int64 tensor on {rows.device}")
    lib = _lib.load()
    with torch.cuda.device(rows.device):
        keys = torch.empty((n, 2), dtype=torch.int64, device=rows.device)
        rc = lib.vqb_row_keys(rows.device.index, flat.data_ptr(), n, words, mult.data_ptr(), keys.data_ptr(),
                              torch.cuda.current_stream(rows.device).cuda_stream)
    _lib.check(rc, "vqb_row_keys")
    return keys


def dedupe_first(keys: torch.Tensor) -> torch.Tensor:
    """(n,) int64: for every row the smallest row index with the same (n, 2) int64 key pair (CUDA) -- vqb_dedupe_first."""
    if not isinstance(keys, torch.Tensor) or not keys.is_cuda or keys.dtype != torch.int64 or keys.dim() != 2 or keys.shape[1] != 2:
        raise RuntimeError("dedupe_first: keys must be an (n, 2) int64 CUDA tensor; there is no CPU fallback")
    keys = keys.contiguous()
    n = keys.shape[0]
    lib = _lib.load()
    with torch.cuda.device(keys.device):
        nbytes = int(lib.vqb_dedupe_scratch_bytes(n))
        scratch = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=keys.device)
        first = torch.empty(n, dtype=torch.int64, device=keys.device)
        rc = lib.vqb_dedupe_first(keys.device.index, keys.data_ptr(), n, scratch.data_ptr(), scratch.numel() * 8,
                                  first.data_ptr(), torch.cuda.current_stream(keys.device).cuda_stream)
    _lib.check(rc, "vqb_dedupe_first")
    return first


def one_hot(indices: torch.Tensor, k: int) -> torch.Tensor:
    """(N, k) fp32 one-hot of (N, 1) int64 indices (model/vector_quantizer.py:98-100)."""
    if not isinstance(indices, torch.Tensor) or indices.dtype != torch.int64:
        raise RuntimeError(f"one_hot: indices must be an int64 tensor, got {getattr(indices, 'dtype', type(indices))}")
    if not indices.is_cuda:
        raise RuntimeError(f"one_hot: indices live on {indices.device}; there is no CPU fallback")
    if k <= 0:
        raise RuntimeError("one_hot: k must be positive")
    lib = _lib.load()
    dev = indices.device
    flat = indices.reshape(-1).contiguous()
    n = flat.numel()
    with torch.cuda.device(dev):
        out = torch.empty((n, k), dtype=torch.float32, device=dev)
        rc = lib.vqb_one_hot(dev.index, flat.data_ptr(), n, k, out.data_ptr(),
                             torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "vqb_one_hot")
    return out


class VQStraightThrough(torch.autograd.Function):
    """loss, z_q, perplexity, indices, counts = f(z, codebook).  Gradients follow the
    reference's autograd exactly: z receives g_zq (straight-through)
    plus the commitment term, the codebook receives the beta-weighted codebook term."""

    @staticmethod
    def forward(ctx, z, weight, beta, path):
        # A strided view that qualifies for the tcgen05 kernel is packed once HERE, and the packed rows are what the
        # backward sees too: its TMA-ring kernel needs contiguous rows (1.32 ms instead of 1.70 ms at N = 2^24 for the
        # reference encoder's permuted layout), and the forward would have made the same copy anyway.
        k, d = weight.shape
        if not z.is_contiguous() and path in ("auto", "tc") and z.numel() % d == 0 and _tc_eligible(z.numel() // d, k, d):
            z = pack_rows(z, d)
        loss, zq, ppl, idx, counts = forward(z, weight, beta, path)
        ctx.save_for_backward(z, weight, idx)
        ctx.beta = float(beta)
        ctx.mark_non_differentiable(ppl, idx, counts)
        return loss, zq, ppl, idx, counts

    @staticmethod
    def backward(ctx, g_loss, g_zq, _g_ppl, _g_idx, _g_counts):
        z, weight, idx = ctx.saved_tensors
        need_z, need_e = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        grad_z, grad_e = backward(g_zq, g_loss, z, idx, weight, ctx.beta, need_z, need_e)
        return grad_z, grad_e, None, None


class HostEncoder:
    """Owns a vqb_host_ctx: host arrays in, ids / quantised vectors back on the host, with
    H2D / kernel / D2H of successive chunks overlapped (vqb_encode_host)."""

    def __init__(self, codebook, device: int = 0, chunk_rows: int = 1 << 20, depth: int = 3):
        import numpy as np
        self._np = np
        self._lib = _lib.load()
        cb = np.ascontiguousarray(codebook.detach().cpu().numpy() if isinstance(codebook, torch.Tensor) else codebook,
                                  dtype=np.float32)
        self.k, self.d = cb.shape
        self.device = int(device)
        self.chunk_rows = int(chunk_rows)
        self._ctx = ctypes.c_void_p()
        _lib.check(self._lib.vqb_host_create(self.device, self.chunk_rows, self.d, self.k, int(depth),
                                             ctypes.byref(self._ctx)), "vqb_host_create")
        _lib.check(self._lib.vqb_host_set_codebook(self._ctx, cb.ctypes.data), "vqb_host_set_codebook")
        self.last_launches = 0

    def encode(self, z_host, beta: float = 0.25, zq_out=None, idx_out=None, counts_out=None, path: str = "auto"):
        """z_host: (n, d) fp32 contiguous numpy array or CPU tensor (pin it for async copies).  idx_out: int64 (the
        reference's id dtype), or uint8 (k <= 256) / uint16 (k <= 65536) for the same ids in 1 / 2 bytes each.
        Returns (loss, perplexity)."""
        np = self._np

        def ptr(a, name, kinds, min_elems):
            """Raw host pointer of a C-contiguous numpy array / CPU tensor of the given dtype with >= min_elems
            elements; anything else would be misread or overrun by the library, so it raises here."""
            if a is None:
                return None
            if isinstance(a, torch.Tensor):
                if a.is_cuda:
                    raise RuntimeError(f"{name} must live in host memory (got {a.device})")
                dt, contiguous, elems, p = str(a.dtype).replace("torch.", ""), a.is_contiguous(), a.numel(), a.data_ptr()
            elif isinstance(a, np.ndarray):
                dt, contiguous, elems, p = a.dtype.name, a.flags["C_CONTIGUOUS"], a.size, a.ctypes.data
            else:
                raise TypeError(f"{name} must be a numpy array or a CPU torch tensor, got {type(a).__name__}")
            if dt not in kinds:
                raise RuntimeError(f"{name} must have dtype {' or '.join(kinds)}, got {dt}")
            if not contiguous:
                raise RuntimeError(f"{name} must be C-contiguous")
            if elems < min_elems:
                raise RuntimeError(f"{name} holds {elems} elements, the call needs {min_elems}")
            return p

        if not isinstance(z_host, (torch.Tensor, np.ndarray)):
            raise TypeError("z_host must be a numpy array or a CPU torch tensor")
        total = z_host.numel() if isinstance(z_host, torch.Tensor) else z_host.size
        if total % self.d:
            raise RuntimeError(f"z_host holds {total} elements, not a multiple of the vector width {self.d}")
        n = total // self.d
        z_ptr = ptr(z_host, "z_host", ("float32",), n * self.d)
        zq_ptr = ptr(zq_out, "zq_out", ("float32",), n * self.d)
        idx_ptr = ptr(idx_out, "idx_out", ("int64", "uint8", "uint16"), n)
        id_kind = "int64"
        if idx_out is not None:
            id_kind = (str(idx_out.dtype).replace("torch.", "") if isinstance(idx_out, torch.Tensor) else idx_out.dtype.name)
            if (id_kind == "uint8" and self.k > 256) or (id_kind == "uint16" and self.k > 65536):
                raise RuntimeError(f"idx_out dtype {id_kind} cannot hold ids of a {self.k}-entry codebook")
        cnt_ptr = ptr(counts_out, "counts_out", ("int64", "uint64"), self.k)
        loss = ctypes.c_float()
        ppl = ctypes.c_float()
        launches = ctypes.c_int()
        rc = self._lib.vqb_encode_host(self._ctx, z_ptr, int(n), float(beta), zq_ptr, idx_ptr,
                                       ctypes.addressof(loss), ctypes.addressof(ppl), cnt_ptr,
                                       PATHS[path] | IDS_FLAGS[id_kind], ctypes.byref(launches))
        _lib.check(rc, "vqb_encode_host")
        self.last_launches = launches.value
        return loss.value, ppl.value

    def close(self):
        if self._ctx:
            self._lib.vqb_host_destroy(self._ctx)
            self._ctx = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def token_linear(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, h: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, mode: int = 0) -> Optional[torch.Tensor]:
    """One fused per-token linear layer of the encoder's residual blocks (vqb_token_linear, csrc/tok_linear.cu):
    mode 0: out = bf16(gelu(a @ w.T + bias));  mode 1: h += a @ w.T + bias (in place), out = bf16(gelu(h)) if given;
    mode 2: h = a @ w.T + bias (written), out = bf16(gelu(h)) if given.
    a (T, K) bf16, w (N, K) bf16, bias (N,) fp32, h (T, N) fp32, out (T, N) bf16 -- all CUDA, contiguous."""
    for t, name, dt in ((a, "a", torch.bfloat16), (w, "w", torch.bfloat16), (bias, "bias", torch.float32)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous CUDA {dt} tensor (no CPU fallback)")
    t_rows, k = a.shape
    n = w.shape[0]
    if w.shape[1] != k or bias.numel() != n:
        raise RuntimeError("token_linear: shape mismatch")
    if mode == 0 and out is None:
        out = torch.empty((t_rows, n), dtype=torch.bfloat16, device=a.device)
    if mode != 0 and (h is None or h.dtype != torch.float32 or not h.is_contiguous() or tuple(h.shape) != (t_rows, n)):
        raise RuntimeError("token_linear modes 1 and 2 need a contiguous fp32 h of shape (T, N)")
    if out is not None and (out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape) != (t_rows, n)):
        raise RuntimeError("token_linear: out must be a contiguous bf16 (T, N) tensor")
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.vqb_token_linear(a.device.index, a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                  h.data_ptr() if h is not None else None,
                                  out.data_ptr() if out is not None else None,
                                  t_rows, k, n, int(mode), torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "vqb_token_linear")
    return out


def bf16_pair(x: torch.Tensor) -> torch.Tensor:
    """(R, C) fp32 -> (R, 2 C) bf16 = [bf16(x) | bf16(x - bf16(x))]: the operand pair of the split layers (weights; built once
    per weight version with torch ops -- activations are paired by the kernels themselves)."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=1).contiguous()


def token_pair(h: torch.Tensor, gelu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(T, N) fp32 -> (T, 2 N) bf16 pair of gelu(h) (erf form) or of h itself -- vqb_token_pair."""
    if not h.is_cuda or h.dtype != torch.float32 or not h.is_contiguous() or h.dim() != 2:
        raise RuntimeError("token_pair: h must be a contiguous CUDA fp32 (T, N) tensor (no CPU fallback)")
    t_rows, n = h.shape
    if out is None:
        out = torch.empty((t_rows, 2 * n), dtype=torch.bfloat16, device=h.device)
    elif out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape) != (t_rows, 2 * n):
        raise RuntimeError("token_pair: out must be a contiguous bf16 (T, 2 N) tensor")
    if t_rows == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(h.device):
        rc = lib.vqb_token_pair(h.device.index, h.data_ptr(), out.data_ptr(), t_rows, n, int(bool(gelu)),
                                torch.cuda.current_stream(h.device).cuda_stream)
    _lib.check(rc, "vqb_token_pair")
    return out


def token_linear_split(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, h: Optional[torch.Tensor] = None,
                       out: Optional[torch.Tensor] = None, mode: int = 0, out_gelu: bool = True) -> Optional[torch.Tensor]:
    """token_linear on bf16 hi + lo operand pairs (vqb_token_linear_split): fp32-faithful to 2^-16 per product.
    a (T, 2 K) = [a_hi | a_lo], w (N, 2 K) = [w_hi | w_lo] (bf16_pair), bias (N,) fp32, h (T, N) fp32,
    out (T, 2 N) bf16 pair of gelu(x) (erf form), or of x itself with out_gelu = False."""
    for t, name, dt in ((a, "a", torch.bfloat16), (w, "w", torch.bfloat16), (bias, "bias", torch.float32)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous CUDA {dt} tensor (no CPU fallback)")
    t_rows, k2 = a.shape
    n = w.shape[0]
    if w.shape[1] != k2 or k2 % 2 or bias.numel() != n:
        raise RuntimeError("token_linear_split: shape mismatch")
    if mode == 0 and out is None:
        out = torch.empty((t_rows, 2 * n), dtype=torch.bfloat16, device=a.device)
    if mode != 0 and (h is None or h.dtype != torch.float32 or not h.is_contiguous() or tuple(h.shape) != (t_rows, n)):
        raise RuntimeError("token_linear_split modes 1 and 2 need a contiguous fp32 h of shape (T, N)")
    if out is not None and (out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape) != (t_rows, 2 * n)):
        raise RuntimeError("token_linear_split: out must be a contiguous bf16 (T, 2 N) tensor")
    if t_rows == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.vqb_token_linear_split(a.device.index, a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                        h.data_ptr() if h is not None else None,
                                        out.data_ptr() if out is not None else None,
                                        t_rows, k2 // 2, n, int(mode), int(bool(out_gelu)),
                                        torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "vqb_token_linear_split")
    return out


def token_conv(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, h: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None, mode: int = 0, taps: int = 3, tokens_per_cycle: int = 16,
               out_gelu: bool = True) -> Optional[torch.Tensor]:
    """token_linear with three taps along the positions of a cycle (vqb_token_conv: the decoder's Conv1d(k=3, pad=1),
    model/vq_vae_patch_embedd.py:60-74,142-147): x[t] = sum_tap w[:, tap*K:(tap+1)*K] a[t + tap - 1] + bias with zeros
    outside the token's cycle, then the epilogue of `mode`; out_gelu=False stores bf16(x) / bf16(h) without the GELU.
    a (T, K) bf16, w (N, taps*K) bf16, bias (N,) fp32, h (T, N) fp32, out (T, N) bf16 -- all CUDA, contiguous."""
    for t, name, dt in ((a, "a", torch.bfloat16), (w, "w", torch.bfloat16), (bias, "bias", torch.float32)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous CUDA {dt} tensor (no CPU fallback)")
    t_rows, k = a.shape
    n = w.shape[0]
    if w.shape[1] != taps * k or bias.numel() != n:
        raise RuntimeError("token_conv: shape mismatch")
    if mode == 0 and out is None:
        out = torch.empty((t_rows, n), dtype=torch.bfloat16, device=a.device)
    if mode != 0 and (h is None or h.dtype != torch.float32 or not h.is_contiguous() or tuple(h.shape) != (t_rows, n)):
        raise RuntimeError("token_conv modes 1 and 2 need a contiguous fp32 h of shape (T, N)")
    if out is not None and (out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape) != (t_rows, n)):
        raise RuntimeError("token_conv: out must be a contiguous bf16 (T, N) tensor")
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.vqb_token_conv(a.device.index, a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                h.data_ptr() if h is not None else None, out.data_ptr() if out is not None else None,
                                t_rows, k, n, int(mode), int(taps), int(tokens_per_cycle), 1 if out_gelu else 0,
                                torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "vqb_token_conv")
    return out


def token_conv_split(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, h: Optional[torch.Tensor] = None,
                     out: Optional[torch.Tensor] = None, mode: int = 0, taps: int = 3, tokens_per_cycle: int = 16,
                     out_gelu: bool = True) -> Optional[torch.Tensor]:
    """token_conv on bf16 hi + lo operand pairs (vqb_token_conv_split; fp32-faithful like token_linear_split).
    a (T, 2 K) = [a_hi | a_lo], w (N, taps * 2 K) = per tap [w_hi | w_lo] (conv_pair), bias (N,) fp32, h (T, N) fp32,
    out (T, 2 N) bf16 pair."""
    for t, name, dt in ((a, "a", torch.bfloat16), (w, "w", torch.bfloat16), (bias, "bias", torch.float32)):
        if not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous CUDA {dt} tensor (no CPU fallback)")
    t_rows, k2 = a.shape
    n = w.shape[0]
    if w.shape[1] != taps * k2 or k2 % 2 or bias.numel() != n:
        raise RuntimeError("token_conv_split: shape mismatch")
    if mode == 0 and out is None:
        out = torch.empty((t_rows, 2 * n), dtype=torch.bfloat16, device=a.device)
    if mode != 0 and (h is None or h.dtype != torch.float32 or not h.is_contiguous() or tuple(h.shape) != (t_rows, n)):
        raise RuntimeError("token_conv_split modes 1 and 2 need a contiguous fp32 h of shape (T, N)")
    if out is not None and (out.dtype != torch.bfloat16 or not out.is_contiguous() or tuple(out.shape) != (t_rows, 2 * n)):
        raise RuntimeError("token_conv_split: out must be a contiguous bf16 (T, 2 N) tensor")
    if t_rows == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.vqb_token_conv_split(a.device.index, a.data_ptr(), w.data_ptr(), bias.data_ptr(),
                                      h.data_ptr() if h is not None else None, out.data_ptr() if out is not None else None,
                                      t_rows, k2 // 2, n, int(mode), int(taps), int(tokens_per_cycle), 1 if out_gelu else 0,
                                      torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "vqb_token_conv_split")
    return out


def conv_pair(w: torch.Tensor) -> torch.Tensor:
    """Conv1d weight (N, K, taps) fp32 -> (N, taps * 2 K) bf16: per tap [bf16(w) | bf16(w - bf16(w))], the weight operand of
    token_conv_split."""
    return torch.cat([bf16_pair(w[:, :, t].contiguous()) for t in range(w.shape[2])], dim=1).contiguous()


def token_out_proj_pair(a: torch.Tensor, w: torch.Tensor, bias: float, group: int) -> torch.Tensor:
    """token_out_proj on a bf16 pair (vqb_token_out_proj_pair): a (T, 2 * group * H) = [hi | lo] -> (T * group, P) fp32."""
    if not a.is_cuda or a.dtype != torch.bfloat16 or not a.is_contiguous() or a.dim() != 2:
        raise RuntimeError("token_out_proj_pair: a must be a contiguous CUDA bf16 (T, 2 * group * H) tensor (no CPU fallback)")
    hdim = w.shape[1]
    if not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous() or w.dim() != 2 or a.shape[1] != 2 * group * hdim:
        raise RuntimeError("token_out_proj_pair: w must be a contiguous CUDA fp32 (P, H) tensor with 2 * group * H = a.shape[1]")
    out = torch.empty((a.shape[0] * group, w.shape[0]), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.vqb_token_out_proj_pair(a.device.index, a.data_ptr(), w.data_ptr(), float(bias), out.data_ptr(), a.shape[0],
                                         int(group), hdim, w.shape[0], torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "vqb_token_out_proj_pair")
    return out


def token_out_proj(a: torch.Tensor, w: torch.Tensor, bias: float) -> torch.Tensor:
    """out[r, j] = sum_c a[r, c] w[j, c] + bias (vqb_token_out_proj: PatchEmbeddingInverse's last ConvTranspose1d,
    model/vq_vae_patch_embedd.py:24-29).  a (R, H) bf16, w (P, H) fp32 -> (R, P) fp32."""
    if not a.is_cuda or a.dtype != torch.bfloat16 or not a.is_contiguous() or a.dim() != 2:
        raise RuntimeError("token_out_proj: a must be a contiguous CUDA bf16 (R, H) tensor (no CPU fallback)")
    if not w.is_cuda or w.dtype != torch.float32 or not w.is_contiguous() or w.dim() != 2 or w.shape[1] != a.shape[1]:
        raise RuntimeError("token_out_proj: w must be a contiguous CUDA fp32 (P, H) tensor")
    out = torch.empty((a.shape[0], w.shape[0]), dtype=torch.float32, device=a.device)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.vqb_token_out_proj(a.device.index, a.data_ptr(), w.data_ptr(), float(bias), out.data_ptr(), a.shape[0],
                                    a.shape[1], w.shape[0], torch.cuda.current_stream(a.device).cuda_stream)
    _lib.check(rc, "vqb_token_out_proj")
    return out


def token_bias_gelu(h: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """h += bias (in place, fp32 (T, N)); returns bf16(gelu(h)) -- vqb_token_bias_gelu."""
    if not h.is_cuda or h.dtype != torch.float32 or not h.is_contiguous() or h.dim() != 2:
        raise RuntimeError("token_bias_gelu: h must be a contiguous CUDA fp32 (T, N) tensor (no CPU fallback)")
    if not bias.is_cuda or bias.dtype != torch.float32 or bias.numel() != h.shape[1]:
        raise RuntimeError("token_bias_gelu: bias must be a CUDA fp32 (N,) tensor")
    if out is None:
        out = torch.empty(h.shape, dtype=torch.bfloat16, device=h.device)
    lib = _lib.load()
    with torch.cuda.device(h.device):
        rc = lib.vqb_token_bias_gelu(h.device.index, h.data_ptr(), bias.contiguous().data_ptr(), out.data_ptr(),
                                     h.shape[0], h.shape[1], torch.cuda.current_stream(h.device).cuda_stream)
    _lib.check(rc, "vqb_token_bias_gelu")
    return out


def patch_embed(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, patch: int, want_act: bool = True):
    """PatchEmbedding.forward (model/vq_vae_patch_embedd.py:13-17) on token rows, fused with the first residual block's
    leading GELU: x (B, L, C) fp32 contiguous -> (h (B*T, H) fp32, a = bf16(gelu(h)) or None) -- vqb_patch_embed.
    weight: the Conv1d weight (H, 1, P) or its (H, P) slice; tokens come out channel-major like the reference's."""
    _require_cuda_fp32(x, "x")
    _require_cuda_fp32(weight, "patch_embed weight")
    _require_cuda_fp32(bias, "patch_embed bias")
    if x.dim() != 3 or not x.is_contiguous():
        raise RuntimeError("patch_embed: x must be a contiguous (B, L, C) tensor")
    b, l, c = x.shape
    w = weight.detach().reshape(weight.shape[0], -1).contiguous()
    hdim = w.shape[0]
    if w.shape[1] != patch or l % patch:
        raise RuntimeError("patch_embed: weight / patch size / sequence length mismatch")
    n_tokens = b * (l // patch) * c
    lib = _lib.load()
    with torch.cuda.device(x.device):
        h = torch.empty((n_tokens, hdim), dtype=torch.float32, device=x.device)
        a = torch.empty((n_tokens, hdim), dtype=torch.bfloat16, device=x.device) if want_act else None
        rc = lib.vqb_patch_embed(x.device.index, x.data_ptr(), b, l, c, int(patch), w.data_ptr(),
                                 bias.detach().contiguous().data_ptr(), h.data_ptr(),
                                 a.data_ptr() if a is not None else None, hdim,
                                 torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "vqb_patch_embed")
    return h, a


_chain_scratch = {}


def projection_rows(wp: torch.Tensor) -> torch.Tensor:
    """The 128 weight rows vqb_encoder_chain expects behind the layers for a fused projection: rows [0, D) = bf16(Wp),
    rows [64, 64 + D) = bf16(Wp - bf16(Wp)) (together Wp to 2^-17 relative), zeros elsewhere.  wp: (D, H) fp32, D <= 64."""
    d, hdim = wp.shape
    rows = torch.zeros(128, hdim, dtype=torch.bfloat16, device=wp.device)
    hi = wp.to(torch.bfloat16)
    rows[:d] = hi
    rows[64:64 + d] = (wp - hi.float()).to(torch.bfloat16)
    return rows


def patch_rows(wpe: torch.Tensor, hidden: int) -> torch.Tensor:
    """The `hidden` weight rows vqb_encoder_chain expects at the end of its weight stack for a fused patch embedding:
    columns [0, 32) = bf16(Wpe), [32, 64) = bf16(Wpe - bf16(Wpe)), zero elsewhere.  wpe: (hidden, P) fp32, P <= 32."""
    hdim, p = wpe.shape
    rows = torch.zeros(hdim, hidden, dtype=torch.bfloat16, device=wpe.device)
    hi = wpe.to(torch.bfloat16)
    rows[:, :p] = hi
    rows[:, 32:32 + p] = (wpe - hi.float()).to(torch.bfloat16)
    return rows


def patch_split(x: torch.Tensor, patch: int) -> torch.Tensor:
    """x (B, L, C) fp32 contiguous -> (B * T, 64) bf16 operand of the fused patch embedding (vqb_patch_split): per token
    the patch's samples as a bf16 hi + lo pair, 32 + 32 columns, tokens in the reference's channel-major order."""
    _require_cuda_fp32(x, "x")
    if x.dim() != 3 or not x.is_contiguous():
        raise RuntimeError("patch_split: x must be a contiguous (B, L, C) tensor")
    b, l, c = x.shape
    if patch > 32 or l % patch:
        raise RuntimeError("patch_split: patch size must divide the sequence length and be <= 32")
    lib = _lib.load()
    with torch.cuda.device(x.device):
        out = torch.empty((b * (l // patch) * c, 64), dtype=torch.bfloat16, device=x.device)
        rc = lib.vqb_patch_split(x.device.index, x.data_ptr(), b, l, c, int(patch), out.data_ptr(),
                                 torch.cuda.current_stream(x.device).cuda_stream)
    _lib.check(rc, "vqb_patch_split")
    return out


def encoder_chain(a0: torch.Tensor, h: Optional[torch.Tensor], w: torch.Tensor, bias: torch.Tensor,
                  proj_bias: Optional[torch.Tensor] = None, pre_bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """All residual blocks of the patch encoder in one launch (vqb_encoder_chain, csrc/enc_chain.cu):
    for every block b:  h <- h + w[2b+1] gelu(w[2b] gelu(h) + bias[2b]) + bias[2b+1], h updated in place and returned.
    a0 (T, H) bf16 = bf16(gelu(h)); h (T, H) fp32; w (L, H, H) bf16 (out x in per layer); bias (L, H) fp32; H in {256, 512}.
    With proj_bias (D,) fp32 the final projection H -> D is fused: w is then 2-D, the L * H layer rows followed by
    projection_rows(Wp), and z_e (T, D) fp32 is returned (h is not written back).
    With pre_bias (H,) fp32 the patch embedding is fused too: a0 is then patch_split(x) (T, 64), w ends with
    patch_rows(Wpe, H), and h is not read (it may be None when proj_bias is given as well)."""
    hidden = bias.shape[1] if bias.dim() == 2 else -1
    checks = [(a0, "a0", torch.bfloat16), (w, "w", torch.bfloat16), (bias, "bias", torch.float32)]
    if h is not None:
        checks.append((h, "h", torch.float32))
    if proj_bias is not None:
        checks.append((proj_bias, "proj_bias", torch.float32))
    if pre_bias is not None:
        checks.append((pre_bias, "pre_bias", torch.float32))
    for t, name, dt in checks:
        if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != dt or not t.is_contiguous():
            raise RuntimeError(f"encoder_chain: {name} must be a contiguous CUDA {dt} tensor (no CPU fallback)")
    layers = bias.shape[0]
    extra = (128 if proj_bias is not None else 0) + (hidden if pre_bias is not None else 0)
    n_tokens = a0.shape[0]
    ok = (a0.dim() == 2 and a0.shape[1] == (64 if pre_bias is not None else hidden)
          and w.numel() == (layers * hidden + extra) * hidden and (extra == 0 or w.dim() == 2)
          and (h is None or tuple(h.shape) == (n_tokens, hidden))
          and (h is not None or (proj_bias is not None and pre_bias is not None))
          and (pre_bias is None or pre_bias.numel() == hidden))
    if not ok:
        raise RuntimeError("encoder_chain: shape mismatch")
    lib = _lib.load()
    dev = a0.device
    with torch.cuda.device(dev):
        key = (dev.index, hidden, torch.cuda.current_stream(dev).cuda_stream)
        scratch = _chain_scratch.get(key)
        if scratch is None:
            nbytes = lib.vqb_encoder_chain_scratch_bytes(dev.index, hidden)
            if nbytes == 0:
                raise RuntimeError("encoder_chain: unsupported device / hidden size")
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _chain_scratch[key] = scratch
        z_e = None
        if proj_bias is not None:
            z_e = torch.empty((n_tokens, proj_bias.numel()), dtype=torch.float32, device=dev)
        rc = lib.vqb_encoder_chain(dev.index, a0.data_ptr(), h.data_ptr() if h is not None else None, w.data_ptr(),
                                   bias.data_ptr(), n_tokens, hidden, layers, scratch.data_ptr(), scratch.numel(),
                                   proj_bias.data_ptr() if proj_bias is not None else None,
                                   z_e.data_ptr() if z_e is not None else None,
                                   proj_bias.numel() if proj_bias is not None else 0,
                                   pre_bias.data_ptr() if pre_bias is not None else None,
                                   torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "vqb_encoder_chain")
    return z_e if z_e is not None else h
