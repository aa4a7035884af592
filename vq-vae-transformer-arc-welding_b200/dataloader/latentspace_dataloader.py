"""The encode calls of the reference's offline tokeniser, on the B200 path.

Mirror of the hot-path part of ``dataloader/latentspace_dataloader.py``:
  get_latent_space / get_latent_space_IDs                       (:144-161)
  create_latent_space_dataset_VQ_VAE[_IDs|_autoreggressive]      (:171-263)
with the same method names, arguments and returned numpy arrays (shapes, dtypes, order),
so ``LatentSpaceDataLoader`` can delegate to it.  Everything around it in that file (CSV
loading, window bookkeeping, pickle cache, wandb artifacts) is data plumbing outside the hot
path and is not rebuilt.

What changes underneath:
  * the reference encodes a batch of windows one 200-sample cycle at a time: per cycle one
    H2D copy, one launch storm and one blocking ``.cpu()`` (:231-235), then ``np.append``s
    (O(n^2) host copying, :238).  Here all cycles of a batch go through the encoder and the
    fused VQ kernel in ONE call (cycles are independent: every op before the decoder is
    per token), results land in preallocated arrays, and only ids cross PCIe for the id tasks;
  * overlapping windows repeat cycles: with `dedupe = True` every distinct cycle of a batch is encoded once and its
    ids are scattered to all its windows (encode_unique; identical arrays, up to 20x less encoder work); with
    `dedupe = "dataset"` once per data set (CycleIdCache); ids leave through pinned staging buffers on a copy stream
    (_AsyncHostWriter), and save_latent_dataset writes the reference's own pickle cache;
  * multi-GPU: batches are sharded across ranks (one process per GPU), the codebook and
    encoder weights are replicated, the only collectives are an optional gather of the ids and
    one K-element all-reduce of the code-usage histogram (bulk_encode_ids / gather_sharded).
"""
from __future__ import annotations

import os
import pickle
import queue
import threading
from typing import Callable, Iterable, List, Optional, Tuple

import numpy as np
import torch
from torch import distributed as dist


# ---------------------------------------------------------------------------------------
# NVTX ranges around the host-side steps: visible in Nsight Systems timelines, free otherwise
# ---------------------------------------------------------------------------------------
import contextlib


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


# ---------------------------------------------------------------------------------------
# duplicate cycles
# ---------------------------------------------------------------------------------------
# The reference's windows overlap: with a stride of one cycle, a batch of 512 windows x 20 cycles holds ~531 distinct
# cycles, and the loops above :171-263 encode every one of the 10 240.  Every op before the decoder is per cycle, so the
# ids of a cycle do not depend on the window it is seen in: encode each distinct cycle once and scatter.  Distinct is
# decided on the BIT PATTERN of the samples: two 64-bit multiplicative hashes group the rows, and every row is then
# compared word for word with its group's representative, so a hash collision costs a second pass, never a wrong id.
_HASH_SEEDS = (0x9E3779B97F4A7C15, 0xC2B2AE3D27D4EB4F)


_weight_cache: dict = {}


def _hash_weights(width: int, device, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed & 0x7FFFFFFF)
    w = torch.randint(-(1 << 62), 1 << 62, (width,), generator=g, dtype=torch.int64)
    return (w | 1).to(device)                 # odd multipliers; int64 arithmetic wraps


def _hash_weight_pair(width: int, device) -> torch.Tensor:
    """(2, width) int64: the multipliers of both hashes on `device`, built once per (width, device)."""
    key = (width, str(device), _HASH_SEEDS)
    w = _weight_cache.get(key)
    if w is None:
        w = torch.stack([_hash_weights(width, device, s) for s in _HASH_SEEDS]).contiguous()
        _weight_cache[key] = w
    return w


def cycle_fingerprints(rows: torch.Tensor) -> torch.Tensor:
    """(n, ...) float32 -> (n, 2) int64: two multiplicative hashes of the rows' bit patterns, key_s = sum_j uint32(word_j) *
    w_s[j] (wrapping; the 32-bit words zero-extended).  CUDA rows: one pass of vqb_row_keys (the rows are read once); CPU rows (host-logic tests): the
    same sums as torch ops."""
    n = rows.shape[0]
    if n == 0:
        return torch.empty((0, 2), dtype=torch.int64, device=rows.device)
    if rows.is_cuda:
        from .. import ops
        bits = rows.reshape(n, -1)
        return ops.row_keys(bits, _hash_weight_pair(bits.shape[1], rows.device))
    wide = rows.reshape(n, -1).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return torch.stack([(wide * _hash_weights(wide.shape[1], rows.device, s)).sum(dim=1) for s in _HASH_SEEDS], dim=1)


def dedupe_rows(rows: torch.Tensor, return_keys: bool = False):
    """rows (n, ...) float32 -> (rep, inverse): rep (u,) int64 indices of one representative per group of bit-identical
    rows (the first occurrence), inverse (n,) int64 with rows[i] bit-identical to rows[rep[inverse[i]]].  Rows that a
    hash collision put into a foreign group come back as their own representatives.

    CUDA rows: fingerprints (vqb_row_keys) -> first row with the same fingerprints (vqb_dedupe_first: a hash table in
    device memory) -> word-for-word check against that row -> representatives in ascending row order; ONE host
    synchronisation (the number of representatives).  CPU rows (host-logic tests): the same grouping with torch.unique.
    return_keys: also the (n, 2) fingerprints of all rows (the data-set cache keys its entries by them)."""
    n = rows.shape[0]
    if n == 0:
        e = torch.empty(0, dtype=torch.int64, device=rows.device)
        return (e, e, torch.empty((0, 2), dtype=torch.int64, device=rows.device)) if return_keys else (e, e)
    bits = rows.reshape(n, -1).contiguous().view(torch.int32)
    keys = cycle_fingerprints(rows)
    ar = torch.arange(n, device=rows.device)
    if rows.is_cuda:
        from .. import ops
        first = ops.dedupe_first(keys)
    else:
        _, group = torch.unique(keys, dim=0, return_inverse=True)
        u = int(group.max().item()) + 1
        first = torch.full((u,), n, dtype=torch.int64, device=rows.device).scatter_reduce_(0, group, ar, reduce="amin")[group]
    same = (bits == bits[first]).all(dim=1)                   # word-for-word check against the row it was grouped with
    first = torch.where(same, first, ar)                      # collided rows: each becomes its own group
    is_rep = first == ar
    rank = torch.cumsum(is_rep, dim=0) - 1                    # position of a representative among the representatives
    rep = is_rep.nonzero().view(-1)                           # ascending row order (the one synchronisation)
    return (rep, rank[first], keys) if return_keys else (rep, rank[first])


def encode_unique(encode_fn: Callable[[torch.Tensor], torch.Tensor], cycles: torch.Tensor, pre=None) -> torch.Tensor:
    """encode_fn(cycles) for a per-cycle encode_fn ((b, ...) -> (b, T)), evaluated once per distinct cycle.
    pre: dedupe_rows(cycles) if the caller has it already (the bulk loops compute it one batch ahead)."""
    rep, inverse = dedupe_rows(cycles) if pre is None else pre
    if rep.numel() == cycles.shape[0]:
        return encode_fn(cycles)
    return encode_fn(cycles[rep])[inverse]


class CycleIdCache:
    """Whole-data-set de-duplication: the ids of every distinct cycle seen so far, keyed by
    its 128-bit fingerprint.  A data set of overlapping windows (stride of one cycle, 20 cycles per window) repeats every
    cycle up to 20 times ACROSS batches too; with the cache each distinct cycle passes through the encoder once per data
    set.  Within a batch, identity is still decided word for word (dedupe_rows); across batches a hit is verified word for
    word against the stored representative while the representatives fit in `max_rep_bytes` of device memory, and rests
    on the fingerprint alone beyond that (two independent 64-bit hashes: ~n^2 / 2^128 expected false matches).
    Device-resident: a sorted key column for torch.searchsorted, re-sorted when a batch adds new cycles."""

    def __init__(self, device, max_rep_bytes: int = 4 << 30):
        self.device = torch.device(device)
        self.k1 = torch.empty(0, dtype=torch.int64, device=self.device)      # sorted
        self.k2 = torch.empty(0, dtype=torch.int64, device=self.device)
        self.slot = torch.empty(0, dtype=torch.int64, device=self.device)    # row of ids / reps for the sorted position
        self._ids: Optional[torch.Tensor] = None                             # (capacity, T): rows [0, m) are used
        self._reps: Optional[torch.Tensor] = None                            # (capacity, words) int32 while they fit
        self._keep_reps = True
        self.max_rep_bytes = max_rep_bytes
        self.hits = 0
        self.misses = 0

    def __len__(self) -> int:
        return int(self.k1.numel())

    @property
    def ids(self) -> Optional[torch.Tensor]:
        """(m, T): the ids of the cached cycles."""
        return None if self._ids is None else self._ids[: len(self)]

    @property
    def reps(self) -> Optional[torch.Tensor]:
        """(m, words) int32: the cached cycles themselves, while they fit max_rep_bytes."""
        return None if self._reps is None else self._reps[: len(self)]

    @staticmethod
    def _append(buf: Optional[torch.Tensor], used: int, rows: torch.Tensor, cap_rows: Optional[int] = None) -> torch.Tensor:
        """rows appended behind buf[:used]; the buffer doubles when it is full (amortised: no copy of the whole store
        per call), up to cap_rows."""
        need = used + rows.shape[0]
        if buf is None or buf.shape[0] < need:
            cap = max(need, 2 * (0 if buf is None else buf.shape[0]), 1024)
            if cap_rows is not None:
                cap = max(min(cap, cap_rows), need)
            grown = torch.empty((cap,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
            if used:
                grown[:used] = buf[:used]
            buf = grown
        buf[used:need] = rows
        return buf

    def lookup(self, cycles: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
        """(n,) int64: row of the cached ids for every cycle, -1 where the cycle is new."""
        n = cycles.shape[0]
        out = torch.full((n,), -1, dtype=torch.int64, device=self.device)
        m = len(self)
        if m == 0 or n == 0:
            return out
        pos = torch.searchsorted(self.k1, keys[:, 0].contiguous()).clamp_(max=m - 1)
        hit = (self.k1[pos] == keys[:, 0]) & (self.k2[pos] == keys[:, 1])
        rows = self.slot[pos]
        if self.reps is not None:
            bits = cycles.reshape(n, -1).contiguous().view(torch.int32)
            hit &= (self.reps[rows] == bits).all(dim=1)                      # word-for-word check
        out[hit] = rows[hit]
        return out

    def insert(self, cycles: torch.Tensor, keys: torch.Tensor, ids: torch.Tensor) -> None:
        """Remember the ids of `cycles` (distinct, not in the cache)."""
        if cycles.shape[0] == 0:
            return
        m = len(self)
        self._ids = self._append(self._ids, m, ids)
        bits = cycles.reshape(cycles.shape[0], -1).contiguous().view(torch.int32)
        if self._keep_reps:
            row_bytes = bits.shape[1] * 4
            if (m + bits.shape[0]) * row_bytes <= self.max_rep_bytes:
                self._reps = self._append(self._reps, m, bits, cap_rows=self.max_rep_bytes // max(row_bytes, 1))
            else:
                self._reps, self._keep_reps = None, False                    # from here on: fingerprints only
        k1 = torch.cat([self.k1, keys[:, 0]])
        order = torch.argsort(k1)
        self.k1 = k1[order]
        self.k2 = torch.cat([self.k2, keys[:, 1]])[order]
        self.slot = torch.cat([self.slot, m + torch.arange(cycles.shape[0], device=self.device)])[order]

    def encode(self, encode_fn: Callable[[torch.Tensor], torch.Tensor], cycles: torch.Tensor, pre=None) -> torch.Tensor:
        """ids of `cycles` ((n, ...) -> (n, T)): distinct cycles of the batch first (word for word), then the cache,
        then the encoder for what is left.  pre: dedupe_rows(cycles, return_keys=True) if the caller has it already."""
        rep, inverse, all_keys = dedupe_rows(cycles, return_keys=True) if pre is None else pre
        uniq = cycles[rep]
        keys = all_keys[rep]
        rows = self.lookup(uniq, keys)
        new = (rows < 0).nonzero().view(-1)
        self.hits += int(cycles.shape[0] - new.numel())
        self.misses += int(new.numel())
        if new.numel():
            new_ids = encode_fn(uniq[new]).view(new.numel(), -1)
            base = len(self)
            self.insert(uniq[new], keys[new], new_ids)
            rows = rows.clone()
            rows[new] = base + torch.arange(new.numel(), device=self.device)
        return self.ids[rows][inverse]


class _AsyncHostWriter:
    """Device results -> ONE host array through a small pool of pinned staging buffers on a copy stream: the D2H copy of
    batch i overlaps the encoder work of batch i + 1 (the reference blocks on `.cpu()` once per cycle slice, :231-235, and
    `np.append`s, :238).  A retired batch is copied once, straight into its rows of the output array (converted to
    `out_dtype` on the way); the array is sized from `rows_hint` (the loops set it to the loader's length x its first batch: exact
    for the equal batches a DataLoader yields) and doubles if that was too small."""

    def __init__(self, device, depth: int = 3, rows_hint: Optional[int] = None, out_dtype=None):
        self.cuda = torch.device(device).type == "cuda"
        self.depth = depth
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.pending: List[Tuple[torch.Tensor, Optional["torch.cuda.Event"], tuple]] = []
        self.free: List[torch.Tensor] = []
        self.rows_hint, self.out_dtype = rows_hint, out_dtype
        self.arr: Optional[np.ndarray] = None
        self.n = 0

    def _append(self, a: np.ndarray) -> None:
        b = a.shape[0]
        if self.arr is None:
            cap = max(self.rows_hint or 0, b, 1)
            self.arr = np.empty((cap,) + a.shape[1:], dtype=self.out_dtype or a.dtype)
        elif self.n + b > self.arr.shape[0]:
            grown = np.empty((max(self.n + b, 2 * self.arr.shape[0]),) + self.arr.shape[1:], dtype=self.arr.dtype)
            grown[: self.n] = self.arr[: self.n]
            self.arr = grown
        self.arr[self.n: self.n + b] = a
        self.n += b

    def _retire(self) -> None:
        buf, ev, shape = self.pending.pop(0)
        if ev is not None:
            ev.synchronize()
        n = int(np.prod(shape))
        self._append(buf[:n].view(shape).numpy())
        self.free.append(buf)

    def put(self, t: torch.Tensor) -> None:
        if not self.cuda:
            self._append(t.cpu().numpy())
            return
        # retire what has arrived (its rows are written while the GPU works on the next batches, not at the end), and
        # whatever must go to keep the pool at `depth` buffers
        while self.pending and (len(self.pending) >= self.depth or self.pending[0][1].query()):
            self._retire()
        n = t.numel()
        k = next((i for i, b in enumerate(self.free) if b.numel() >= n and b.dtype == t.dtype), None)
        if k is not None:
            buf = self.free.pop(k)         # (by position: list.remove would compare tensors element-wise)
        else:
            buf = torch.empty(max(n, 1), dtype=t.dtype, pin_memory=True)
        self.stream.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(self.stream):
            buf[:n].copy_(t.reshape(-1), non_blocking=True)
            t.record_stream(self.stream)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending.append((buf, ev, tuple(t.shape)))

    def finish(self) -> Optional[np.ndarray]:
        """All rows written so far (None if there were none)."""
        while self.pending:
            self._retire()
        return None if self.arr is None else self.arr[: self.n]


def _n_batches(loader) -> Optional[int]:
    try:
        return len(loader)
    except TypeError:
        return None


class _DevicePrefetcher:
    """Loader batches -> their cycles on the device, in GROUPS of up to `group_cycles` cycles, one group ahead.

    The host->device copy of every loader batch runs on a copy stream, straight into its rows of the group's device slot,
    while the encoder works on the previous group (the reference copies one cycle slice at a time on the compute stream
    and then blocks on it, :231-232).  A pageable loader tensor (or one whose windows are longer than seq_len cycles)
    first goes through one of three pinned staging buffers -- `copy_` on CPU tensors of this size runs on all host
    threads -- a pinned, exactly-sized one is copied as it is.  Three device slots rotate (being encoded / handed out /
    being filled); a slot is overwritten only after the compute stream has passed the work that read it (`consumed`
    events), a staging buffer only after its copy has finished.  The loader is walked and the copies are issued by a
    helper thread.  Grouping amortises the per-call host work of the encoder over several loader batches (a batch of 512
    windows x 20 cycles keeps the GPU busy for 1.5 ms, about what its launches cost the host) and lets the
    de-duplication see the overlap BETWEEN consecutive batches.  On a CPU device, or for batches that are on a GPU
    already, it degenerates to `.to(device)` per batch.  (A pinned batch is read by the copy engine after the loader has
    moved on: like any non_blocking copy this wants loaders that hand out a batch and leave it alone -- DataLoader does.)

    Yields (cycles_on_device (rows, window, C), items) with `items` the list of what the loader produced for the group,
    rows = sum over the items of batch size x seq_len, in loader order."""

    SLOTS = 3

    def __init__(self, loader: Iterable, device, seq_len: int, window_size: int, no_labels: bool = False,
                 group_cycles: int = 65536):
        self.loader, self.device = loader, torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:      # "cuda": the caller's current device, by number
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.seq_len, self.window = seq_len, window_size
        self.no_labels = no_labels
        self.group_cycles = max(int(group_cycles), 1)
        self.cuda = self.device.type == "cuda"
        if self.cuda:
            self.stream = torch.cuda.Stream(device=self.device)
            self.dev: List[Optional[torch.Tensor]] = [None] * self.SLOTS
            self.consumed: List[Optional["torch.cuda.Event"]] = [None] * self.SLOTS
            self.stage: List[Optional[torch.Tensor]] = [None] * self.SLOTS
            self.copied: List[Optional["torch.cuda.Event"]] = [None] * self.SLOTS
        self.n_groups = 0
        self.n_staged = 0

    def _source(self, item) -> torch.Tensor:
        x = item if self.no_labels else item[0]
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        return x[:, : self.seq_len * self.window, :]

    def _copy_in(self, src: torch.Tensor, dst: torch.Tensor) -> None:
        """One loader batch (b, seq_len * window, C) on the host -> its rows `dst` (b * seq_len, window, C) of a slot."""
        if not (src.is_contiguous() and src.is_pinned()):
            k = self.n_staged % self.SLOTS
            self.n_staged += 1
            if self.copied[k] is not None:
                self.copied[k].synchronize()               # the copy that last read this staging buffer
            n = src.numel()
            if self.stage[k] is None or self.stage[k].numel() < n or self.stage[k].dtype != src.dtype:
                self.stage[k] = torch.empty(max(n, 1), dtype=src.dtype, pin_memory=True)
            host = self.stage[k][:n].view(src.shape)
            host.copy_(src)
        else:
            k, host = -1, src
        with torch.cuda.stream(self.stream):
            dst.copy_(host.reshape(dst.shape), non_blocking=True)
            if k >= 0:
                ev = torch.cuda.Event()
                ev.record(self.stream)
                self.copied[k] = ev

    def _groups(self):
        """The loader's items in groups of at most group_cycles cycles (a single larger batch is a group of its own)."""
        group, rows = [], 0
        for item in self.loader:
            src = self._source(item)
            r = src.shape[0] * self.seq_len
            if group and (rows + r > self.group_cycles or src.is_cuda != group[0][1].is_cuda or
                          src.dtype != group[0][1].dtype or src.shape[2] != group[0][1].shape[2]):
                yield group
                group, rows = [], 0
            group.append((item, src))
            rows += r
        if group:
            yield group

    def _issue(self, group):
        items = [it for it, _ in group]
        srcs = [s for _, s in group]
        c, dtype = srcs[0].shape[2], srcs[0].dtype
        rows = [s.shape[0] * self.seq_len for s in srcs]
        if not self.cuda or srcs[0].is_cuda:               # nothing to overlap: a CPU run, or cycles that are on a GPU already
            parts = [s.reshape(r, self.window, c).to(self.device) for s, r in zip(srcs, rows)]
            return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)), items, -1, None
        k = self.n_groups % self.SLOTS
        self.n_groups += 1
        per_row = self.window * c
        total = sum(rows)
        with torch.cuda.device(self.device):
            old = self.dev[k]
            if old is None or old.numel() < total * per_row or old.dtype != dtype:
                self.dev[k] = torch.empty(max(total, min(self.group_cycles, 1 << 20)) * per_row, dtype=dtype, device=self.device)
                # (a block the allocator may be recycling from the compute stream)
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
            if self.consumed[k] is not None:
                self.stream.wait_event(self.consumed[k])   # the encoder work that read this slot three groups ago
            out = self.dev[k][: total * per_row].view(total, self.window, c)
            lo = 0
            for s, r in zip(srcs, rows):
                self._copy_in(s, out[lo: lo + r])
                lo += r
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return out, items, k, ready

    def __iter__(self):
        if not self.cuda:
            for group in self._groups():
                cyc, items, _k, _ready = self._issue(group)
                yield cyc, items
            return
        # A helper thread walks the loader and issues the copies (ATen's host copy and the CUDA calls release the GIL), so the
        # staging of the next group overlaps the launches -- and, in the de-duplicating modes, the synchronisations -- of
        # this one on the consumer's thread.  The queue holds one issued group: one handed out + one queued + one being
        # issued = SLOTS.
        q: "queue.Queue" = queue.Queue(maxsize=self.SLOTS - 2)
        stop = threading.Event()
        end = object()

        def walk():
            try:
                torch.cuda.set_device(self.device)
                for group in self._groups():
                    if stop.is_set():
                        break
                    q.put(self._issue(group))
            except BaseException as e:           # handed to the consumer, raised there
                q.put(e)
            finally:
                q.put(end)

        worker = threading.Thread(target=walk, name="vqb200-prefetch", daemon=True)
        worker.start()
        try:
            while True:
                got = q.get()
                if got is end:
                    break
                if isinstance(got, BaseException):
                    raise got
                cyc, items, k, ready = got
                here = torch.cuda.current_stream(self.device)
                if k >= 0:
                    here.wait_event(ready)
                    cyc.record_stream(here)
                yield cyc, items
                if k >= 0:
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(self.device))
                    self.consumed[k] = ev        # set before the next q.get() lets the helper move on to this slot
        finally:
            stop.set()
            while worker.is_alive():             # a consumer that stops early: let the helper run into `stop`
                try:
                    q.get(timeout=0.05)
                except queue.Empty:
                    pass
            worker.join()


def latent_dataset_name(task: str, model_name: str, cycle_seq_number: int, model_id: str) -> str:
    """The directory name the reference gives a latent data set (dataloader/latentspace_dataloader.py:21-26)."""
    if task in ("classification", "classification_ids"):
        return f"asimow_ls_{task}_{model_name}_cycle_{cycle_seq_number}_{model_id}"
    if task in ("autoregressive_ids", "autoregressive_ids_classification"):
        return f"{task}_cycle_{cycle_seq_number}_{model_id}"
    raise ValueError(f"task {task} not supported")


def save_latent_dataset(splits, data_directory_path: str, dataset_name: str) -> str:
    """Writes (train, val, test) -- each an (x, y) pair of numpy arrays, what `preprocessing` returns (:79-98) -- where and how
    the reference caches it: <data>/quality_prediction_data/<dataset_name>/dataset.pickle, one pickle.dump of the tuple
    (dataloader/base_dataloader.py:149,236-246, dataloader/utils.py:38-42), so that the reference's `load_dataset` /
    `get_data_loader` find a finished data set and skip their own encode loops."""
    path = os.path.join(data_directory_path, "quality_prediction_data", dataset_name)
    os.makedirs(path, exist_ok=True)
    file = os.path.join(path, "dataset.pickle")
    with open(file, "wb") as f:
        pickle.dump(tuple((np.asarray(x), np.asarray(y)) for x, y in splits), f)
    return file


class LatentSpaceEncoder:
    """Encode-side of ``LatentSpaceDataLoader`` (reference :16-39 for the constructor fields
    that matter here: the model, ``window_size`` and the device)."""

    def __init__(self, latent_space_model, window_size: int = 200, device: Optional[str] = None,
                 encoder_mode: Optional[str] = "auto"):
        """encoder_mode: "auto" (default) selects the model's fused tcgen05 encoder ("fused_bf16": bf16 operands for the
        hidden layers, fp32 accumulation and residual stream) when the model offers it -- the reference builds its latent
        data sets under torch.set_float32_matmul_precision('medium') (train_transformer_mtasks.py:245), i.e. with bf16
        matmuls too; "fused_fp32" is the fp32-faithful form of the same kernels (bf16 hi + lo operand pairs, ~3.5x slower,
        ids equal to the fp32 layers' except on ~1e-5 of the tokens); "torch" keeps the fp32 PyTorch layers; None leaves
        the model's setting alone (VQVAEPatch's own default is "auto" = "fused_fp32" for inference calls)."""
        if device is None:
            # the reference hard-codes cuda:0 (:34); one process per GPU uses its own device
            device = f"cuda:{torch.cuda.current_device()}" if torch.cuda.is_available() else "cpu"
        self.device = device
        self.latent_space_model = latent_space_model.to(self.device)
        if encoder_mode is not None and hasattr(self.latent_space_model, "encoder_mode"):
            self.latent_space_model.encoder_mode = "fused_bf16" if encoder_mode == "auto" else encoder_mode
        self.window_size = window_size
        self.code_counts = None
        #: encode every distinct cycle of a batch once (overlapping windows repeat cycles; see encode_unique).  The ids
        #: are the same whenever the encoder is row-wise deterministic (eval mode; the fused encoder always is).  The
        #: code-usage histogram then counts distinct cycles.
        #: "dataset": additionally across the batches of one create_latent_space_dataset_* call (CycleIdCache).
        self.dedupe = False
        self.cycle_cache: Optional[CycleIdCache] = None
        #: the bulk loops hand the encoder up to this many cycles per call: consecutive loader batches are copied into one
        #: device buffer (_DevicePrefetcher); 1 = one call per loader batch.  Same arrays either way (every op is per cycle).
        self.group_cycles = 65536

    # ---- single encode calls (:144-161) -------------------------------------------------
    def _encode(self, x, has_patch_embed: bool):
        model = self.latent_space_model
        if has_patch_embed and hasattr(model, "encode"):
            # VQVAEPatch.encode = patch_embed -> encoder, in the model's `encoder_mode` (the fused tcgen05 layers when
            # selected); a reference model without that method takes the reference's own two calls below (:145-149)
            return model.encode(x)
        x = model.patch_embed(x) if has_patch_embed else x.permute(0, 2, 1)
        return model.encoder(x)

    def get_latent_space(self, x, has_patch_embed: bool = False):
        z_e = self._encode(x, has_patch_embed)
        _loss, z_q, _ppl, _oh, _idx = self.latent_space_model.vector_quantization(z_e)
        return z_q

    def get_latent_space_IDs(self, x, has_patch_embed: bool = False):
        z_e = self._encode(x, has_patch_embed)
        vq = self.latent_space_model.vector_quantization
        if hasattr(vq, "encode_indices"):
            return vq.encode_indices(z_e)
        return vq(z_e)[4]

    # ---- bulk loops (:171-263) ------------------------------------------------------------
    def _cycles(self, x: torch.Tensor, seq_len: int) -> torch.Tensor:
        """(B, >= seq_len*window, C) -> (B*seq_len, window, C): cycle i of window b at row b*seq_len+i."""
        b = x.shape[0]
        w = self.window_size
        return x[:, : seq_len * w, :].reshape(b * seq_len, w, x.shape[2])

    def create_latent_space_dataset_VQ_VAE(self, loader: Iterable, seq_len: int, has_patch_embed: bool = False):
        """Quantised latents per window: (n, seq_len, embedding_dim*enc_out_len) float64, labels (n,)."""
        model = self.latent_space_model
        width = int(model.embedding_dim * model.enc_out_len)
        ys = []
        writer = _AsyncHostWriter(self.device, out_dtype=np.float64)
        n_batches = _n_batches(loader)
        model.eval()
        with torch.no_grad():
            for cyc, items in _DevicePrefetcher(loader, self.device, seq_len, self.window_size, group_cycles=self.group_cycles):
                if writer.rows_hint is None and n_batches:
                    writer.rows_hint = n_batches * int(items[0][0].shape[0])
                with nvtx_range("vqb200.encode_batch"):
                    z_q = self.get_latent_space(cyc, has_patch_embed=has_patch_embed)
                writer.put(z_q.reshape(cyc.shape[0] // seq_len, seq_len, -1))
                for _x, y in items:
                    ys.append(np.asarray(y.cpu().numpy() if isinstance(y, torch.Tensor) else y, dtype=np.float64))
        new_x = writer.finish()
        if new_x is None:
            return np.empty((0, seq_len, width)), np.empty((0,))
        return new_x, np.concatenate(ys, axis=0)

    def create_latent_space_dataset_VQ_VAE_IDs(self, loader: Iterable, seq_len: int, has_patch_embed: bool = False,
                                               no_labels: bool = False):
        """Token ids per window: (n, seq_len, enc_out_len) int64, labels (n,) (zeros when no_labels)."""
        model = self.latent_space_model
        enc_out_len = int(model.enc_out_len)
        ys = []
        counts = None
        writer = _AsyncHostWriter(self.device)          # ids leave through pinned buffers on a copy stream
        n_batches = _n_batches(loader)
        cache = CycleIdCache(self.device) if self.dedupe == "dataset" else None
        self.cycle_cache = cache
        enc = lambda c: self.get_latent_space_IDs(c, has_patch_embed).view(c.shape[0], -1)
        model.eval()
        with torch.no_grad():
            for cyc, items in _DevicePrefetcher(loader, self.device, seq_len, self.window_size, no_labels=no_labels,
                                                group_cycles=self.group_cycles):
                b = cyc.shape[0] // seq_len
                if writer.rows_hint is None and n_batches:
                    writer.rows_hint = n_batches * int((items[0] if no_labels else items[0][0]).shape[0])
                model.vector_quantization.code_counts = None
                with nvtx_range("vqb200.encode_batch"):
                    if cache is not None:
                        ids = cache.encode(enc, cyc)
                    elif self.dedupe:
                        ids = encode_unique(enc, cyc)
                    else:
                        ids = self.get_latent_space_IDs(cyc, has_patch_embed)
                c = getattr(model.vector_quantization, "code_counts", None)
                if c is not None:
                    counts = c.clone() if counts is None else counts + c
                writer.put(ids.view(b, seq_len, -1))
                if not no_labels:
                    for _x, y in items:
                        ys.append(np.asarray(y.cpu().numpy() if isinstance(y, torch.Tensor) else y, dtype=np.float64))
        new_x = writer.finish()
        self.code_counts = counts
        if new_x is None:
            new_x = np.empty((0, seq_len, enc_out_len), dtype=int)
        new_y = np.zeros(new_x.shape[0]) if no_labels else (np.concatenate(ys, axis=0) if ys else np.empty((0,)))
        return new_x, new_y

    def create_latent_space_dataset_VQ_VAE_autoreggressive(self, loader: Iterable, seq_len: int = 1,
                                                          has_patch_embed: bool = False, task: str = "autoregressive_ids"):
        """Flattened ids (n, seq_len*enc_out_len) for the autoregressive task (:245-263)."""
        new_x, new_y = self.create_latent_space_dataset_VQ_VAE_IDs(
            loader, seq_len=seq_len, has_patch_embed=has_patch_embed, no_labels=(task == "autoregressive_ids"))
        return new_x.reshape((new_x.shape[0], -1)), new_y


    # ---- the same data sets from the CYCLE STREAM ---------------------------------
    def create_latent_space_dataset_from_cycles(self, cycles, labels=None, seq_len: int = 1, has_patch_embed: bool = False,
                                                kind: str = "ids", batch: int = 65536, shard: bool = False,
                                                materialize: bool = True):
        """The arrays the loops above build from the reference's windows, built from the cycles the windows are made of.

        The reference slides a window of `seq_len` cycles with a stride of ONE cycle over the (n, window, C) cycle array
        (`ASIMoWDataLoader.create_sequence_ds`, dataloader/asimow_dataloader.py:185-206: n - seq_len windows, window i =
        cycles i .. i + seq_len - 1, label y[i + seq_len]; with seq_len = 1 the cycles themselves, :173) and then encodes
        every window cycle by cycle (:225-238) -- every cycle seq_len times.  Every op before the decoder is per cycle, so
        the ids (latents) of a window are the ids (latents) of its cycles: each cycle is encoded ONCE, in `batch`-cycle
        calls through the loops above, and the windows are a sliding view of the result.  Same arrays, seq_len times
        less encoder work and host->device traffic (test: the reference's own create_sequence_ds + loop fixture).

        cycles: (n, window, C) numpy array or tensor, already scaled the way the loader's windows are (the reference's
                scaler is per channel, dataloader/utils.py:81-93, so scaling commutes with windowing); float64 is cast
                to float32 as the reference's Dataset classes do (dataloader/base_dataloader.py:29,72)
        labels: (n,) per-cycle labels or None (zeros, as the loops return for no_labels)
        kind:   "ids"      -> (n_windows, seq_len, enc_out_len) int64           (create_latent_space_dataset_VQ_VAE_IDs)
                "ar_ids"   -> (n_windows, seq_len * enc_out_len) int64          (..._VQ_VAE_autoreggressive)
                "latents"  -> (n_windows, seq_len, embedding_dim * enc_out_len) float64   (create_latent_space_dataset_VQ_VAE)
        shard:  with an initialised process group of several ranks (one process per GPU), every rank encodes its
                contiguous shard of the cycles (shard_range) and the per-cycle results are all-gathered -- 128 bytes per
                cycle for the ids -- so that every rank returns the whole data set; no other collective
        materialize: False returns the windows as a read-only sliding VIEW of the per-cycle array (same shape and values,
                no seq_len-fold copy: the reference's Dataset classes only index it, dataloader/base_dataloader.py:14-72);
                "ar_ids" flattens the windows and is always materialised
        Returns (array, labels (n_windows,) float64)."""
        if kind not in ("ids", "ar_ids", "latents"):
            raise ValueError(f"kind must be 'ids', 'ar_ids' or 'latents', got {kind!r}")
        if seq_len < 1:
            raise ValueError("seq_len must be >= 1")
        if isinstance(cycles, np.ndarray):
            cycles = torch.from_numpy(cycles)
        if cycles.dim() != 3 or cycles.shape[1] < self.window_size:
            raise ValueError(f"cycles must be (n, >= {self.window_size}, C), got {tuple(cycles.shape)}")
        if cycles.dtype != torch.float32:
            cycles = cycles.to(torch.float32)
        n = cycles.shape[0]
        n_windows = n if seq_len == 1 else max(n - seq_len, 0)
        if labels is None:
            new_y = np.zeros(n_windows)
        else:
            y = np.asarray(labels.cpu().numpy() if isinstance(labels, torch.Tensor) else labels, dtype=np.float64)
            if y.shape != (n,):
                raise ValueError(f"labels must be ({n},), got {y.shape}")
            new_y = y.copy() if seq_len == 1 else y[seq_len:].copy()
        model = self.latent_space_model
        width = int(model.enc_out_len) * (int(model.embedding_dim) if kind == "latents" else 1)
        sharded = shard and _dist_ready()
        lo, hi = shard_range(n, dist.get_rank(), dist.get_world_size()) if sharded else (0, n)
        loader = [cycles[s: min(s + batch, hi)] for s in range(lo, hi, batch)]
        keep, self.group_cycles = self.group_cycles, batch         # one encoder call per `batch` cycles
        try:
            if kind == "latents":
                per_cycle, _ = self.create_latent_space_dataset_VQ_VAE([(c, np.zeros(c.shape[0])) for c in loader], seq_len=1,
                                                                       has_patch_embed=has_patch_embed)
            else:
                per_cycle, _ = self.create_latent_space_dataset_VQ_VAE_IDs(loader, seq_len=1, has_patch_embed=has_patch_embed,
                                                                           no_labels=True)
        finally:
            self.group_cycles = keep
        per_cycle = per_cycle.reshape(hi - lo, width)              # (n, T) ids or (n, D * T) latents
        if sharded:
            per_cycle = gather_sharded(torch.from_numpy(np.ascontiguousarray(per_cycle)).to(self.device), n).cpu().numpy()
        if seq_len == 1:
            new_x = per_cycle.reshape(n, 1, -1)
        elif n_windows == 0:
            new_x = np.empty((0, seq_len, per_cycle.shape[1]), dtype=per_cycle.dtype)
        else:
            # (n - seq_len + 1, width, seq_len) sliding view -> the windows as their own array, seq_len x the per-cycle
            # result (what the reference's format asks for); torch's strided copy runs on all host threads
            if materialize or kind == "ar_ids":
                view = torch.from_numpy(np.ascontiguousarray(per_cycle)).unfold(0, seq_len, 1)
                new_x = view[:n_windows].permute(0, 2, 1).contiguous().numpy()
            else:
                new_x = np.lib.stride_tricks.sliding_window_view(per_cycle, seq_len, axis=0)[:n_windows].transpose(0, 2, 1)
        if kind == "ar_ids":
            new_x = new_x.reshape(new_x.shape[0], -1)
        return new_x, new_y


class OnTheFlyTokenizer:
    """Windows of raw cycles -> the transformer's training batch, on the GPU, per step:
    what the reference prepares offline as a pickled data set -- the id loop of dataloader/latentspace_dataloader.py:
    205-263 followed by MyLatentAutoregressiveDataset (dataloader/base_dataloader.py:74-110) -- as one call:
        (x, cond, y) = tok(windows, labels)
    windows (B, n_cycles * window, C) fp32, labels (B,) or None;  x = [start, ids...], y = [ids..., end] int64
    (B, n_cycles * enc_out_len + 1), cond = labels as int64 (zeros without labels), like the data set's __getitem__.

    start / end tokens: the reference derives them from the largest id PRESENT in its pickled data set (max_id + 1 / + 2,
    :85-90) while the model is built with num_embeddings + 2 classes (train_transformer_mtasks.py:146); a stream has no
    "whole data set", so `max_token` defaults to num_embeddings - 1 (the two agree whenever the last code is used)."""

    def __init__(self, latent_space_model, window_size: int = 200, device: Optional[str] = None,
                 encoder_mode: Optional[str] = "auto", max_token: Optional[int] = None):
        self.encoder = LatentSpaceEncoder(latent_space_model, window_size=window_size, device=device,
                                          encoder_mode=encoder_mode)
        model = self.encoder.latent_space_model
        model.eval()
        self.max_token = int(model.num_embeddings - 1 if max_token is None else max_token)
        self.start_token, self.end_token = self.max_token + 1, self.max_token + 2
        self.num_classes = self.max_token + 3

    def __call__(self, windows: torch.Tensor, labels: Optional[torch.Tensor] = None, n_cycles: Optional[int] = None):
        from .. import ops
        enc = self.encoder
        b = windows.shape[0]
        if n_cycles is None:
            n_cycles = windows.shape[1] // enc.window_size
        with torch.no_grad():
            cyc = enc._cycles(windows, n_cycles).to(enc.device, non_blocking=True)
            ids = enc.get_latent_space_IDs(cyc, has_patch_embed=True).view(b, -1)
            x, y = ops.ar_pairs(ids, self.start_token, self.end_token)
        if labels is None:
            cond = torch.zeros((b, 1), dtype=torch.long, device=x.device)
        else:
            cond = torch.as_tensor(labels).to(x.device).to(torch.long)
        return x, cond, y


# ---------------------------------------------------------------------------------------
# batch sharding across ranks (one process per GPU)
# ---------------------------------------------------------------------------------------
def shard_range(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items owned by `rank`; the first n % world_size ranks
    get one extra item, so concatenating the shards in rank order restores the input order."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def gather_sharded(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """All-gathers per-rank shards (split with shard_range along dim 0) into the full tensor
    in input order, on every rank.  Ragged shards are padded by one row at most."""
    if not _dist_ready():
        return local
    world = dist.get_world_size()
    base, extra = divmod(n_total, world)
    rows = base + (1 if extra else 0)
    padded = local.new_zeros((rows,) + tuple(local.shape[1:]))
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    out = []
    for r, part in enumerate(parts):
        lo, hi = shard_range(n_total, r, world)
        out.append(part[: hi - lo])
    return torch.cat(out, dim=0)


def reduce_counts(counts: torch.Tensor, async_op: bool = False):
    """Sum of the per-rank code-usage histograms (the one collective bulk encoding needs), in place.  async_op=True
    returns the work handle instead (None on a single rank): the 2 KB all-reduce then overlaps whatever the caller
    launches next, and `handle.wait()` orders the current stream behind it."""
    if _dist_ready():
        work = dist.all_reduce(counts, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            return work
    return None if async_op else counts


def bulk_encode_ids(encode_fn: Callable[[torch.Tensor], torch.Tensor], cycles: torch.Tensor, batch: int = 65536,
                    rank: Optional[int] = None, world_size: Optional[int] = None, gather: bool = True):
    """Tokenise `cycles` (n, window, C) with `encode_fn` (cycles -> (b, T) int64 ids), this rank
    encoding only its shard.  Returns the ids of all cycles in input order when gather=True
    (every rank), else this rank's shard and its [lo, hi) range."""
    if rank is None:
        rank = dist.get_rank() if _dist_ready() else 0
    if world_size is None:
        world_size = dist.get_world_size() if _dist_ready() else 1
    n = cycles.shape[0]
    lo, hi = shard_range(n, rank, world_size)
    outs = []
    for s in range(lo, hi, batch):
        outs.append(encode_fn(cycles[s: min(s + batch, hi)]))
    local = torch.cat(outs, dim=0) if outs else None
    if world_size > 1 and _dist_ready():
        # a rank whose shard is empty learns the id layout (tokens per cycle, dtype) from its peers
        meta = [None] * world_size
        dist.all_gather_object(meta, None if local is None else (tuple(local.shape[1:]), local.dtype,
                                                                 local.device.type))
        known = next((m for m in meta if m is not None), ((0,), torch.int64, cycles.device.type))
        if local is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if known[2] == "cuda" else torch.device("cpu")
            local = torch.empty((0,) + known[0], dtype=known[1], device=dev)
    elif local is None:
        local = torch.empty((0, 0), dtype=torch.int64, device=cycles.device)
    if gather and world_size > 1:
        return gather_sharded(local, n)
    return local if gather else (local, (lo, hi))
