"""Drop-in for the reference's ``model/vector_quantizer.py:VectorQuantizer`` (lines 59-131).

Same constructor, attributes, ``forward`` signature, returned 5-tuple, ``embedding.weight``
state-dict key and Lightning base class -- but ``forward`` is one fused CUDA kernel
(libvqb200.so) instead of five (N, K) temporaries, two GEMMs and a CPU-built one-hot.
There is no CPU path: calling it with CPU tensors raises.
"""
from __future__ import annotations

import torch
from torch import distributed as dist
from torch import nn

from .. import ops
from .._lightning_compat import LightningModule


class VectorQuantizer(LightningModule):
    """Discretisation bottleneck of the VQ-VAE.

    Args (model/vector_quantizer.py:68):
        n_e:   number of codebook entries
        e_dim: dimension of one entry
        beta:  weight of the codebook term, loss = mean((sg[e]-z)^2) + beta*mean((e-sg[z])^2)
               (the reference's weighting, :107-108 -- reversed w.r.t. the VQ-VAE paper)
    Extra keyword arguments (not in the reference, defaults keep its behaviour):
        one_hot: "dense" materialises the (N, n_e) fp32 ``min_encodings`` like the reference;
                 "none" returns None in its place (4*n_e bytes per vector saved; no caller
                 in the reference reads it).
        path:    "auto" | "fma" | "tc" kernel selection (see include/vqb200.h).
    """

    def __init__(self, n_e, e_dim, beta, one_hot: str = "dense", path: str = "auto"):
        super().__init__()
        if one_hot not in ("dense", "none"):
            raise ValueError(f"one_hot must be 'dense' or 'none', got {one_hot!r}")
        if path not in ops.PATHS:
            raise ValueError(f"path must be one of {sorted(ops.PATHS)}, got {path!r}")
        self.n_e = n_e
        self.e_dim = e_dim
        self.beta = beta
        self.one_hot = one_hot
        self.path = path
        self.embedding = nn.Embedding(self.n_e, self.e_dim)
        self.embedding.weight.data.uniform_(-1.0 / self.n_e, 1.0 / self.n_e)   # :74
        # code-usage histogram of the last forward (int64, K): a plain attribute, so the state
        # dict keeps its single key `embedding.weight` and DDP has no extra buffer to broadcast
        self.code_counts = None

    def forward(self, z, need_one_hot=None):
        """z: fp32 CUDA tensor whose trailing elements group into vectors of ``e_dim``
        (any shape, may be a non-contiguous view; read in place).

        Returns (loss, z_q, perplexity, min_encodings, min_encoding_indices), :119.
        need_one_hot: False = return None for min_encodings in this call whatever ``one_hot`` says (a caller that
        drops it, like VQVAEPatch.forward -- model/vq_vae_patch_embedd.py:161 --, saves the 4 * n_e bytes per vector
        of its write); None = as configured.
        """
        loss, z_q, perplexity, indices, counts = ops.VQStraightThrough.apply(
            z, self.embedding.weight, self.beta, self.path)
        self.code_counts = counts
        dense = self.one_hot == "dense" if need_one_hot is None else bool(need_one_hot)
        min_encodings = ops.one_hot(indices, self.n_e) if dense else None
        return loss, z_q, perplexity, min_encodings, indices

    def encode_indices(self, z):
        """min_encoding_indices only (what dataloader/latentspace_dataloader.py:160-161 keeps):
        no z_q write, no one-hot, no autograd graph."""
        with torch.no_grad():
            _, _, _, indices, counts = ops.forward(z, self.embedding.weight, self.beta, self.path, want_zq=False,
                                                    want_loss=False)
        self.code_counts = counts
        return indices

    def get_embedding_from_one_hot(self, min_encoding_indices, target_shape):
        """E[min_encoding_indices].view(target_shape), :121-131."""
        return ops.gather(min_encoding_indices, self.embedding.weight, target_shape)

    def reduce_stats(self, loss, async_op: bool = True):
        """Global loss and code usage of the last forward over all ranks in ONE small all-reduce of
        [counts[K], loss * n, n], issued asynchronously by default so that it overlaps the
        backward pass; `.result()` on the returned handle gives (global_loss, global_perplexity, global_counts).
        The per-rank `loss` that drives autograd is untouched (the reference's semantics: per-rank mean, DDP averages
        the gradients)."""
        if self.code_counts is None:
            raise RuntimeError("reduce_stats() needs a forward pass first")
        return fused_stats_all_reduce(self.code_counts, loss, async_op=async_op)

    def global_perplexity(self):
        """Perplexity of the code usage summed over all ranks (opt-in; the reference never
        reduces it).  One K-element all-reduce over NCCL."""
        if self.code_counts is None:
            raise RuntimeError("global_perplexity() needs a forward pass first")
        counts = self.code_counts.clone()
        all_reduce(counts)
        p = counts.to(torch.float32) / counts.sum().to(torch.float32)
        return torch.exp(-torch.sum(p * torch.log(p + 1e-10)))


class StatsHandle:
    """Pending fused all-reduce of [counts[K], loss * n, n] (fused_stats_all_reduce)."""

    def __init__(self, buf, work, k):
        self._buf, self._work, self._k = buf, work, k

    def result(self):
        """(global_loss, global_perplexity, global_counts int64 (K,)) -- waits for the collective if it is pending."""
        if self._work is not None:
            self._work.wait()
            self._work = None
        k = self._k
        counts, n = self._buf[:k], self._buf[k + 1]
        p = (counts / n).to(torch.float32)
        ppl = torch.exp(-torch.sum(p * torch.log(p + 1e-10)))
        return (self._buf[k] / n).to(torch.float32), ppl, counts.round().to(torch.int64)


def fused_stats_all_reduce(counts, loss, async_op: bool = True) -> StatsHandle:
    """One all-reduce (SUM) of the float64 vector [counts[K], loss * n, n] with n = counts.sum(): counts stay exact up
    to 2^53, the global loss is the n-weighted mean of the per-rank means.  Single-process runs skip the collective."""
    k = counts.numel()
    buf = torch.empty(k + 2, dtype=torch.float64, device=counts.device)
    buf[:k] = counts
    n = buf[:k].sum()
    buf[k] = loss.detach().to(torch.float64) * n
    buf[k + 1] = n
    work = None
    if get_world_size() > 1:
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=async_op)
        if not async_op:
            work = None
    return StatsHandle(buf, work, k)


def get_world_size():
    """model/vector_quantizer.py:134-141."""
    if not dist.is_available() or not dist.is_initialized():
        return 1
    return dist.get_world_size()


def all_reduce(tensor, op=dist.ReduceOp.SUM):
    """In-place all-reduce guarded for single-process runs (model/vector_quantizer.py:144-152).
    Unlike the reference, the reduced tensor is returned in the multi-rank case too."""
    if get_world_size() == 1:
        return tensor
    dist.all_reduce(tensor, op=op)
    return tensor
