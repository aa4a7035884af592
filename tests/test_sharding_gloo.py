"""N>1 host logic on CPU: two gloo ranks shard a batch of cycles, encode their shards with a
stand-in encoder, gather the ids and all-reduce the code histogram -- the same plumbing the
B200 ranks run over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import vqb200  # noqa: F401
from vqb200.dataloader import bulk_encode_ids, gather_sharded, reduce_counts, shard_range


def test_shard_range_partitions_in_order():
    for n in (0, 1, 7, 16, 1000, 10_000_001):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _fake_encode(cycles: torch.Tensor) -> torch.Tensor:
    # deterministic stand-in for model.encode_ids: 4 "tokens" per cycle in [0, 16)
    b = cycles.shape[0]
    feat = cycles.reshape(b, 4, -1).sum(-1)
    return (feat.abs() * 1000).long() % 16


def _worker(rank, world, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(5)
        cycles = torch.randn(n, 8, 2, generator=g)
        ids = bulk_encode_ids(_fake_encode, cycles, batch=5)
        local, (lo, hi) = bulk_encode_ids(_fake_encode, cycles, batch=5, gather=False)
        counts = reduce_counts(torch.bincount(local.reshape(-1), minlength=16))
        again = torch.bincount(local.reshape(-1), minlength=16)
        work = reduce_counts(again, async_op=True)          # same sum through the asynchronous form
        work.wait()
        assert torch.equal(again, counts)
        assert vqb200.get_world_size() == world
        t = torch.ones(3)
        assert vqb200.all_reduce(t) is t and torch.equal(t, torch.full((3,), float(world)))
        # the training step's fused small all-reduce: [counts[K], loss * n, n] in one collective
        from vqb200.model.vector_quantizer import fused_stats_all_reduce
        my_counts = torch.bincount(local.reshape(-1), minlength=16)
        my_loss = torch.tensor(0.5 + rank, dtype=torch.float32)
        g_loss, g_ppl, g_counts = fused_stats_all_reduce(my_counts, my_loss, async_op=True).result()
        torch.save(dict(ids=ids, counts=counts, lo=lo, hi=hi, my_n=int(my_counts.sum()), g_loss=g_loss, g_ppl=g_ppl,
                        g_counts=g_counts), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [23, 16, 1])
def test_two_rank_bulk_encode_equals_single_process(tmp_path, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    g = torch.Generator().manual_seed(5)
    cycles = torch.randn(n, 8, 2, generator=g)
    want = _fake_encode(cycles)
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    assert torch.equal(r0["ids"], want) and torch.equal(r1["ids"], want)
    assert r0["hi"] == r1["lo"] and r0["lo"] == 0 and r1["hi"] == n
    full = torch.bincount(want.reshape(-1), minlength=16)
    assert torch.equal(r0["counts"], full) and torch.equal(r1["counts"], full)
    # fused statistics: exact counts, n-weighted mean of the per-rank losses, perplexity of the global histogram
    n0, n1 = r0["my_n"], r1["my_n"]
    want_loss = (0.5 * n0 + 1.5 * n1) / max(n0 + n1, 1)
    p = full.float() / full.sum().float()
    want_ppl = torch.exp(-torch.sum(p * torch.log(p + 1e-10)))
    for r in (r0, r1):
        assert torch.equal(r["g_counts"], full)
        assert abs(float(r["g_loss"]) - want_loss) < 1e-6
        assert abs(float(r["g_ppl"]) - float(want_ppl)) < 1e-5


def _stream_worker(rank, world, port, n, seq_len, out_dir):
    """create_latent_space_dataset_from_cycles(shard=True) on two gloo ranks with a stand-in per-cycle encoder."""
    import types
    import numpy as np
    from vqb200.dataloader import LatentSpaceEncoder
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Stub(torch.nn.Module):
            enc_out_len, embedding_dim = 4, 2
            vector_quantization = types.SimpleNamespace(code_counts=None)
        enc = LatentSpaceEncoder(Stub(), window_size=8, device="cpu", encoder_mode=None)
        seen = []
        enc.get_latent_space_IDs = lambda x, p=False: (seen.append(x.shape[0]), _fake_encode(x).view(-1, 1))[1]
        cycles = torch.randn(n, 8, 2, generator=torch.Generator().manual_seed(5))
        y = np.arange(n, dtype=np.float64)
        ids, labels = enc.create_latent_space_dataset_from_cycles(cycles, y, seq_len=seq_len, batch=5, shard=True)
        torch.save(dict(ids=torch.from_numpy(ids), labels=torch.from_numpy(labels), encoded=sum(seen)),
                   os.path.join(out_dir, f"s{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,seq_len", [(23, 4), (9, 1), (1, 1), (3, 4)])
def test_two_rank_data_set_from_the_cycle_stream(tmp_path, n, seq_len):
    """Every rank encodes its shard of the cycles once, the per-cycle ids are all-gathered, and both ranks hold the windows
    a single process builds (n - seq_len windows, window i = cycles i .. i + seq_len - 1, label y[i + seq_len])."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_stream_worker, args=(2, port, n, seq_len, str(tmp_path)), nprocs=2, join=True)
    cycles = torch.randn(n, 8, 2, generator=torch.Generator().manual_seed(5))
    per_cycle = _fake_encode(cycles)
    n_windows = n if seq_len == 1 else max(n - seq_len, 0)
    want = torch.stack([per_cycle[i:i + seq_len] for i in range(n_windows)]) if n_windows else torch.empty(0, seq_len, 4, dtype=torch.int64)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"s{r}.pt")) for r in range(2))
    for r in (r0, r1):
        assert tuple(r["ids"].shape) == (n_windows, seq_len, 4) and torch.equal(r["ids"], want)
        assert torch.equal(r["labels"], torch.arange(n, dtype=torch.float64)[(seq_len if seq_len > 1 else 0):][:n_windows])
    assert r0["encoded"] + r1["encoded"] == n and abs(r0["encoded"] - r1["encoded"]) <= 1      # each cycle once, on one rank


def test_single_process_paths():
    cycles = torch.randn(10, 8, 2)
    ids = bulk_encode_ids(_fake_encode, cycles, batch=3)
    assert torch.equal(ids, _fake_encode(cycles))
    assert gather_sharded(ids, 10) is ids
    c = torch.ones(4, dtype=torch.int64)
    assert reduce_counts(c) is c
    assert reduce_counts(c, async_op=True) is None
