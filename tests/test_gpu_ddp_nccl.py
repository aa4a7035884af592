"""Two-rank NCCL test (needs >= 2 GPUs on the box; skipped otherwise): a DDP-wrapped VQ-VAE-Patch training step on
shards of a batch gives the gradients, the code counts and the loss of one GPU on the concatenated batch
("DDP grads + count all-reduce equal single-GPU on the concatenated batch").
Reference: train_reconstruction_embedding.py:190-202, model/autencoder_lightning_base.py:86-97."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases as C

pytestmark = pytest.mark.gpu


def _build(patch_sd, case, dev):
    import vqb200
    model = vqb200.VQVAEPatch(hidden_dim=case["hidden_dim"], input_dim=case["input_dim"],
                              num_embeddings=case["num_embeddings"], embedding_dim=case["embedding_dim"],
                              n_resblocks=case["n_resblocks"], learning_rate=1e-3, dropout_p=0.0,
                              patch_size=case["patch_size"], seq_len=case["seq_len"],
                              batch_norm=case["batch_norm"], beta=case["beta"])
    model.load_state_dict(patch_sd, strict=True)
    # eval mode with gradients: the decoder's BatchNorm1d (PatchEmbeddingInverse, model/vq_vae_patch_embedd.py:19-57) then
    # uses its running statistics -- in training mode each rank would normalise with the statistics of its own shard
    # (the reference does not use SyncBatchNorm), which no single-GPU run on the whole batch reproduces
    return model.to(dev).eval()


def _worker(rank, world, port, patch_sd, case, x_all, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        model = _build(patch_sd, case, dev)
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[rank])
        per = x_all.shape[0] // world
        x = x_all[rank * per:(rank + 1) * per].to(dev)
        emb_loss, x_hat, _ = net(x)
        handle = model.vector_quantization.reduce_stats(emb_loss)          # one fused small all-reduce, overlaps backward
        loss = torch.nn.functional.mse_loss(x_hat, x) + emb_loss
        loss.backward()
        g_loss, g_ppl, g_counts = handle.result()
        grads = {k: p.grad.detach().cpu() for k, p in model.named_parameters() if p.grad is not None}
        torch.save(dict(grads=grads, g_loss=g_loss.cpu(), g_ppl=g_ppl.cpu(), g_counts=g_counts.cpu(),
                        loss=loss.detach().cpu()), os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_ddp_step_equals_single_gpu_on_the_concatenated_batch(tmp_path, patch_golden):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = C.PATCH_CASES[0]
    pre = f"{case['name']}/sd/"
    sd = {k[len(pre):]: torch.from_numpy(patch_golden[k]) for k in patch_golden.files if k.startswith(pre)}
    rs = np.random.RandomState(77)
    x_all = torch.from_numpy(rs.standard_normal((16, case["seq_len"], case["input_dim"])).astype(np.float32))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, sd, case, x_all, str(tmp_path)), nprocs=2, join=True)
    # single GPU, whole batch
    dev = torch.device("cuda", 0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    model = _build(sd, case, dev)
    x = x_all.to(dev)
    emb_loss, x_hat, ppl = model(x)
    loss = torch.nn.functional.mse_loss(x_hat, x) + emb_loss
    loss.backward()
    counts = model.vector_quantization.code_counts.cpu()
    r0 = torch.load(os.path.join(tmp_path, "r0.pt"))
    r1 = torch.load(os.path.join(tmp_path, "r1.pt"))
    # DDP averages the per-rank gradients of per-rank means = gradient of the global mean for equal shards
    for k, p in model.named_parameters():
        if p.grad is None:
            continue
        ref = p.grad.cpu()
        for r in (r0, r1):
            torch.testing.assert_close(r["grads"][k], ref, rtol=2e-4, atol=2e-6 + 2e-5 * float(ref.abs().max()))
    for r in (r0, r1):
        assert torch.equal(r["g_counts"], counts)
        assert float(r["g_loss"]) == pytest.approx(float(emb_loss), rel=1e-5)
        assert float(r["g_ppl"]) == pytest.approx(float(ppl), rel=1e-5)
    assert 0.5 * (float(r0["loss"]) + float(r1["loss"])) == pytest.approx(float(loss), rel=1e-5)


def _stream_worker(rank, world, port, patch_sd, case, stream, labels, seq_len, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from vqb200.dataloader import LatentSpaceEncoder
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        enc = LatentSpaceEncoder(_build(patch_sd, case, dev), window_size=200, device=str(dev), encoder_mode="torch")
        ids, y = enc.create_latent_space_dataset_from_cycles(stream, labels, seq_len=seq_len, has_patch_embed=True, batch=4,
                                                             shard=True)
        np.savez(os.path.join(out_dir, f"s{rank}.npz"), ids=ids, y=y)
    finally:
        dist.destroy_process_group()


def test_sharded_data_set_from_the_cycle_stream_matches_the_reference_fixture(tmp_path, patch_golden, bulk_golden):
    """Bulk latent-dataset building batch-sharded over two GPUs (north star: codebook and encoder replicated, every rank
    encodes its shard of the cycles, the ids are all-gathered over NCCL): both ranks hold the arrays the unmodified
    reference built from its own windows (tests/golden: bulk_overlap/seq_*)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    case = C.BULK_CASE
    mcase = next(c for c in C.PATCH_CASES if c["name"] == case["model"])
    pre = f"{mcase['name']}/sd/"
    sd = {k[len(pre):]: torch.from_numpy(patch_golden[k]) for k in patch_golden.files if k.startswith(pre)}
    stream, cycle_labels = C.make_stream(case)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_stream_worker, args=(2, port, sd, mcase, stream, cycle_labels, case["seq_len"], str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        got = np.load(os.path.join(tmp_path, f"s{r}.npz"))
        assert np.array_equal(got["ids"], bulk_golden[f"{case['name']}/seq_ids"])
        assert np.array_equal(got["y"], bulk_golden[f"{case['name']}/seq_labels"])
