"""GPU box: BASELINE configs[3] -- VQ-VAE-Patch training step (train_reconstruction_embedding.py path) with the
fused VQ forward + straight-through backward inside stock-PyTorch encoder/decoder, RAdam, B = 1024 cycles per GPU
(script default), synthetic randn cycles.  Under torchrun the model is wrapped in DDP (NCCL all-reduce of the
gradients incl. the dense codebook gradient).  Reports ms/step, patches/s and the share of the fused VQ kernels.

    python tools/train_step.py            |   python -m torch.distributed.run --nproc-per-node N tools/train_step.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vqb200

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
res = {}
for mode in ("fp32", "tf32"):
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.set_float32_matmul_precision("high" if mode == "tf32" else "highest")
    torch.manual_seed(0)
    model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                              learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.RAdam(net.parameters(), lr=1e-3)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    x = torch.randn(B, 200, 2, device=dev, generator=g)
    vq_ms = [0.0]
    vq = model.vector_quantization
    orig = vq.forward
    def timed_forward(z, _orig=orig):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = _orig(z); b.record()
        pairs.append((a, b))
        return out
    vq.forward = timed_forward
    import contextlib
    stats = {}
    def step(sync=True):
        opt.zero_grad(set_to_none=True)
        with (contextlib.nullcontext() if sync or world == 1 else net.no_sync()):
            emb_loss, x_hat, ppl = net(x)
            # global loss / perplexity / code counts: ONE fused small all-reduce, issued now, collected after backward
            handle = vq.reduce_stats(emb_loss) if sync else None
            loss = torch.nn.functional.mse_loss(x_hat, x) + emb_loss
            loss.backward()
        if handle is not None:
            stats["global"] = handle.result()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        return loss
    pairs = []
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    pairs = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    vq_fwd_ms = sum(a.elapsed_time(b) for a, b in pairs) / steps
    # the same step without the gradient all-reduce (DDP no_sync) and without the statistics all-reduce: what NCCL
    # adds to the step as exposed (non-overlapped) time
    ms_nosync = ms
    if world > 1:
        for _ in range(2):
            step(sync=False)
        torch.cuda.synchronize(); dist.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(steps):
            step(sync=False)
        f1.record(); torch.cuda.synchronize()
        t2 = torch.tensor([f0.elapsed_time(f1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms_nosync = t2.item() / steps
    res[mode] = {"ms_per_step": ms, "patches_per_s": world * B * model.enc_out_len / (ms * 1e-3),
                 "vq_forward_ms": vq_fwd_ms, "vq_forward_share": vq_fwd_ms / ms, "loss": float(loss.item()),
                 "ms_per_step_without_nccl": ms_nosync, "nccl_exposed_ms": ms - ms_nosync,
                 "nccl_exposed_share": (ms - ms_nosync) / ms,
                 "grad_bytes_all_reduced": int(sum(p.numel() for p in model.parameters()) * 4),
                 "global_stats": ({"loss": float(stats["global"][0]), "perplexity": float(stats["global"][1]),
                                   "codes_used": int((stats["global"][2] > 0).sum())} if "global" in stats else None)}
    vq.forward = orig
if rank == 0:
    print(json.dumps({"workload": f"VQ-VAE-Patch train step, B={B} cycles/GPU, H=512, 8 resblocks, K=256, D=32, RAdam, "
                      "grad clip 1.0", "n_gpus": world, "modes": res}))
if world > 1:
    dist.destroy_process_group()
