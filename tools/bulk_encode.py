"""GPU box: BASELINE configs[2] -- bulk latent-dataset encoding (dataloader/latentspace_dataloader.py path):
synthetic welding cycles randn(n, 200, 2) generated on the device in 65536-cycle chunks, random-init
VQ-VAE-Patch (repo defaults, seed 0), ids out (what the id tasks keep).  Reports patches/s for the encode call
(patchify + encoder + fused VQ) per encoder precision, the share of the fused VQ kernel, and the index match
rate against the fp32 ('highest') encoder on the first chunk.  Modes: the stock PyTorch layers in fp32 / TF32 / bf16
autocast (round 1's baseline rows) and `fused_bf16` = the one-launch tcgen05 encoder chain;
the quantiser is the tcgen05 kernel in all of them.  `bench.py`'s `bulk_encode` object is the figure of record.

    python tools/bulk_encode.py [n_cycles]            (torchrun: every rank encodes its own n_cycles)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import vqb200

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); dist.init_process_group("nccl", device_id=dev)
n_cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
chunk = 1 << 16
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
if world > 1:
    for t in model.state_dict().values():
        dist.broadcast(t, 0)
T = model.enc_out_len

def encode_all(mode):
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.set_float32_matmul_precision("high" if mode == "tf32" else "highest")
    model.encoder_mode = "fused_bf16" if mode == "fused_bf16" else "torch"
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    counts = torch.zeros(256, dtype=torch.int64, device=dev)
    first = None
    vq_ms = 0.0
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for s in range(0, n_cycles, chunk):
            x = torch.randn(min(chunk, n_cycles - s), 200, 2, device=dev, generator=g)
            if mode == "bf16":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    z_e = model.encode(x)
                z_e = z_e.float()
            else:
                z_e = model.encode(x)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ids = model.vector_quantization.encode_indices(z_e)
            b.record()
            counts += model.vector_quantization.code_counts
            if first is None:
                first = ids.view(-1, T).clone()
            pairs.append((a, b))
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), first, counts

out = {}
ref_ids = None
for mode in ("fp32", "tf32", "bf16", "fused_bf16"):
    pairs = []
    n_keep, n_cycles = n_cycles, min(n_cycles, chunk)
    encode_all(mode)                 # one untimed chunk per mode: allocator pools, cuBLAS handles, kernel attributes
    n_cycles = n_keep
    pairs = []
    ms, first, counts = encode_all(mode)
    vq_ms = sum(a.elapsed_time(b) for a, b in pairs)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(counts)
    if ref_ids is None:
        ref_ids = first
    out[mode] = {"patches_per_s": world * n_cycles * T / (t.item() * 1e-3), "ms": t.item(), "vq_share": vq_ms / ms,
                 "index_match_vs_fp32_encoder": float((first == ref_ids).float().mean().item()),
                 "codes_used": int((counts > 0).sum().item())}
if rank == 0:
    print(json.dumps({"workload": f"bulk encode, {n_cycles} cycles x {T} patches per GPU, chunks of {chunk} cycles, "
                      "VQ-VAE-Patch H=512 8 resblocks K=256 D=32, ids out", "n_gpus": world, "modes": out}))
if world > 1:
    dist.destroy_process_group()
