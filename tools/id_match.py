"""GPU box: id match of the fused bf16 encoder variants against the fp32 encoder (TF32 off) on synthetic cycles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
x = torch.randn(n, 200, 2, device=dev, generator=torch.Generator(device=dev).manual_seed(1000))
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
with torch.no_grad():
    ref = torch.cat([model.encode_ids(x[i:i + 8192]) for i in range(0, n, 8192)]).view(-1)
    model.encoder_mode = "fused_bf16"
    for chain, proj in ((False, False), (True, False), (True, True)):
        model.fused_chain, model.fused_projection = chain, proj
        ids = torch.cat([model.encode_ids(x[i:i + 8192]) for i in range(0, n, 8192)]).view(-1)
        print(f"chain={chain} fused_projection={proj}: id match {float((ids == ref).float().mean()):.5f} on {ids.numel()} tokens")
