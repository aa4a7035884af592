"""GPU box, for ncu: a few fully fused encode calls (vqb_patch_split + vqb_encoder_chain with the patch embedding as its
first and the projection as its last GEMM) of the default model on `n` cycles (default 16384 = 2^18 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
model.encoder_mode = "fused_bf16"
x = torch.randn(n, 200, 2, device=dev)
with torch.no_grad():
    for _ in range(4):
        z = model.encode(x)
torch.cuda.synchronize()
print("done", tuple(z.shape))
