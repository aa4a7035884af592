"""GPU box: torch-profiler kernel table of one 65536-cycle bulk-encode chunk in fused_bf16 mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, vqb200
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
model.encoder_mode = "fused_bf16"
x = torch.randn(65536, 200, 2, device=dev)
with torch.no_grad():
    for _ in range(2): model.encode_ids(x)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): model.encode_ids(x)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
