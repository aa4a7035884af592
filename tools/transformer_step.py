"""GPU box: BASELINE configs[4] context -- a next-token training step of a minGPT-style decoder (the consumer of the
ids: d_model 512, 8 blocks, 8 heads, T = n_cycles * 16 + 1 = 321, num_embeddings + 2 = 258 classes;
train_transformer_mtasks.py:145-146,175-176) fed by ON-THE-FLY tokenisation of synthetic windows (16 windows of
20 cycles per step, :214) instead of the pickled data set.  The decoder is stock PyTorch (out of the hot-path scope;
attention through F.scaled_dot_product_attention); what is measured is the tokeniser -- fully fused encoder launch +
fused quantiser + vqb_ar_pairs -- as a share of the step.

    python tools/transformer_step.py [batch_windows]"""
import json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn, torch.nn.functional as F
import vqb200
from vqb200.dataloader import OnTheFlyTokenizer

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
N_CYCLES, D_MODEL, N_HEAD, N_BLOCKS = 20, 512, 8, 8


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.ln1, self.ln2 = nn.LayerNorm(D_MODEL), nn.LayerNorm(D_MODEL)
        self.qkv, self.proj = nn.Linear(D_MODEL, 3 * D_MODEL), nn.Linear(D_MODEL, D_MODEL)
        self.fc, self.out = nn.Linear(D_MODEL, 4 * D_MODEL), nn.Linear(4 * D_MODEL, D_MODEL)

    def forward(self, x):
        b, t, c = x.shape
        q, k, v = self.qkv(self.ln1(x)).view(b, t, 3, N_HEAD, c // N_HEAD).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(b, t, c)
        x = x + self.proj(a)
        return x + self.out(F.gelu(self.fc(self.ln2(x)), approximate="tanh"))


class Decoder(nn.Module):
    def __init__(self, n_classes, seq_len):
        super().__init__()
        self.tok = nn.Embedding(n_classes, D_MODEL)
        pe = torch.zeros(seq_len, D_MODEL)
        pos = torch.arange(seq_len).float().unsqueeze(1)
        div = (torch.arange(0, D_MODEL, 2).float() * -(math.log(10000.0) / D_MODEL)).exp()
        pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
        self.register_buffer("pe", pe)
        self.blocks = nn.ModuleList(Block() for _ in range(N_BLOCKS))
        self.ln_f, self.lm_head = nn.LayerNorm(D_MODEL), nn.Linear(D_MODEL, n_classes, bias=False)

    def forward(self, x):
        h = self.tok(x) + self.pe[: x.shape[1]]
        for blk in self.blocks:
            h = blk(h)
        return self.lm_head(self.ln_f(h))


torch.manual_seed(0)
vqvae = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                          learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
tok = OnTheFlyTokenizer(vqvae, window_size=200, device=str(dev))
T = N_CYCLES * vqvae.enc_out_len + 1
model = Decoder(tok.num_classes, T).to(dev).train()
opt = torch.optim.RAdam(model.parameters(), lr=1e-4)
g = torch.Generator(device=dev).manual_seed(7)
windows = torch.randn(B, N_CYCLES * 200, 2, device=dev, generator=g)


def step(timed=None):
    if timed:
        timed[0].record()
    x, cond, y = tok(windows)
    if timed:
        timed[1].record()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(x)
    loss = F.cross_entropy(logits.float().view(-1, logits.shape[-1]), y.view(-1))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.8)
    opt.step()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
steps, tk_ms = 10, 0.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
pairs = []
e0.record()
for _ in range(steps):
    p = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    loss = step(p)
    pairs.append(p)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
tk_ms = sum(a.elapsed_time(b) for a, b in pairs) / steps
print(json.dumps({"workload": f"next-token step, d_model {D_MODEL}, {N_BLOCKS} blocks, {N_HEAD} heads, T = {T}, "
                  f"{tok.num_classes} classes, {B} windows x {N_CYCLES} cycles per step, bf16 autocast, RAdam, "
                  "tokens from OnTheFlyTokenizer (fused_bf16 encoder + fused quantiser + vqb_ar_pairs)",
                  "ms_per_step": ms, "tokens_per_s": B * T / (ms * 1e-3), "tokenizer_ms": tk_ms,
                  "tokenizer_share": tk_ms / ms, "cycles_tokenised_per_step": B * N_CYCLES, "loss": float(loss)}))
