"""GPU box, under compute-sanitizer (ONE tool per gpurun call):
    compute-sanitizer --tool memcheck  --kernel-name kns=3vqb python tools/sanitize_small.py
    compute-sanitizer --tool racecheck --kernel-name kns=3vqb python tools/sanitize_small.py
Small-N calls of every hand-written kernel of the path (all live in namespace vqb), each checked against its
reference so that the run is also a functional test under the tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, vqb200
from vqb200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
n = 128 * 5 + 37
z = 0.1 * torch.randn(n, 32, device=dev)
w = (torch.rand(256, 32, device=dev) * 2 - 1) / 256
ref = ops.forward(z, w, 0.25, path="fma")
out = ops.forward(z, w, 0.25, path="tc")
assert torch.equal(ref[3], out[3]) and torch.equal(ref[1], out[1])
print("forward tc / fma ok")
wide = 0.1 * torch.randn(n, 64, device=dev); w64 = (torch.rand(300, 64, device=dev) * 2 - 1) / 300
a, b = ops.forward(wide, w64, 0.25, path="fma"), ops.forward(wide, w64, 0.25, path="tc")
assert torch.equal(a[3], b[3])
print("forward wide / chunked ok")
gq = torch.randn(n, 32, device=dev); gl = torch.tensor(1.0, device=dev)
gz, ge = ops.backward(gq, gl, z, out[3], w, 0.25)
res = w[out[3].view(-1)] - z
ge_ref = torch.zeros_like(w).index_add_(0, out[3].view(-1), res) * (0.25 * 2 / (n * 32))
assert torch.allclose(ge, ge_ref, rtol=1e-4, atol=1e-8) and torch.allclose(gz, gq - 2 * res / (n * 32), rtol=1e-5, atol=1e-8)
print("backward ok")
perm = torch.randn(40, 32, 16, device=dev).permute(0, 2, 1)
assert torch.equal(ops.pack_rows(perm, 32), perm.contiguous())
print("pack_rows ok")
T, H, L = 128 * 2 + 9, 512, 4
h0 = torch.randn(T, H, device=dev)
wl = (torch.randn(L, H, H, device=dev) * (1.0 / H) ** 0.5).to(torch.bfloat16)
bl = 0.1 * torch.randn(L, H, device=dev)
a0 = torch.nn.functional.gelu(h0).to(torch.bfloat16)
href = h0.clone(); aa, uu = a0.clone(), torch.empty_like(a0)
for i in range(L // 2):
    ops.token_linear(aa, wl[2 * i], bl[2 * i], out=uu, mode=0)
    ops.token_linear(uu, wl[2 * i + 1], bl[2 * i + 1], h=href, out=aa if 2 * i + 2 < L else None, mode=1)
hc = ops.encoder_chain(a0, h0.clone(), wl, bl)
assert torch.allclose(hc, href, rtol=1e-5, atol=1e-5)
wp = torch.randn(32, H, device=dev) * (1.0 / H) ** 0.5
zp = ops.encoder_chain(a0, h0.clone(), torch.cat([wl.reshape(-1, H), ops.projection_rows(wp)]).contiguous(), bl,
                       proj_bias=torch.zeros(32, device=dev))
assert torch.allclose(zp, hc @ wp.t(), rtol=1e-3, atol=1e-3)
print("token_linear / encoder_chain ok")
torch.manual_seed(1)
model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=64, embedding_dim=32, n_resblocks=2,
                          learning_rate=1e-3, patch_size=25, batch_norm=False).to(dev).eval()
xm = torch.randn(19, 200, 2, device=dev)
with torch.no_grad():
    model.encoder_mode = "fused_bf16"
    model.fused_patch_embed = False
    z1 = model.encode(xm)
    model.fused_patch_embed = True
    z2 = model.encode(xm)                      # vqb_patch_split + the fully fused chain launch
assert (z1 - z2).abs().max().item() <= 0.02 * z1.abs().max().item()
print("fully fused encoder launch ok")
x = torch.randn(5, 200, 2, device=dev)
conv = torch.nn.Conv1d(1, 512, 25, stride=25)
with torch.no_grad():      # fp32 reference on the CPU (cuDNN would run the conv in TF32)
    r = conv(x.cpu().permute(0, 2, 1).reshape(5, 1, -1)).permute(0, 2, 1).reshape(-1, 512).to(dev)
conv = conv.to(dev)
hh, act = ops.patch_embed(x, conv.weight, conv.bias, 25)
assert torch.allclose(hh, r, rtol=1e-4, atol=1e-5)
print("patch_embed ok")
# tile-stationary kernel: large codebook / wide vectors against the exact FMA kernel
for K_, D_ in ((700, 32), (300, 128)):
    zz_ = 0.1 * torch.randn(128 * 3 + 5, D_, device=dev); ww_ = 0.1 * torch.randn(K_, D_, device=dev)
    a, b = ops.forward(zz_, ww_, 0.25, path="fma"), ops.forward(zz_, ww_, 0.25, path="tc")
    assert torch.equal(a[3], b[3]) and torch.equal(a[1], b[1]) and torch.equal(a[4], b[4])
print("forward tile-stationary ok")
# decoder: three-tap layer against Conv1d, then the whole fused decoder against the PyTorch modules
a3 = torch.randn(5 * 16, 256, device=dev).to(torch.bfloat16)
w3 = (torch.randn(256, 256, 3, device=dev) * 0.04).to(torch.bfloat16)
b3 = 0.1 * torch.randn(256, device=dev)
got = ops.token_conv(a3, w3.permute(0, 2, 1).reshape(256, 768).contiguous(), b3, mode=0, taps=3, tokens_per_cycle=16, out_gelu=False)
want = torch.nn.functional.conv1d(a3.float().view(5, 16, 256).permute(0, 2, 1), w3.float(), b3, padding=1).permute(0, 2, 1).reshape(80, 256)
assert torch.allclose(got.float(), want, rtol=1.0 / 128, atol=2e-3)
with torch.no_grad():
    zq_ = 0.3 * torch.randn(19, 16, 32, device=dev)
    ref_hat = model.decode(zq_)
    model.decoder_mode = "fused_bf16"
    hat = model.decode(zq_)
assert (hat - ref_hat).abs().max().item() <= 0.02 * ref_hat.abs().max().item()
print("token_conv / fused decoder ok")
torch.cuda.synchronize()
print("all ok")
