"""GPU tests of the fused per-token linear layer (csrc/tok_linear.cu) against a plain PyTorch fp32 reference of the
same op on the same bf16-rounded operands (tolerances: fp32 accumulation order + one bf16 rounding of the output)."""
import numpy as np
import pytest
import torch

import cases as C
import vqb200
from vqb200 import ops

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _ref(a, w, bias, h, mode):
    acc = a.float() @ w.float().t() + bias
    if mode == 1:
        acc = acc + h
    return acc, torch.nn.functional.gelu(acc)


@pytest.mark.parametrize("T", [1, 127, 128, 129, 1000, 4096 + 77])
@pytest.mark.parametrize("K,N", [(512, 512), (64, 256), (128, 768)])
def test_token_linear_matches_fp32_reference(T, K, N):
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(T * 7 + K + N)
    a = torch.randn(T, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device=dev, generator=g) * (2.0 / K) ** 0.5).to(torch.bfloat16)
    bias = 0.1 * torch.randn(N, device=dev, generator=g)
    # mode 0
    out = ops.token_linear(a, w, bias, mode=0)
    _, ref = _ref(a, w, bias, None, 0)
    torch.testing.assert_close(out.float(), ref, rtol=1.0 / 128, atol=2e-3)      # bf16 output: 2^-8 relative
    # mode 1: residual stream updated in place, next activation written
    h = torch.randn(T, N, device=dev, generator=g)
    h0 = h.clone()
    nxt = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
    ops.token_linear(a, w, bias, h=h, out=nxt, mode=1)
    acc, ref = _ref(a, w, bias, h0, 1)
    torch.testing.assert_close(h, acc, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nxt.float(), ref, rtol=1.0 / 128, atol=2e-3)
    # mode 2: h written (not read), next activation written
    h3 = torch.full((T, N), float("nan"), device=dev)
    nxt3 = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
    ops.token_linear(a, w, bias, h=h3, out=nxt3, mode=2)
    acc3, ref3 = _ref(a, w, bias, None, 0)
    torch.testing.assert_close(h3, acc3, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nxt3.float(), ref3, rtol=1.0 / 128, atol=2e-3)
    # mode 1 without the activation output (last block)
    h2 = h0.clone()
    ops.token_linear(a, w, bias, h=h2, out=None, mode=1)
    assert torch.equal(h2, h)


def _pair_value(p):
    n = p.shape[1] // 2
    return p[:, :n].double() + p[:, n:].double()


@pytest.mark.parametrize("T", [1, 127, 129, 1000, 128 * 150 + 77])
@pytest.mark.parametrize("K,N", [(512, 512), (64, 256), (256, 768)])
def test_token_linear_split_is_fp32_faithful(T, K, N):
    """vqb_token_linear_split against an fp64 reference on the SAME fp32 operands: hi + lo pairs carry 2^-17 per operand,
    the dropped lo x lo term 2^-18, so the result must sit within ~2^-15 of |a| . |w| -- 200x tighter than the bf16 form,
    and the same order as an fp32 GEMM's own rounding for these K."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(T * 11 + K + N)
    x = torch.randn(T, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) * (2.0 / K) ** 0.5
    bias = 0.1 * torch.randn(N, device=dev, generator=g)
    a = ops.token_pair(x, gelu=False)
    assert (_pair_value(a) - x.double()).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()
    ag = ops.token_pair(x, gelu=True)
    gx = torch.nn.functional.gelu(x.double())
    assert bool(((_pair_value(ag) - gx).abs() <= 2.0 ** -17 * gx.abs() + 5e-7).all())      # erff + the pair's 2^-18
    wp = ops.bf16_pair(w)
    bound = 2.0 ** -14 * (x.double().abs() @ w.double().abs().t()) + 1e-6      # elementwise
    acc = x.double() @ w.double().t() + bias.double()
    # mode 0: pair of gelu(acc)
    out = ops.token_linear_split(a, wp, bias, mode=0)
    assert out.shape == (T, 2 * N)
    assert bool(((_pair_value(out) - torch.nn.functional.gelu(acc)).abs() <= bound).all())
    # mode 1: residual stream in place, pair of gelu(h) / of h itself
    h = torch.randn(T, N, device=dev, generator=g)
    h0 = h.clone()
    nxt = torch.empty(T, 2 * N, dtype=torch.bfloat16, device=dev)
    ops.token_linear_split(a, wp, bias, h=h, out=nxt, mode=1)
    want = acc + h0.double()
    assert bool(((h.double() - want).abs() <= bound + 2.0 ** -22 * want.abs()).all())
    assert bool(((_pair_value(nxt) - torch.nn.functional.gelu(want)).abs() <= bound + 2.0 ** -16 * want.abs()).all())
    h1 = h0.clone()
    ops.token_linear_split(a, wp, bias, h=h1, out=nxt, mode=1, out_gelu=False)
    assert torch.equal(h1, h)
    assert bool(((_pair_value(nxt) - h.double()).abs() <= 2.0 ** -16 * h.double().abs() + 1e-30).all())
    # mode 2: h written (not read), no activation output
    h2 = torch.full((T, N), float("nan"), device=dev)
    ops.token_linear_split(a, wp, bias, h=h2, mode=2, out_gelu=False)
    assert bool(((h2.double() - acc).abs() <= bound).all())
    # and the fp32-operand error is far below the bf16 form's on the same data
    h16 = torch.empty(T, N, device=dev)
    ops.token_linear(x.to(torch.bfloat16).contiguous(), w.to(torch.bfloat16).contiguous(), bias, h=h16, mode=2)
    if T >= 127:
        assert (h2.double() - acc).abs().max().item() * 50 < (h16.double() - acc).abs().max().item()


def test_fused_fp32_encoder_matches_the_fp32_encoder():
    """VQVAEPatch.encode in 'fused_fp32' mode at the repo-default architecture: z_e within 1e-5 of its range of the fp32
    PyTorch layers, ids equal except where two codes are that close (none expected on 4800 tokens)."""
    dev = _dev()
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    for hidden, n_res, bn in ((512, 8, False), (256, 2, True), (512, 0, False)):
        model = vqb200.VQVAEPatch(hidden_dim=hidden, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=n_res,
                                  learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=bn).to(dev)
        if bn:      # non-trivial running statistics, then eval mode (the statistics are folded into the weights)
            model.train()
            with torch.no_grad():
                for _ in range(3):
                    model.encode(torch.randn(64, 200, 2, device=dev))
        model.eval()
        x = torch.randn(300, 200, 2, device=dev)
        with torch.no_grad():
            model.encoder_mode = "torch"
            z_ref = model.encode(x).double()
            ids_ref = model.encode_ids(x)
            model.encoder_mode = "fused_fp32"
            assert model._fused_ok(x)
            z = model.encode(x)
            ids = model.encode_ids(x)
        assert z.shape == z_ref.shape and z.is_contiguous() and z.dtype == torch.float32
        err = (z.double() - z_ref).abs().max().item()
        # (16 layers of 2^-18-per-operand pairs: measured ~1e-5 of the range; the bf16 form sits at ~1e-2)
        assert err <= 1e-4 * z_ref.abs().max().item(), (hidden, n_res, err, z_ref.abs().max().item())
        E = model.vector_quantization.embedding.weight.detach().cpu().numpy()
        assert C.unexplained_mismatches(z_ref.float().cpu().numpy().reshape(-1, 32), E, ids.cpu().numpy().reshape(-1),
                                        ids_ref.cpu().numpy().reshape(-1), rel=1e-4) == 0
        assert (ids == ids_ref).float().mean().item() >= 0.9995
        model.train()
        assert not model._fused_ok(x)


def test_token_linear_gelu_is_the_erf_form():
    """The epilogue's GELU is the exact (erf) form to bf16 accuracy, not the textbook two-term tanh approximation:
    |error| <= 2.5e-5 + 2.5e-4 |x| before the bf16 rounding (csrc/tok_linear.cu gelu_fast)."""
    dev = _dev()
    K, N, T = 64, 256, 256
    a = torch.zeros(T, K, device=dev, dtype=torch.bfloat16)
    w = torch.zeros(N, K, device=dev, dtype=torch.bfloat16)
    x = torch.cat([torch.linspace(-6, 6, N - 32, device=dev),
                   torch.tensor([-1e4, -300., -40., -12., -9., -8., -7., -6.5, 6.5, 7., 8., 9., 12., 40., 300., 1e4] * 2,
                                device=dev)])
    h = x.repeat(T, 1).contiguous()
    out = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
    ops.token_linear(a, w, torch.zeros(N, device=dev), h=h, out=out, mode=1)
    ref = torch.nn.functional.gelu(x)
    bound = 2.5e-5 + 2.5e-4 * x.abs() + 2.0 ** -8 * ref.abs()      # formula + MUFU.TANH + bf16 rounding
    assert bool(((out[0].float() - ref).abs() <= bound).all())


def test_token_linear_rejects_bad_arguments():
    dev = _dev()
    a = torch.zeros(8, 64, device=dev, dtype=torch.bfloat16)
    w = torch.zeros(256, 64, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(256, device=dev)
    with pytest.raises(RuntimeError):
        ops.token_linear(a.float(), w, b)
    with pytest.raises(RuntimeError):
        ops.token_linear(a, w, b, mode=1)                        # no residual stream
    with pytest.raises(RuntimeError):
        ops.token_linear(torch.zeros(8, 48, device=dev, dtype=torch.bfloat16),
                         torch.zeros(256, 48, device=dev, dtype=torch.bfloat16), b)   # K not a multiple of 64


def test_split_entry_points_reject_bad_arguments():
    dev = _dev()
    a = torch.zeros(8, 128, device=dev, dtype=torch.bfloat16)
    w = torch.zeros(256, 128, device=dev, dtype=torch.bfloat16)
    b = torch.zeros(256, device=dev)
    with pytest.raises(RuntimeError):
        ops.token_linear_split(a.float(), w, b)
    with pytest.raises(RuntimeError):
        ops.token_linear_split(a, w, b, mode=1)                                   # no residual stream
    with pytest.raises(RuntimeError):
        ops.token_linear_split(a, w, b, out=torch.empty(8, 256, device=dev, dtype=torch.bfloat16))   # out must be the (T, 2 N) pair
    with pytest.raises(RuntimeError):
        ops.token_linear_split(torch.zeros(8, 96, device=dev, dtype=torch.bfloat16),
                               torch.zeros(256, 96, device=dev, dtype=torch.bfloat16), b)            # K = 48 not a multiple of 64
    with pytest.raises(RuntimeError):
        ops.token_conv_split(torch.zeros(48, 128, device=dev, dtype=torch.bfloat16),
                             torch.zeros(256, 384, device=dev, dtype=torch.bfloat16), b, taps=3, tokens_per_cycle=24)
    with pytest.raises(RuntimeError):
        ops.token_pair(torch.zeros(4, 6, device=dev))                             # n not a multiple of 4
    assert ops.token_linear_split(torch.zeros(0, 128, device=dev, dtype=torch.bfloat16), w, b).shape == (0, 512)   # empty input


def test_fused_bf16_encoder_tracks_the_fp32_encoder():
    """VQVAEPatch.encode in 'fused_bf16' mode: same shapes, z_e within bf16-operand error of the fp32 encoder,
    ids equal for the overwhelming majority of tokens (the quantiser itself is exact on whatever z_e it gets)."""
    dev = _dev()
    torch.manual_seed(0)
    model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32, n_resblocks=8,
                              learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=False).to(dev).eval()
    x = torch.randn(300, 200, 2, device=dev)
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        model.encoder_mode = "torch"
        z_ref = model.encode(x)
        ids_ref = model.encode_ids(x)
        model.encoder_mode = "fused_bf16"
        z_fused = model.encode(x)
        ids_fused = model.encode_ids(x)
        model.encoder_mode = "torch"
    assert z_fused.shape == z_ref.shape and z_fused.is_contiguous()
    err = (z_fused - z_ref).abs().max().item()
    assert err <= 0.05 * z_ref.abs().max().item(), err
    # stated tolerance of the bf16-operand encoder: >= 99.5 % of the ids equal to the fp32 encoder's at the repo-default
    # size (measured 99.86 % on 2^20 tokens; 4800 tokens here)
    assert (ids_fused == ids_ref).float().mean().item() >= 0.995
    # the one-launch chain and the layer-at-a-time kernels are the same arithmetic, layer for layer
    with torch.no_grad():
        model.encoder_mode = "fused_bf16"
        model.fused_chain = False
        z_layers = model.encode(x)
        model.fused_chain, model.fused_projection = True, False
        z_chain = model.encode(x)
        model.fused_projection = True
        model.encoder_mode = "torch"
    torch.testing.assert_close(z_chain, z_layers, rtol=1e-5, atol=1e-6)
    # the fused projection rounds its operand h to bf16 (2^-9 relative per component, like every other layer's operand)
    assert (z_fused - z_layers).abs().max().item() <= 0.01 * z_layers.abs().max().item()
    # training mode / autograd never takes the fused path
    model.train()
    model.encoder_mode = "fused_bf16"
    assert not model._fused_ok(x)


def test_token_bias_gelu():
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(3)
    h = torch.randn(1000, 512, device=dev, generator=g) * 4      # pre-activations out to |x| ~ 20
    b = torch.randn(512, device=dev, generator=g)
    ref_h = h + b
    out = ops.token_bias_gelu(h, b)
    assert torch.equal(h, ref_h)
    ref = torch.nn.functional.gelu(ref_h)
    bound = 5e-5 + 3e-4 * ref_h.abs() + 2.0 ** -8 * ref.abs()     # MUFU.TANH (2^-11 on the tanh) + bf16 rounding
    assert bool(((out.float() - ref).abs() <= bound).all())


def test_fused_bf16_encoder_folds_eval_batchnorm():
    """The module's constructor default is batch_norm=True: in eval mode the fused path folds the BatchNorm affine maps
    into the layer weights and still tracks the fp32 encoder."""
    dev = _dev()
    torch.manual_seed(1)
    model = vqb200.VQVAEPatch(hidden_dim=256, input_dim=2, num_embeddings=64, embedding_dim=32, n_resblocks=2,
                              learning_rate=1e-3, patch_size=25, batch_norm=True).to(dev)
    with torch.no_grad():                      # non-trivial running statistics and affine parameters
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.8, 1.2); m.bias.normal_(0, 0.1)
    model.eval()
    x = torch.randn(200, 200, 2, device=dev)
    with torch.no_grad():
        model.encoder_mode = "torch"
        z_ref = model.encode(x)
        model.encoder_mode = "fused_bf16"
        assert model._fused_ok(x)
        z_fused = model.encode(x)
    assert (z_fused - z_ref).abs().max().item() <= 0.05 * z_ref.abs().max().item()


@pytest.mark.parametrize("B,L,C,P", [(37, 200, 2, 25), (64, 200, 2, 10), (5, 200, 2, 50), (3, 64, 3, 64)])
def test_patch_embed_kernel(B, L, C, P):
    """vqb_patch_embed against the reference's PatchEmbedding arithmetic (model/vq_vae_patch_embedd.py:13-17:
    permute, reshape, stride-P conv == linear per patch): h within 1e-5 relative (fp32, different summation order),
    the bf16 activation within the fused GELU's stated bound; ragged token tiles (B * T not a multiple of 32)."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(B * 1000 + P)
    x = torch.randn(B, L, C, device=dev, generator=g)
    conv = torch.nn.Conv1d(1, 512, kernel_size=P, stride=P)
    with torch.no_grad():        # the reference's own ops on the CPU in fp32 (cuDNN would run the conv in TF32 by default)
        ref = conv(x.cpu().permute(0, 2, 1).reshape(B, 1, -1)).permute(0, 2, 1).reshape(-1, 512).to(dev)   # (B*T, H)
    conv = conv.to(dev)
    h, a = ops.patch_embed(x, conv.weight, conv.bias, P)
    assert h.shape == ref.shape and a.shape == ref.shape and a.dtype == torch.bfloat16
    torch.testing.assert_close(h, ref, rtol=1e-5, atol=2e-6)
    act = torch.nn.functional.gelu(h)
    bound = 5e-5 + 3e-4 * h.abs() + 2.0 ** -8 * act.abs()
    assert bool(((a.float() - act).abs() <= bound).all())
    h2, a2 = ops.patch_embed(x, conv.weight, conv.bias, P, want_act=False)
    assert a2 is None and torch.equal(h, h2)


def _wide_model(patch_wide_golden, dev):
    """The fused-eligible fixture model (hidden_dim = 256): only the tensors the encode path reads were stored."""
    case = C.PATCH_WIDE_CASE
    name = case["name"]
    torch.manual_seed(1)
    model = vqb200.VQVAEPatch(hidden_dim=case["hidden_dim"], input_dim=case["input_dim"],
                              num_embeddings=case["num_embeddings"], embedding_dim=case["embedding_dim"],
                              n_resblocks=case["n_resblocks"], learning_rate=1e-3, dropout_p=0.0,
                              patch_size=case["patch_size"], seq_len=case["seq_len"],
                              batch_norm=case["batch_norm"], beta=case["beta"])
    sd = model.state_dict()
    for k in patch_wide_golden.files:
        if k.startswith(f"{name}/sd/"):
            sd[k[len(name) + 4:]] = torch.from_numpy(patch_wide_golden[k])
        elif k.startswith(f"{name}/centre/"):
            key = k[len(name) + 8:]
            wfull = sd[key].clone()
            wfull[:, :, 1] = torch.from_numpy(patch_wide_golden[k])
            sd[key] = wfull
    model.load_state_dict(sd, strict=True)
    model.encoder_mode = "torch"
    return model.to(dev).eval()


def test_encoder_modes_against_the_reference_fixture(patch_wide_golden):
    """a1-a3 against the UNMODIFIED reference (tests/golden/patch_wide_golden.npz: z_e and ids of a hidden_dim = 256
    model on 256 cycles).  fp32 encoder: ids equal except fp32 near-ties (1e-5 relative).  bf16-operand fused encoder
    (one-launch chain and layer-at-a-time): stated tolerance >= 99.8 % of the ids, z_e within 2 % of its range."""
    dev = _dev()
    case = C.PATCH_WIDE_CASE
    name = case["name"]
    model = _wide_model(patch_wide_golden, dev)
    x = torch.from_numpy(C.make_cycles(case)).to(dev)
    ref_z, ref_ids = patch_wide_golden[f"{name}/z_e"], patch_wide_golden[f"{name}/idx"].astype(np.int64)
    E = patch_wide_golden[f"{name}/sd/vector_quantization.embedding.weight"]
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        z = model.encode(x)
        ids = model.encode_ids(x).cpu().numpy().reshape(-1)
    np.testing.assert_allclose(z.cpu().numpy(), ref_z, rtol=1e-4, atol=1e-5)
    assert C.unexplained_mismatches(ref_z, E, ids, ref_ids, rel=1e-5) == 0
    assert (ids == ref_ids).mean() >= 0.999
    model.encoder_mode = "fused_bf16"
    with torch.no_grad():
        assert model._fused_ok(x)
    for chain in (True, False):
        model.fused_chain = chain
        with torch.no_grad():
            zf = model.encode(x)
            idf = model.encode_ids(x).cpu().numpy().reshape(-1)
        assert zf.shape == z.shape
        assert np.abs(zf.cpu().numpy() - ref_z).max() <= 0.02 * np.abs(ref_z).max()
        rate = (idf == ref_ids).mean()
        assert rate >= 0.998, (chain, rate)
    # fp32-faithful fused encoder (bf16 hi + lo pairs, three products per layer): the fp32 encoder's own tolerances
    model.encoder_mode = "fused_fp32"
    with torch.no_grad():
        zs = model.encode(x)
        ids_s = model.encode_ids(x).cpu().numpy().reshape(-1)
    np.testing.assert_allclose(zs.cpu().numpy(), ref_z, rtol=1e-4, atol=1e-4 * np.abs(ref_z).max())
    assert C.unexplained_mismatches(ref_z, E, ids_s, ref_ids, rel=1e-4) == 0
    assert (ids_s == ref_ids).mean() >= 0.999


@pytest.mark.parametrize("T,H,L", [(1, 512, 2), (127, 512, 4), (128 * 3 + 5, 512, 16), (128 * 150 + 77, 512, 16),
                                   (1000, 256, 2), (128 * 149, 256, 6)])
def test_encoder_chain_equals_layerwise(T, H, L):
    """vqb_encoder_chain against the layer-at-a-time vqb_token_linear sequence on the same operands: same bf16 operands,
    same accumulation order per layer, so the residual streams agree to fp32 rounding; ragged and multi-tile shapes."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(T + H + L)
    h0 = torch.randn(T, H, device=dev, generator=g)
    w = (torch.randn(L, H, H, device=dev, generator=g) * (1.0 / H) ** 0.5).to(torch.bfloat16)
    b = 0.1 * torch.randn(L, H, device=dev, generator=g)
    a0 = torch.nn.functional.gelu(h0).to(torch.bfloat16)
    ref = h0.clone()
    a, u = a0.clone(), torch.empty_like(a0)
    for i in range(L // 2):
        ops.token_linear(a, w[2 * i], b[2 * i], out=u, mode=0)
        ops.token_linear(u, w[2 * i + 1], b[2 * i + 1], h=ref, out=a if 2 * i + 2 < L else None, mode=1)
    out = ops.encoder_chain(a0, h0.clone(), w, b)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    # and against plain fp32 PyTorch on the same bf16-rounded operands (tolerance: bf16 rounding of the activations)
    hh = h0.clone()
    for i in range(L // 2):
        t = torch.nn.functional.gelu(hh).to(torch.bfloat16).float() @ w[2 * i].float().t() + b[2 * i]
        hh = hh + torch.nn.functional.gelu(t).to(torch.bfloat16).float() @ w[2 * i + 1].float().t() + b[2 * i + 1]
    torch.testing.assert_close(out, hh, rtol=2e-2, atol=2e-2 * L ** 0.5)


@pytest.mark.parametrize("T,H,L,D", [(1000, 512, 4, 32), (128 * 149 + 3, 512, 16, 32), (777, 256, 2, 8), (300, 512, 2, 64)])
def test_encoder_chain_fused_projection(T, H, L, D):
    """The final H -> D projection fused into the chain launch, both operands as bf16 hi + lo pairs (two K passes for h,
    two column groups for Wp): z_e = h_final Wp^T + bp to 2^-16 relative -- against the chain without projection
    followed by the fp32 product in PyTorch."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(T + H + L + D)
    h0 = torch.randn(T, H, device=dev, generator=g)
    w = (torch.randn(L, H, H, device=dev, generator=g) * (1.0 / H) ** 0.5).to(torch.bfloat16)
    b = 0.1 * torch.randn(L, H, device=dev, generator=g)
    wp = torch.randn(D, H, device=dev, generator=g) * (1.0 / H) ** 0.5
    bp = 0.1 * torch.randn(D, device=dev, generator=g)
    a0 = torch.nn.functional.gelu(h0).to(torch.bfloat16)
    h_final = ops.encoder_chain(a0, h0.clone(), w, b)
    rows = ops.projection_rows(wp)
    w_exact = rows[:D].float() + rows[64:64 + D].float()
    assert (w_exact - wp).abs().max().item() <= 2.0 ** -16 * wp.abs().max().item()
    ref = h_final @ w_exact.t() + bp
    stack = torch.cat([w.reshape(-1, H), rows]).contiguous()
    h_in = h0.clone()
    z = ops.encoder_chain(a0, h_in, stack, b, proj_bias=bp)
    assert z.shape == (T, D) and z.dtype == torch.float32
    torch.testing.assert_close(z, ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))


@pytest.mark.parametrize("hidden,n_res,patch,B", [(512, 8, 25, 300), (256, 2, 10, 129), (512, 1, 25, 7)])
def test_fully_fused_encoder_launch(hidden, n_res, patch, B):
    """Raw samples in, z_e out in one launch (patch embedding as the first GEMM, projection as the last, both with bf16
    hi + lo operand pairs) against the same chain fed by the fp32 patch-embedding kernel and followed by an fp32
    projection: the two differ by 2^-16-relative operand errors only."""
    dev = _dev()
    torch.manual_seed(hidden + patch)
    model = vqb200.VQVAEPatch(hidden_dim=hidden, input_dim=2, num_embeddings=64, embedding_dim=32, n_resblocks=n_res,
                              learning_rate=1e-3, patch_size=patch, batch_norm=False).to(dev).eval()
    x = torch.randn(B, 200, 2, device=dev)
    model.encoder_mode = "fused_bf16"
    with torch.no_grad():
        assert model._fused_ok(x)
        model.fused_patch_embed, model.fused_projection = False, False
        z_ref = model.encode(x)
        model.fused_projection = True
        z_proj = model.encode(x)
        model.fused_patch_embed = True
        z_all = model.encode(x)
    scale = z_ref.abs().max().item()
    assert z_all.shape == z_ref.shape == (B, 16 if patch == 25 else 40, 32)
    assert (z_proj - z_ref).abs().max().item() <= 2e-4 * scale       # hi + lo projection: fp32-faithful
    # the fused patch embedding perturbs h0 by 2^-16 relative; the bf16 roundings of the 2 * n_res layers then differ
    # on a few elements each, so the bound is the bf16-operand tolerance of the path, not 2^-16
    assert (z_all - z_ref).abs().max().item() <= 0.02 * scale
    assert ((z_all - z_ref).abs().mean() / z_ref.abs().mean()).item() <= 2e-3


# ---------------------------------------------------------------------------------------
# decoder half on the fused layer kernels
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_cycles,tpc", [(1, 16), (8, 16), (9, 16), (300, 16), (50, 8), (7, 32)])
@pytest.mark.parametrize("K,N", [(256, 256), (512, 512)])
def test_token_conv_three_taps_matches_conv1d(n_cycles, tpc, K, N):
    """vqb_token_conv (taps = 3) against torch's own Conv1d(k=3, pad=1) in fp32 on the same bf16-rounded operands: zero
    padding at BOTH ends of every cycle, tiles that end inside the tensor, modes 0 / 1, with and without the output GELU."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(n_cycles * 31 + tpc + K)
    T = n_cycles * tpc
    a = torch.randn(T, K, device=dev, generator=g).to(torch.bfloat16)
    w3 = (torch.randn(N, K, 3, device=dev, generator=g) * (2.0 / (3 * K)) ** 0.5).to(torch.bfloat16)     # Conv1d weight layout
    bias = 0.1 * torch.randn(N, device=dev, generator=g)
    w = w3.permute(0, 2, 1).reshape(N, 3 * K).contiguous()
    x = a.float().view(n_cycles, tpc, K).permute(0, 2, 1)                                                # (cycles, K, positions)
    conv = torch.nn.functional.conv1d(x, w3.float(), bias, padding=1).permute(0, 2, 1).reshape(T, N)
    out = ops.token_conv(a, w, bias, mode=0, taps=3, tokens_per_cycle=tpc)
    torch.testing.assert_close(out.float(), torch.nn.functional.gelu(conv), rtol=1.0 / 128, atol=2e-3)
    out_lin = ops.token_conv(a, w, bias, mode=0, taps=3, tokens_per_cycle=tpc, out_gelu=False)
    torch.testing.assert_close(out_lin.float(), conv, rtol=1.0 / 128, atol=2e-3)
    h = torch.randn(T, N, device=dev, generator=g)
    h0 = h.clone()
    nxt = torch.empty(T, N, dtype=torch.bfloat16, device=dev)
    ops.token_conv(a, w, bias, h=h, out=nxt, mode=1, taps=3, tokens_per_cycle=tpc, out_gelu=False)
    torch.testing.assert_close(h, h0 + conv, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(nxt.float(), h0 + conv, rtol=1.0 / 128, atol=2e-3)
    with pytest.raises(RuntimeError):                      # cycles must tile the 128-token tiles
        ops.token_conv(a[: 2 * 24].contiguous(), w, bias, mode=0, taps=3, tokens_per_cycle=24)


@pytest.mark.parametrize("n_cycles,tpc", [(1, 16), (9, 16), (300, 16), (50, 8)])
@pytest.mark.parametrize("K,N", [(256, 256), (512, 512)])
def test_token_conv_split_is_fp32_faithful(n_cycles, tpc, K, N):
    """vqb_token_conv_split (three taps, bf16 hi + lo operand pairs) against Conv1d(k=3, pad=1) in fp64 on the same fp32
    operands: within 2^-14 of sum |a||w| per element (measured ~100x tighter), padding at both ends of every cycle."""
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(n_cycles * 37 + tpc + K)
    T = n_cycles * tpc
    a32 = torch.randn(T, K, device=dev, generator=g)
    w3 = torch.randn(N, K, 3, device=dev, generator=g) * (2.0 / (3 * K)) ** 0.5
    bias = 0.1 * torch.randn(N, device=dev, generator=g)
    a = ops.token_pair(a32, gelu=False)
    w = ops.conv_pair(w3)
    assert w.shape == (N, 6 * K)
    x = a32.double().view(n_cycles, tpc, K).permute(0, 2, 1)
    conv = torch.nn.functional.conv1d(x, w3.double(), bias.double(), padding=1).permute(0, 2, 1).reshape(T, N)
    bound = 2.0 ** -14 * torch.nn.functional.conv1d(x.abs(), w3.double().abs(), None, padding=1).permute(0, 2, 1).reshape(T, N) + 1e-6
    out = ops.token_conv_split(a, w, bias, mode=0, taps=3, tokens_per_cycle=tpc)
    assert bool(((_pair_value(out) - torch.nn.functional.gelu(conv)).abs() <= bound).all())
    h = torch.randn(T, N, device=dev, generator=g)
    h0 = h.clone()
    nxt = torch.empty(T, 2 * N, dtype=torch.bfloat16, device=dev)
    ops.token_conv_split(a, w, bias, h=h, out=nxt, mode=1, taps=3, tokens_per_cycle=tpc, out_gelu=False)
    want = h0.double() + conv
    assert bool(((h.double() - want).abs() <= bound + 2.0 ** -22 * want.abs()).all())
    assert bool(((_pair_value(nxt) - h.double()).abs() <= 2.0 ** -16 * h.double().abs() + 1e-30).all())
    # the last transposed convolution on the pair: (hi + lo) w^T + bias, `group` runs of H channels per token
    for hdim, group in ((256, 2), (512, 1)):
        src = torch.randn(T, group * hdim, device=dev, generator=g)
        pr = ops.token_pair(src, gelu=False)
        w_out = torch.randn(5, hdim, device=dev, generator=g) * 0.1
        got = ops.token_out_proj_pair(pr, w_out, 0.25, group)
        ref = _pair_value(pr).view(T * group, hdim) @ w_out.double().t() + 0.25
        assert got.shape == (T * group, 5)
        torch.testing.assert_close(got.double(), ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("R,H,P", [(1, 512, 5), (1000, 512, 5), (333, 256, 2), (64, 512, 8)])
def test_token_out_proj_matches_fp32_reference(R, H, P):
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(R + H + P)
    a = torch.randn(R, H, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(P, H, device=dev, generator=g) * H ** -0.5
    out = ops.token_out_proj(a, w, 0.25)
    torch.testing.assert_close(out, a.float() @ w.t() + 0.25, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("hidden,n_res,batch_norm", [(256, 2, False), (512, 1, True), (256, 3, True)])
def test_fused_decoder_matches_the_torch_decoder(hidden, n_res, batch_norm):
    """decoder_mode = "fused_bf16" against the stock PyTorch modules in fp32 on the same weights (eval mode, BatchNorm with
    non-trivial running statistics): bf16 operands through 2 * n_res + 2 layers -- stated tolerance 2 % of the output's
    largest magnitude (measured: below 1 %)."""
    dev = _dev()
    torch.manual_seed(hidden + n_res)
    model = vqb200.VQVAEPatch(hidden_dim=hidden, input_dim=2, num_embeddings=64, embedding_dim=32, n_resblocks=n_res,
                              learning_rate=1e-3, batch_norm=batch_norm).to(dev)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    model.eval()
    z_q = 0.5 * torch.randn(37, 16, 32, device=dev)
    model.encoder_mode = model.decoder_mode = "torch"
    with torch.no_grad():
        want = model.decode(z_q)
        model.decoder_mode = "fused_bf16"
        assert model._fused_decoder_ok(z_q)
        got = model.decode(z_q)
        assert got.shape == want.shape == (37, 200, 2)
        err = (got - want).abs().max().item()
        assert err <= 0.02 * want.abs().max().item(), (err, want.abs().max().item())
        # the whole forward with both halves fused: same contract, reconstruction within the same tolerance
        x = torch.randn(37, 200, 2, device=dev)
        model.encoder_mode = "torch"
        loss0, xhat0, ppl0 = model(x)
        model.decoder_mode = "torch"
        loss1, xhat1, ppl1 = model(x)
        assert torch.equal(loss0, loss1) and torch.equal(ppl0, ppl1)
        assert (xhat0 - xhat1).abs().max().item() <= 0.02 * xhat1.abs().max().item()
        # the fp32-faithful form of the same launches (the module's default for inference): 2^-18 operand pairs through the
        # 2 n_res + 2 layers -- stated tolerance 1e-4 of the output's largest magnitude (measured ~1e-5)
        for mode in ("fused_fp32", "auto"):
            model.decoder_mode = mode
            got32 = model.decode(z_q)
            assert got32.shape == want.shape
            err32 = (got32 - want).abs().max().item()
            assert err32 <= 1e-4 * want.abs().max().item(), (mode, err32, want.abs().max().item())
        model.encoder_mode = model.decoder_mode = "auto"
        loss2, xhat2, ppl2 = model(x)
        assert (xhat2 - xhat1).abs().max().item() <= 2e-4 * xhat1.abs().max().item()
        assert loss2.item() == pytest.approx(loss1.item(), rel=1e-4)
    # training / autograd keep the PyTorch modules
    model.decoder_mode = "fused_bf16"
    model.train()
    assert not model._fused_decoder_ok(z_q)
