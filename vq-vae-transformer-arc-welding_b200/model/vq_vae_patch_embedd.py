"""VQ-VAE-Patch with the B200 encode-and-quantise path -- mirror of the reference's
``model/vq_vae_patch_embedd.py`` (same constructor, attributes, sub-module names and
therefore the same 80-key state dict, so reference checkpoints load strictly).

Encode half (the hot path):
  * PatchEmbedding  : channel-major patchify + linear P -> H            (:7-17)
  * CNNBlock        : per-token residual MLP.  The reference applies Conv1d(k=3, pad=1) to
                      length-1 slices in a 16-iteration Python loop (:103-111), where only the
                      centre tap W[:, :, 1] ever meets non-zero data; here the same weights are
                      applied to all tokens at once as dense layers (one GEMM per conv instead
                      of 16 tiny convolutions).  With BatchNorm in training mode the per-position
                      batch statistics of the reference are kept by looping like it does.
  * SepCNNBlock     : per-token linear H -> D, written as contiguous (B, T, D) rows (the reference returns
                      the same values as a permuted view of (B, D, T), :83-91)
  * VectorQuantizer : fused CUDA kernel; contiguous rows go straight to the tcgen05 path, any strided view
                      (e.g. the reference encoder's permuted output) is accepted too.
Decode half (:19-57, :142-147): stock PyTorch modules by default; `decoder_mode = "fused_bf16"` runs inference on the same
tcgen05 layer kernels (three-tap convolutions as shifted TMA boxes).
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import functional as F

from .autencoder_lightning_base import Autoencoder
from .vector_quantizer import VectorQuantizer


class PatchEmbedding(nn.Module):
    def __init__(self, patch_size, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv1d(1, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        """(B, L, C) -> (B, H, T) with T = L*C/patch: channel-major, all patches of channel 0
        first (:14-15).  A stride-P conv over the flattened signal is a linear map per patch."""
        b = x.shape[0]
        patches = x.permute(0, 2, 1).reshape(b, -1, self.patch_size)        # (B, T, P)
        tokens = F.linear(patches, self.proj.weight[:, 0, :], self.proj.bias)  # (B, T, H)
        return tokens.permute(0, 2, 1)


class PatchEmbeddingInverse(nn.Module):
    _KERNELS = {25: (5, 5), 10: (2, 5), 50: (10, 5)}   # (:24-45)

    def __init__(self, patch_size, embed_dim, input_dim):
        super().__init__()
        if patch_size not in self._KERNELS:
            raise NotImplementedError(f"Patch size not implemented: {patch_size}")
        k0, k1 = self._KERNELS[patch_size]
        self.patch_size = patch_size
        self.input_dim = input_dim
        self.proj = nn.Sequential(
            nn.ConvTranspose1d(embed_dim, embed_dim, kernel_size=k0, stride=k0),
            nn.BatchNorm1d(embed_dim),
            nn.GELU(),
            nn.ConvTranspose1d(embed_dim, 1, kernel_size=k1, stride=k1),
        )

    def forward(self, x):
        x = self.proj(x)
        return x.reshape(x.shape[0], -1, self.input_dim)   # (:56) not the inverse of the patchify; kept


class ResBlock(nn.Module):
    def __init__(self, channels: int, kernel_size: int = 3, stride: int = 1, padding: int = 1,
                 dropout_p: float = 0.1, batch_norm: bool = True):
        super().__init__()
        def norm():
            return nn.BatchNorm1d(channels) if batch_norm else nn.Identity()
        self.block = nn.Sequential(
            nn.GELU(),
            nn.Conv1d(channels, channels, kernel_size=kernel_size, stride=stride, padding=padding),
            norm(),
            nn.GELU(),
            nn.Conv1d(channels, channels, kernel_size=kernel_size, stride=stride, padding=padding),
            norm(),
            nn.Dropout(p=dropout_p),
        )
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        return x + self.block(x)

    def _centre_tap_ok(self) -> bool:
        return self.stride == 1 and self.kernel_size == 2 * self.padding + 1

    def forward_tokens(self, t):
        """Same block applied to independent tokens t: (..., C).  On a length-1 sequence with
        zero padding only the centre tap of each conv sees data."""
        blk = self.block
        c = self.padding
        h = F.gelu(t)
        h = F.linear(h, blk[1].weight[:, :, c], blk[1].bias)
        h = _token_norm(blk[2], h)
        h = F.gelu(h)
        h = F.linear(h, blk[4].weight[:, :, c], blk[4].bias)
        h = _token_norm(blk[5], h)
        return t + blk[6](h)


def _token_norm(norm, t):
    """BatchNorm1d in eval mode is a per-channel affine map; Identity passes through."""
    if isinstance(norm, nn.Identity):
        return t
    return F.batch_norm(t.reshape(-1, t.shape[-1]), norm.running_mean, norm.running_var,
                        norm.weight, norm.bias, False, 0.0, norm.eps).reshape(t.shape)


class SepCNNBlock(nn.Module):
    def __init__(self, hidden_dim: int, embedding_dim: int) -> None:
        super().__init__()
        self.shared_conv = nn.Conv1d(hidden_dim, embedding_dim, kernel_size=1, stride=1, padding=0)

    def forward(self, x):
        """(B, H, T) -> (B, T, D).  The reference returns the same values as a permuted view of a
        (B, D, T) tensor (:91); here the per-token projection writes (B, T, D) rows directly, which is
        the layout the tcgen05 quantiser streams with TMA (a permuted view would cost a packing copy)."""
        tokens = x.permute(0, 2, 1)                # a view of the contiguous token rows CNNBlock produced
        return F.linear(tokens, self.shared_conv.weight[:, :, 0], self.shared_conv.bias)


class CNNBlock(nn.Module):
    def __init__(self, embed_dim: int, seperate: bool = True, kernel_size: int = 3, stride: int = 1,
                 padding: int = 1, dropout_p: float = 0.1, batch_norm: bool = True, n_resblocks: int = 1):
        super().__init__()
        self.seperate = seperate
        self.shared_conv = nn.Sequential(
            *[ResBlock(channels=embed_dim, kernel_size=kernel_size, stride=stride, padding=padding,
                       dropout_p=dropout_p, batch_norm=batch_norm) for _ in range(n_resblocks)])
        self._has_bn = batch_norm

    def forward(self, x):
        if not self.seperate:
            return self.shared_conv(x)
        per_position_stats = self._has_bn and self.training
        if per_position_stats or not all(b._centre_tap_ok() for b in self.shared_conv):
            # training-mode BatchNorm normalises each position with its own batch statistics and
            # updates the running statistics once per position: do exactly what the reference does
            cols = [self.shared_conv(x[:, :, i].unsqueeze(2)) for i in range(x.shape[2])]
            return torch.cat(cols, dim=2)
        t = x.permute(0, 2, 1)                     # (B, T, H) tokens
        for blk in self.shared_conv:
            t = blk.forward_tokens(t)
        return t.permute(0, 2, 1)


class VQVAEPatch(Autoencoder):

    def __init__(self, hidden_dim: int, input_dim: int, num_embeddings: int, embedding_dim: int,
                 n_resblocks: int, learning_rate: float, dropout_p: float = 0.1, patch_size: int = 25,
                 seq_len: int = 200, batch_norm: bool = True, beta: float = 0.25,
                 use_improved_vq: bool = False, kmeans_iters: int = 0, threshold_ema_dead_code: int = 2):
        super().__init__(hidden_dim=hidden_dim, input_dim=input_dim, num_embeddings=num_embeddings,
                         embedding_dim=embedding_dim, n_resblocks=n_resblocks, learning_rate=learning_rate,
                         seq_len=seq_len, dropout_p=dropout_p)
        if use_improved_vq:
            # ResidualVQLightning wraps the un-vendored `vector-quantize-pytorch` package
            # (model/vector_quantizer.py:9-56); out of scope, parity unpinned.
            raise NotImplementedError("use_improved_vq=True (ResidualVQ) is outside the B200 hot path")
        self.patch_embed = PatchEmbedding(patch_size=patch_size, embed_dim=hidden_dim)
        self.encoder = nn.Sequential(
            CNNBlock(embed_dim=hidden_dim, n_resblocks=n_resblocks, dropout_p=dropout_p, batch_norm=batch_norm),
            SepCNNBlock(hidden_dim=hidden_dim, embedding_dim=embedding_dim),
        )
        self.vector_quantization = VectorQuantizer(n_e=num_embeddings, e_dim=embedding_dim, beta=beta)
        self.decoder = nn.Sequential(
            nn.Conv1d(embedding_dim, hidden_dim, kernel_size=1, stride=1, padding=0),
            CNNBlock(embed_dim=hidden_dim, seperate=False, n_resblocks=n_resblocks, dropout_p=dropout_p,
                     batch_norm=batch_norm),
        )
        self.reverse_patch_embed = PatchEmbeddingInverse(patch_size=patch_size, embed_dim=hidden_dim,
                                                         input_dim=input_dim)
        self.enc_out_len = seq_len // patch_size * input_dim
        self.patch_size = patch_size
        self.apply(self.weights_init)

    # ---- encode half: the hot path ------------------------------------------------------
    #: "auto" (default): "fused_fp32" whenever the call qualifies (CUDA fp32 input, eval mode, no autograd, hidden size a
    #: multiple of 256, centre-tap residual blocks), the PyTorch layers otherwise (training, CPU, narrow models);
    #: "torch": stock PyTorch layers in the ambient matmul precision (fp32-faithful; the only mode that trains);
    #: "fused_fp32": inference on the tcgen05 layer kernel with bf16 hi + lo operand pairs (three products per layer, fp32
    #: accumulation, erf GELU): operand precision 2^-16, ids equal to the fp32 layers' except on ~1e-5 of the tokens;
    #: "fused_bf16": the residual blocks run on the fused tcgen05 layer kernel (csrc/tok_linear.cu) with bf16
    #: operands / fp32 accumulation -- the operand precision the reference itself selects with
    #: torch.set_float32_matmul_precision('medium') (train_*.py), without the element-wise passes.
    encoder_mode = "auto"
    #: in "fused_bf16" mode: all residual blocks in one launch (csrc/enc_chain.cu) instead of one launch per layer
    fused_chain = True
    #: ... with the final H -> D projection fused into the same launch (operand bf16(h), weights exact to 2^-17)
    fused_projection = True
    #: ... and the patch embedding as its first GEMM (samples and weights as bf16 hi + lo pairs: fp32-accurate to 2^-16)
    fused_patch_embed = True

    def encode(self, x):
        """x (B, seq_len, input_dim) -> z_e (B, T, D)."""
        if self.encoder_mode in ("auto", "fused_fp32") and self._fused_ok(x) and self.encoder[1].shared_conv.out_channels <= 256:
            torch.cuda.nvtx.range_push("vqb200.encode_fused_fp32")
            try:
                return self.encode_fused_fp32(x)
            finally:
                torch.cuda.nvtx.range_pop()
        if self.encoder_mode == "fused_bf16" and self._fused_ok(x):
            torch.cuda.nvtx.range_push("vqb200.encode_fused_bf16")
            try:
                return self.encode_fused_bf16(x)
            finally:
                torch.cuda.nvtx.range_pop()
        return self.encoder(self.patch_embed(x))

    def _fused_ok(self, x) -> bool:
        cnn = self.encoder[0]
        # (BatchNorm layers are fine in eval mode: a per-channel affine map, folded into the layer's weights below)
        return (x.is_cuda and x.dtype == torch.float32 and not self.training and not torch.is_grad_enabled()
                and cnn.seperate and all(b._centre_tap_ok() for b in cnn.shared_conv)
                and self.patch_embed.proj.out_channels % 256 == 0)

    def _fused_weights(self):
        """bf16 centre-tap weights of the residual blocks (eval-mode BatchNorm folded in: y = (W x + b - mean) *
        gamma / sqrt(var + eps) + beta is again a linear layer), rebuilt when a parameter or running statistic changes."""
        blocks = list(self.encoder[0].shared_conv)
        tracked = [t for blk in blocks for m in (blk.block[1], blk.block[2], blk.block[4], blk.block[5])
                   for t in list(m.parameters()) + list(m.buffers())] + list(self.encoder[1].shared_conv.parameters()) \
            + list(self.patch_embed.proj.parameters())
        key = tuple((t.data_ptr(), t._version) for t in tracked)
        cache = getattr(self, "_fused_cache", None)
        if cache is None or cache[0] != key:
            def fold(conv, norm, c):
                w, b = conv.weight[:, :, c].float(), conv.bias.float()
                if isinstance(norm, nn.BatchNorm1d):
                    scale = norm.weight.float() / torch.sqrt(norm.running_var.float() + norm.eps)
                    w, b = w * scale[:, None], (b - norm.running_mean.float()) * scale + norm.bias.float()
                return w.to(torch.bfloat16).contiguous(), b.contiguous()
            ws = []
            with torch.no_grad():
                for blk in blocks:
                    w1, b1 = fold(blk.block[1], blk.block[2], blk.padding)
                    w2, b2 = fold(blk.block[4], blk.block[5], blk.padding)
                    ws.append((w1, b1, w2, b2))
            # the same layers stacked for the one-launch chain kernel: (L, H, H) bf16 and (L, H) fp32
            stack_w = torch.stack([w for blk in ws for w in (blk[0], blk[2])]).contiguous()
            stack_b = torch.stack([b for blk in ws for b in (blk[1], blk[3])]).contiguous()
            # ... and, behind them, the final projection as an exact bf16 hi + lo pair (ops.projection_rows)
            from .. import ops
            proj = self.encoder[1].shared_conv
            with torch.no_grad():
                wp = proj.weight[:, :, 0].float()
                stack_wp = None
                if wp.shape[0] <= 64 and wp.shape[0] % 4 == 0:
                    stack_wp = torch.cat([stack_w.reshape(-1, stack_w.shape[-1]), ops.projection_rows(wp)]).contiguous()
                # ... and, last, the patch embedding as a bf16 hi + lo pair (ops.patch_rows) for the fully fused launch
                stack_all = None
                pe = self.patch_embed.proj
                if stack_wp is not None and pe.kernel_size[0] <= 32:
                    stack_all = torch.cat([stack_wp, ops.patch_rows(pe.weight[:, 0, :].float(), stack_w.shape[-1])]).contiguous()
            cache = (key, ws, stack_w, stack_b, stack_wp, stack_all)
            object.__setattr__(self, "_fused_cache", cache)
        return cache[1]

    def _fused_stack(self):
        self._fused_weights()
        return self._fused_cache[2], self._fused_cache[3], self._fused_cache[4], self._fused_cache[5]

    def encode_fused_bf16(self, x):
        """The encoder with its 16 hidden x hidden layers on the fused kernel: tokens (B*T, H) stay row-major,
        the residual stream h is fp32 and updated in place, activations between the layers are bf16."""
        from .. import ops
        b = x.shape[0]
        pe = self.patch_embed
        hidden = pe.proj.out_channels
        chain_ok = self.fused_chain and hidden in (256, 512) and len(self.encoder[0].shared_conv) >= 1
        if chain_ok and self.fused_projection and self.fused_patch_embed and x.is_contiguous():
            stack_all = self._fused_stack()[3]
            if stack_all is not None:
                # raw samples in, z_e out: patch embedding, every residual block and the projection in ONE launch
                # (vqb_patch_split writes the 128-byte-per-token operand of the first GEMM)
                proj = self.encoder[1].shared_conv
                z_e = ops.encoder_chain(ops.patch_split(x, pe.patch_size), None, stack_all, self._fused_stack()[1],
                                        proj_bias=proj.bias.detach().float().contiguous(),
                                        pre_bias=pe.proj.bias.detach().float().contiguous())
                return z_e.view(b, -1, z_e.shape[-1])
        # (the patch embedding stays fp32 arithmetic: K = 25 is 0.3 % of the FLOPs, and feeding the raw signal as bf16
        # through vqb_token_linear mode 2 was measured to cost index matches -- 99.86 % -> 99.59 %)
        if pe.proj.out_channels == 512 and pe.patch_size <= 64 and x.is_contiguous():
            # one kernel: patchify + linear + bias -> h fp32, a = bf16(gelu(h))  (vqb_patch_embed)
            h, a = ops.patch_embed(x, pe.proj.weight, pe.proj.bias, pe.patch_size)
        else:
            patches = x.permute(0, 2, 1).reshape(-1, pe.patch_size)                   # (B*T, P)
            h = torch.matmul(patches, pe.proj.weight[:, 0, :].t())                    # (B*T, H) fp32, bias added below
            a = ops.token_bias_gelu(h, pe.proj.bias)                                   # h += b; a = bf16(gelu(h)), one pass
        if chain_ok:
            # every residual block in ONE launch: the token tile never leaves the SM between the layers (vqb_encoder_chain)
            stack_w, stack_b, stack_wp, _ = self._fused_stack()
            proj = self.encoder[1].shared_conv
            if self.fused_projection and stack_wp is not None:
                # ... and the final projection as one more GEMM on the resident tile: z_e comes straight out of the chain
                z_e = ops.encoder_chain(a, h, stack_wp, stack_b, proj_bias=proj.bias.detach().float().contiguous())
            else:
                ops.encoder_chain(a, h, stack_w, stack_b)
                z_e = F.linear(h, proj.weight[:, :, 0], proj.bias)
            return z_e.view(b, -1, z_e.shape[-1])
        u = torch.empty_like(a)
        blocks = self._fused_weights()
        for i, (w1, b1, w2, b2) in enumerate(blocks):
            ops.token_linear(a, w1, b1, out=u, mode=0)                                 # u = gelu(W1 a + b1)
            ops.token_linear(u, w2, b2, h=h, out=a if i + 1 < len(blocks) else None, mode=1)   # h += W2 u + b2; a = gelu(h)
        proj = self.encoder[1].shared_conv
        z_e = F.linear(h, proj.weight[:, :, 0], proj.bias)
        return z_e.view(b, -1, z_e.shape[-1])

    def _split_weights(self):
        """Operands of the fp32-faithful fused encoder: every layer's centre tap (eval-mode BatchNorm folded in, in fp32) as a
        bf16 hi + lo pair (ops.bf16_pair), the projection likewise on 256 zero-padded rows, the patch embedding as the
        three-product operand of one bf16 GEMM.  Rebuilt when a parameter or running statistic changes."""
        from .. import ops
        blocks = list(self.encoder[0].shared_conv)
        tracked = [t for blk in blocks for m in (blk.block[1], blk.block[2], blk.block[4], blk.block[5])
                   for t in list(m.parameters()) + list(m.buffers())] + list(self.encoder[1].shared_conv.parameters()) \
            + list(self.patch_embed.proj.parameters())
        key = tuple((t.data_ptr(), t._version) for t in tracked)
        cache = getattr(self, "_split_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]

        def fold(conv, norm, c):
            w, b = conv.weight[:, :, c].float(), conv.bias.float()
            if isinstance(norm, nn.BatchNorm1d):
                scale = norm.weight.float() / torch.sqrt(norm.running_var.float() + norm.eps)
                w, b = w * scale[:, None], (b - norm.running_mean.float()) * scale + norm.bias.float()
            return ops.bf16_pair(w), b.contiguous()

        with torch.no_grad():
            layers = []
            for blk in blocks:
                layers.append(fold(blk.block[1], blk.block[2], blk.padding) + fold(blk.block[4], blk.block[5], blk.padding))
            proj = self.encoder[1].shared_conv
            d, hidden = proj.weight.shape[0], proj.weight.shape[1]
            wp = torch.zeros(256, hidden, device=proj.weight.device)
            wp[:d] = proj.weight[:, :, 0].float()
            bp = torch.zeros(256, device=proj.weight.device)
            bp[:d] = proj.bias.float()
            pe = self.patch_embed.proj
            w_pe = None
            if pe.kernel_size[0] <= 32:
                # [a_hi | a_lo | a_hi | 0] (T, 128) against [w_hi | w_hi | w_lo | 0]: a_hi w_hi + a_lo w_hi + a_hi w_lo
                wpe = pe.weight[:, 0, :].float()
                hi = wpe.to(torch.bfloat16)
                lo = (wpe - hi.float()).to(torch.bfloat16)
                w_pe = torch.zeros(hidden, 128, dtype=torch.bfloat16, device=wpe.device)
                w_pe[:, :wpe.shape[1]] = hi
                w_pe[:, 32:32 + wpe.shape[1]] = hi
                w_pe[:, 64:64 + wpe.shape[1]] = lo
            ops_ = dict(layers=layers, wp=ops.bf16_pair(wp), bp=bp, d=d, w_pe=w_pe, b_pe=pe.bias.float().contiguous())
        object.__setattr__(self, "_split_cache", (key, ops_))
        return ops_

    def encode_fused_fp32(self, x):
        """The encoder on the tcgen05 layer kernel in its fp32-faithful form (vqb_token_linear_split): activations and
        weights travel as bf16 hi + lo pairs (2^-17 each), three products per layer accumulate in fp32, GELU in the erf
        form, fp32 residual stream.  ids equal the fp32 PyTorch encoder's except where two codes are within ~1e-5 of
        each other (tests/test_gpu_encoder.py); ~3x the tensor work of `fused_bf16`, one launch per layer."""
        from .. import ops
        b = x.shape[0]
        pe = self.patch_embed
        hidden = pe.proj.out_channels
        w = self._split_weights()
        if hidden == 512 and pe.patch_size <= 64 and x.is_contiguous():
            h, _ = ops.patch_embed(x, pe.proj.weight, pe.proj.bias, pe.patch_size, want_act=False)   # fp32 FMA arithmetic
        elif w["w_pe"] is not None and x.is_contiguous():
            a0 = ops.patch_split(x, pe.patch_size)                                    # (B*T, 64) = [hi | lo]
            a0 = torch.cat([a0, a0[:, :32], torch.zeros_like(a0[:, :32])], dim=1)     # (B*T, 128)
            h = torch.empty(a0.shape[0], hidden, dtype=torch.float32, device=x.device)
            ops.token_linear(a0, w["w_pe"], w["b_pe"], h=h, mode=2)
        else:
            patches = x.permute(0, 2, 1).reshape(-1, pe.patch_size)
            h = torch.addmm(pe.proj.bias, patches, pe.proj.weight[:, 0, :].t())
        a = ops.token_pair(h, gelu=True)                                              # (B*T, 2H) pair of gelu(h)
        u = torch.empty_like(a)
        layers = w["layers"]
        for i, (w1, b1, w2, b2) in enumerate(layers):
            ops.token_linear_split(a, w1, b1, out=u, mode=0)                          # u = gelu(W1 a + b1)
            # h += W2 u + b2; a = pair(gelu(h)) for the next block, pair(h) for the projection after the last one
            ops.token_linear_split(u, w2, b2, h=h, out=a, mode=1, out_gelu=i + 1 < len(layers))
        if not layers:
            ops.token_pair(h, gelu=False, out=a)
        z = torch.empty(h.shape[0], 256, dtype=torch.float32, device=x.device)
        ops.token_linear_split(a, w["wp"], w["bp"], h=z, mode=2, out_gelu=False)                      # z[:, :D] = Wp h + bp
        return z[:, :w["d"]].contiguous().view(b, -1, w["d"])

    def encode_ids(self, x):
        """Token ids (B, T) int64 -- the encode call of dataloader/latentspace_dataloader.py:154-161
        without z_q, loss, one-hot or autograd."""
        with torch.no_grad():
            z_e = self.encode(x)
            return self.vector_quantization.encode_indices(z_e).view(x.shape[0], -1)

    # ---- decode half ---------------------------------------
    #: "torch": stock PyTorch modules (training, fp32-faithful);
    #: "fused_bf16": inference on the tcgen05 layer kernels -- the 1x1 input convolution, every three-tap convolution
    #: of the residual blocks and the first transposed convolution as vqb_token_conv launches on row-major tokens
    #: (bf16 operands, fp32 accumulation, fp32 residual stream, bias / eval-mode BatchNorm / GELU / residual in the
    #: epilogues), the last transposed convolution as vqb_token_out_proj.  Selected like the encoder's mode.
    #: "fused_fp32": the same launches in their fp32-faithful form (bf16 hi + lo operand pairs, three products per layer,
    #: erf GELU: vqb_token_conv_split / vqb_token_out_proj_pair); "auto" (default) selects it for inference calls that
    #: qualify, like the encoder's.
    decoder_mode = "auto"

    def decode(self, z_q):
        """z_q (B, T, D) -> x_hat (B, seq_len, input_dim)   (:164-165)."""
        if self.decoder_mode in ("auto", "fused_fp32") and self._fused_decoder_ok(z_q):
            torch.cuda.nvtx.range_push("vqb200.decode_fused_fp32")
            try:
                return self.decode_fused_fp32(z_q)
            finally:
                torch.cuda.nvtx.range_pop()
        if self.decoder_mode == "fused_bf16" and self._fused_decoder_ok(z_q):
            torch.cuda.nvtx.range_push("vqb200.decode_fused_bf16")
            try:
                return self.decode_fused_bf16(z_q)
            finally:
                torch.cuda.nvtx.range_pop()
        return self.reverse_patch_embed(self.decoder(z_q.permute(0, 2, 1)))

    def _fused_decoder_ok(self, z_q) -> bool:
        cnn = self.decoder[1]
        rp = self.reverse_patch_embed.proj
        hidden = self.decoder[0].out_channels
        t = z_q.shape[1] if z_q.dim() == 3 else 0
        return (z_q.is_cuda and z_q.dtype == torch.float32 and not self.training and not torch.is_grad_enabled()
                and z_q.dim() == 3 and z_q.shape[2] <= 64 and t >= 1 and 128 % t == 0
                and hidden in (256, 512) and not cnn.seperate
                and all(b._centre_tap_ok() and b.kernel_size == 3 for b in cnn.shared_conv)
                and rp[0].stride == rp[0].kernel_size and rp[3].stride == rp[3].kernel_size and rp[3].kernel_size[0] <= 8)

    def _fused_decoder_weights(self):
        """Operands of the fused decoder, rebuilt when a parameter or running statistic changes: every convolution as a
        (out, taps * in) bf16 matrix with eval-mode BatchNorm folded in, biases fp32."""
        tracked = [t for m in (self.decoder, self.reverse_patch_embed) for t in list(m.parameters()) + list(m.buffers())]
        key = tuple((t.data_ptr(), t._version) for t in tracked)
        cache = getattr(self, "_fused_dec_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]

        def bn_fold(w, b, norm):          # w (out, ...), b (out): y = norm(conv(x)) is again a convolution
            if isinstance(norm, nn.BatchNorm1d):
                scale = norm.weight.float() / torch.sqrt(norm.running_var.float() + norm.eps)
                w = w * scale.view(-1, *([1] * (w.dim() - 1)))
                b = (b - norm.running_mean.float()) * scale + norm.bias.float()
            return w, b

        with torch.no_grad():
            conv0 = self.decoder[0]
            hidden, d = conv0.out_channels, conv0.in_channels
            w0 = torch.zeros(hidden, 64, device=conv0.weight.device)
            w0[:, :d] = conv0.weight[:, :, 0].float()
            blocks = []
            for blk in self.decoder[1].shared_conv:
                pair = []
                for conv, norm in ((blk.block[1], blk.block[2]), (blk.block[4], blk.block[5])):
                    w, b = bn_fold(conv.weight.float(), conv.bias.float(), norm)          # (out, in, 3)
                    pair += [w.permute(0, 2, 1).reshape(hidden, -1).to(torch.bfloat16).contiguous(), b.contiguous()]
                blocks.append(tuple(pair))
            rp = self.reverse_patch_embed.proj
            k0 = rp[0].kernel_size[0]
            # ConvTranspose1d(H, H, k0, stride k0): y[o, k0 t + j] = sum_c x[c, t] w[c, o, j] + b[o] -- a linear layer
            # H -> k0 * H on tokens, row (j, o); BatchNorm acts per output channel o
            wt = rp[0].weight.float().permute(2, 1, 0)                                    # (j, o, c)
            bt = rp[0].bias.float().expand(k0, -1)
            if isinstance(rp[1], nn.BatchNorm1d):
                scale = rp[1].weight.float() / torch.sqrt(rp[1].running_var.float() + rp[1].eps)
                wt = wt * scale.view(1, -1, 1)
                bt = (bt - rp[1].running_mean.float()) * scale + rp[1].bias.float()
            w_up = wt.reshape(k0 * hidden, hidden).to(torch.bfloat16).contiguous()
            b_up = bt.reshape(-1).contiguous()
            # ConvTranspose1d(H, 1, k1, stride k1): k1 samples per row
            w_out = rp[3].weight.float()[:, 0, :].t().contiguous()                        # (k1, H)
            b_out = float(rp[3].bias.float().item())
            ops_ = dict(w0=w0.to(torch.bfloat16).contiguous(), b0=conv0.bias.float().contiguous(), blocks=blocks,
                        w_up=w_up, b_up=b_up, w_out=w_out, b_out=b_out, k0=k0, hidden=hidden, d=d)
        object.__setattr__(self, "_fused_dec_cache", (key, ops_))
        return ops_

    def decode_fused_bf16(self, z_q):
        """Decoder + PatchEmbeddingInverse on the fused layer kernels (:142-147, :19-57): tokens (B*T, H) stay row-major,
        the residual stream h is fp32, activations between the layers are bf16."""
        from .. import ops
        w = self._fused_decoder_weights()
        b, t, d = z_q.shape
        n = b * t
        a0 = torch.zeros(n, 64, dtype=torch.bfloat16, device=z_q.device)
        a0[:, :d] = z_q.reshape(n, d)
        h = torch.empty(n, w["hidden"], dtype=torch.float32, device=z_q.device)
        a = torch.empty(n, w["hidden"], dtype=torch.bfloat16, device=z_q.device)
        u = torch.empty_like(a)
        # h = W0 z_q + b0; a = gelu(h)   (Conv1d(D, H, 1), then the first block's leading GELU)
        ops.token_conv(a0, w["w0"], w["b0"], h=h, out=a, mode=2, taps=1, tokens_per_cycle=1, out_gelu=bool(w["blocks"]))
        for i, (w1, b1, w2, b2) in enumerate(w["blocks"]):
            last = i + 1 == len(w["blocks"])
            ops.token_conv(a, w1, b1, out=u, mode=0, taps=3, tokens_per_cycle=t)                       # u = gelu(conv1(a))
            ops.token_conv(u, w2, b2, h=h, out=a, mode=1, taps=3, tokens_per_cycle=t, out_gelu=not last)   # h += conv2(u)
        # first transposed convolution (+ BatchNorm + GELU): H -> k0 * H per token = k0 rows of H
        up = ops.token_conv(a, w["w_up"], w["b_up"], mode=0, taps=1, tokens_per_cycle=1)
        x = ops.token_out_proj(up.view(n * w["k0"], w["hidden"]), w["w_out"], w["b_out"])           # (n * k0, k1)
        return x.view(b, -1, self.reverse_patch_embed.input_dim)

    def _split_decoder_weights(self):
        """Operands of the fp32-faithful fused decoder (bf16 hi + lo pairs of the BatchNorm-folded fp32 weights), rebuilt when a
        parameter or running statistic changes."""
        from .. import ops
        tracked = [t for m in (self.decoder, self.reverse_patch_embed) for t in list(m.parameters()) + list(m.buffers())]
        key = tuple((t.data_ptr(), t._version) for t in tracked)
        cache = getattr(self, "_split_dec_cache", None)
        if cache is not None and cache[0] == key:
            return cache[1]

        def bn_fold(w, b, norm):
            if isinstance(norm, nn.BatchNorm1d):
                scale = norm.weight.float() / torch.sqrt(norm.running_var.float() + norm.eps)
                w = w * scale.view(-1, *([1] * (w.dim() - 1)))
                b = (b - norm.running_mean.float()) * scale + norm.bias.float()
            return w, b

        with torch.no_grad():
            conv0 = self.decoder[0]
            hidden, d = conv0.out_channels, conv0.in_channels
            w0 = torch.zeros(hidden, 64, device=conv0.weight.device)
            w0[:, :d] = conv0.weight[:, :, 0].float()
            blocks = []
            for blk in self.decoder[1].shared_conv:
                pair = []
                for conv, norm in ((blk.block[1], blk.block[2]), (blk.block[4], blk.block[5])):
                    w, b = bn_fold(conv.weight.float(), conv.bias.float(), norm)          # (out, in, 3)
                    pair += [ops.conv_pair(w), b.contiguous()]
                blocks.append(tuple(pair))
            rp = self.reverse_patch_embed.proj
            k0 = rp[0].kernel_size[0]
            wt = rp[0].weight.float().permute(2, 1, 0)                                    # (j, o, c)
            bt = rp[0].bias.float().expand(k0, -1)
            if isinstance(rp[1], nn.BatchNorm1d):
                scale = rp[1].weight.float() / torch.sqrt(rp[1].running_var.float() + rp[1].eps)
                wt = wt * scale.view(1, -1, 1)
                bt = (bt - rp[1].running_mean.float()) * scale + rp[1].bias.float()
            ops_ = dict(w0=ops.bf16_pair(w0), b0=conv0.bias.float().contiguous(), blocks=blocks,
                        w_up=ops.bf16_pair(wt.reshape(k0 * hidden, hidden).contiguous()), b_up=bt.reshape(-1).contiguous(),
                        w_out=rp[3].weight.float()[:, 0, :].t().contiguous(), b_out=float(rp[3].bias.float().item()),
                        k0=k0, hidden=hidden, d=d)
        object.__setattr__(self, "_split_dec_cache", (key, ops_))
        return ops_

    def decode_fused_fp32(self, z_q):
        """decode_fused_bf16's launches in their fp32-faithful form: activations and weights as bf16 hi + lo pairs, three tcgen05
        products per layer in the fp32 accumulator, erf GELU, fp32 residual stream (vqb_token_conv_split); the last transposed
        convolution reads the pair's two halves (vqb_token_out_proj_pair)."""
        from .. import ops
        w = self._split_decoder_weights()
        b, t, d = z_q.shape
        n = b * t
        z64 = torch.zeros(n, 64, dtype=torch.float32, device=z_q.device)
        z64[:, :d] = z_q.reshape(n, d)
        a0 = ops.token_pair(z64, gelu=False)                                            # (n, 128)
        h = torch.empty(n, w["hidden"], dtype=torch.float32, device=z_q.device)
        a = torch.empty(n, 2 * w["hidden"], dtype=torch.bfloat16, device=z_q.device)
        u = torch.empty_like(a)
        ops.token_conv_split(a0, w["w0"], w["b0"], h=h, out=a, mode=2, taps=1, tokens_per_cycle=1, out_gelu=bool(w["blocks"]))
        for i, (w1, b1, w2, b2) in enumerate(w["blocks"]):
            last = i + 1 == len(w["blocks"])
            ops.token_conv_split(a, w1, b1, out=u, mode=0, taps=3, tokens_per_cycle=t)
            ops.token_conv_split(u, w2, b2, h=h, out=a, mode=1, taps=3, tokens_per_cycle=t, out_gelu=not last)
        up = ops.token_conv_split(a, w["w_up"], w["b_up"], mode=0, taps=1, tokens_per_cycle=1)      # (n, 2 * k0 * H)
        x = ops.token_out_proj_pair(up, w["w_out"], w["b_out"], w["k0"])                           # (n * k0, k1)
        return x.view(b, -1, self.reverse_patch_embed.input_dim)

    def forward(self, x):
        z_e = self.encode(x)
        vq = self.vector_quantization
        # (the (N, n_e) one-hot is dropped here, :161 of the reference: the fused quantiser does not write it)
        embedding_loss, z_q, perplexity, _, _ = vq(z_e, need_one_hot=False) if isinstance(vq, VectorQuantizer) else vq(z_e)
        x_hat = self.decode(z_q)
        return embedding_loss, x_hat, perplexity
