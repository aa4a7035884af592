/*
 * vqb200.h -- C ABI of the B200-native vector-quantisation hot path.
 *
 * Drop-in boundary for tmdt-buw/VQ-VAE-Transformer-Arc-Welding.  The reference
 * has no FFI of its own (it is pure PyTorch); the boundary it exposes for this
 * path is the Python class VectorQuantizer (model/vector_quantizer.py:59-131).
 * Each entry point below replaces one piece of that class and is what a ctypes
 * binding inside the reference would call (INTEGRATION.md shows the stub):
 *
 *   vqb_forward   <- VectorQuantizer.forward                model/vector_quantizer.py:76-119
 *   vqb_backward  <- autograd of :103-111 (straight-through + commitment/codebook loss)
 *   vqb_gather    <- VectorQuantizer.get_embedding_from_one_hot          :121-131
 *   vqb_one_hot   <- the (N, n_e) `min_encodings` matrix                 :98-100
 *   vqb_encode_host / vqb_host_* <- the host-side encode call of
 *                    dataloader/latentspace_dataloader.py:154-161,225-238
 *                    (host buffers in, ids out, copies inside the call)
 *
 * Conventions
 *   - Plain C: pointers and sizes only, no torch types, no C++ exceptions.
 *   - Device entry points take DEVICE pointers owned by the caller (for the
 *     PyTorch binding: tensors from the caching allocator).  The library never
 *     allocates or frees device memory on those paths, never synchronises, and
 *     enqueues everything on the caller's stream (a cudaStream_t passed as void*).
 *   - Every function returns VQB_OK (0), a negative VQB_E_* argument error, or a
 *     positive cudaError_t.  vqb_error_string() names any of them.
 *   - Thread-safe for concurrent calls on different devices/streams.
 *   - fp32 only, like the reference.
 *
 * Numerics contract (DESIGN.md section "Exactness")
 *   idx[i] = argmin_k fl(fl(zz_i + ee_k) - fl(2 * dot_ik)) with every sum an
 *   ascending fmaf chain from +0 ("oracle order", oracle/vq_oracle.c), lowest
 *   index on ties, first NaN wins -- bit-identical to the oracle on every
 *   kernel path, including the tensor-core path (which only FILTERS candidates
 *   with tcgen05 and decides with the exact expression).
 *   zq[i] = fl(z_i + fl(E[idx_i] - z_i)) elementwise           (:111)
 *   loss  = m + beta*m, m = mean((E[idx]-z)^2)                  (:107-108)
 *   perplexity = exp(-sum_k p_k log(p_k + 1e-10)), p = counts/N (:114-115)
 */
#ifndef VQB200_H
#define VQB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB_VERSION 100 /* 0.1.0 */

/* ---- status codes ------------------------------------------------------- */
#define VQB_OK 0
#define VQB_E_ARG (-1)         /* null pointer / non-positive size / bad stride */
#define VQB_E_WORKSPACE (-2)   /* workspace too small (see vqb_workspace_bytes) */
#define VQB_E_UNSUPPORTED (-3) /* shape not supported by the requested kernel path */
#define VQB_E_DEVICE (-4)      /* device is not sm_100 */
#define VQB_E_DRIVER (-5)      /* driver entry point (tensor-map encode) unavailable */
#define VQB_E_HOSTCTX (-6)     /* host context misuse */

/* ---- flags for vqb_forward ---------------------------------------------- */
#define VQB_PATH_AUTO 0u   /* tcgen05 filter+refine where the shape allows, else FMA */
#define VQB_PATH_FMA 1u    /* force the CUDA-core FMA kernel (exact, any K/D/strides) */
#define VQB_PATH_TC 2u     /* force the tcgen05 kernel; VQB_E_UNSUPPORTED if not possible */
#define VQB_PATH_MASK 3u
/* The caller vouches that `workspace` still holds what an earlier vqb_forward call prepared from THIS codebook
 * (same contents, same k and d): the per-call codebook kernels are skipped.  VQB_KEEP_CODEBOOK covers the code norms
 * and the non-finite census (every path), VQB_KEEP_TC_IMAGE additionally the tcgen05 operand image (only meaningful
 * after a call that took the tcgen05 path with k <= 256; ignored for larger codebooks).  The PyTorch binding sets them
 * from (data_ptr, _version) of the weight tensor. */
/* id width of vqb_encode_host's idx_host (default: int64) */
#define VQB_IDS_I64 0u
#define VQB_IDS_U8 16u
#define VQB_IDS_U16 32u
#define VQB_IDS_MASK 48u
#define VQB_KEEP_CODEBOOK 4u
#define VQB_KEEP_TC_IMAGE 8u

typedef struct vqb_device_info {
    int device;
    int cc_major;
    int cc_minor;
    int sm_count;
    int max_smem_per_block; /* opt-in bytes */
    size_t l2_bytes;
    size_t total_mem;
} vqb_device_info;

int vqb_version(void);
const char *vqb_error_string(int code);
int vqb_query(int device, vqb_device_info *out);

/* Bytes of caller-owned device scratch every device entry point needs for a
 * codebook of (k, d).  Contents are private to the library and need not be
 * preserved between calls.  The buffer must be 256-byte aligned. */
size_t vqb_workspace_bytes(int k, int d);

/* Which kernel path VQB_PATH_AUTO would take for this shape on `device`
 * (returns VQB_PATH_FMA or VQB_PATH_TC, or a negative error). */
int vqb_select_path(int device, int64_t n, int k, int d, int64_t stride_row, int64_t stride_d);

/*
 * Forward: quantise N = n_outer * n_inner vectors of dimension d.
 *
 * z addressing (element strides): vector (b, t), component j lives at
 *     z[b * stride_outer + t * stride_inner + j * stride_d].
 * Contiguous (N, d):            n_outer = N, n_inner = 1, stride_outer = d, stride_d = 1.
 * The encoder's permuted view   (model/vq_vae_patch_embedd.py:91; physical (B, d, T)):
 *                               n_outer = B, n_inner = T, stride_outer = d*T,
 *                               stride_inner = 1, stride_d = T  -- read in place, no copy.
 *
 * Outputs (device; any of zq/idx/loss/perplexity/counts may be NULL):
 *   zq          (N, d) contiguous fp32 -- straight-through value z + (E[idx] - z)
 *   idx         (N)    int64           -- min_encoding_indices
 *   loss        (1)    fp32
 *   perplexity  (1)    fp32
 *   counts      (k)    uint64          -- code-usage histogram of this call
 *   stats       (4)    uint64 or NULL  -- [0] rows decided by the tcgen05 filter alone,
 *                                         [1] rows re-evaluated exactly, [2] rows on the
 *                                         non-finite path, [3] SM clocks (clock64) of the
 *                                         longest-running CTA of the tcgen05 kernel
 */
int vqb_forward(int device, const float *z, int64_t n_outer, int64_t n_inner, int d,
                int64_t stride_outer, int64_t stride_inner, int64_t stride_d,
                const float *codebook, int k, float beta,
                float *zq, int64_t *idx, float *loss, float *perplexity,
                unsigned long long *counts, unsigned long long *stats,
                void *workspace, size_t workspace_bytes, unsigned flags, void *stream);

/*
 * Backward of vqb_forward's (loss, zq) with respect to z and the codebook:
 *   grad_z[i]        = g_zq[i] + g_loss * 2 * (z_i - E[idx_i]) / (N*d)
 *   grad_codebook[c] = g_loss * beta * 2 / (N*d) * sum_{i: idx_i = c} (E[c] - z_i)
 * g_zq: (N, d) contiguous or NULL (zeros); g_loss: device scalar or NULL (zero);
 * grad_z: (N, d) contiguous or NULL; grad_codebook: (k, d), OVERWRITTEN, or NULL.
 * z uses the same strided addressing as vqb_forward.
 */
int vqb_backward(int device, const float *g_zq, const float *g_loss,
                 const float *z, int64_t n_outer, int64_t n_inner, int d,
                 int64_t stride_outer, int64_t stride_inner, int64_t stride_d,
                 const int64_t *idx, const float *codebook, int k, float beta,
                 float *grad_z, float *grad_codebook,
                 void *workspace, size_t workspace_bytes, void *stream);

/*
 * One per-token linear layer of the patch encoder's residual blocks with the element-wise work around it fused
 * (model/vq_vae_patch_embedd.py:60-74 ResBlock, :103-111 CNNBlock: Conv1d(k=3, pad=1) on a length-1 slice is a
 * dense layer with the centre tap W[:, :, 1]).  bf16 operands, fp32 accumulation (tcgen05), erf-form GELU to
 * bf16 accuracy (|error| <= 2.5e-5 + 2.5e-4 |x| before the bf16 rounding):
 *   mode 0:  out = bf16(gelu(a w^T + bias))                                   first conv of a block
 *   mode 1:  h += a w^T + bias (fp32, in place);  out = bf16(gelu(h)) or NULL second conv + residual + next GELU
 *   mode 2:  h  = a w^T + bias (fp32, written);   out = bf16(gelu(h)) or NULL patch embedding (:13-17, k zero-padded
 *            to 64) + the first block's leading GELU
 * a: (n_tokens, k) bf16 row-major, w: (n, k) bf16 row-major, bias: (n) fp32, h: (n_tokens, n) fp32,
 * out: (n_tokens, n) bf16.  k a multiple of 64, n a multiple of 256; all pointers 16-byte aligned.
 */
int vqb_token_linear(int device, const void *a_bf16, const void *w_bf16, const float *bias, float *h, void *out_bf16,
                     int64_t n_tokens, int k, int n, unsigned mode, void *stream);

/*
 * The same layer with THREE taps: the decoder's Conv1d(k = 3, pad = 1) along the positions of a cycle
 * (model/vq_vae_patch_embedd.py:60-74 ResBlock inside CNNBlock(seperate=False), :142-147), on row-major tokens:
 *     x[t] = sum_{tap = 0..2} w[:, tap * k_in : (tap + 1) * k_in] a[t + tap - 1]  + bias,   a[.] = 0 outside the token's cycle
 * (cycles are runs of tokens_per_cycle consecutive tokens; n_tokens a multiple of it, 128 a multiple of it), then the
 * epilogue of `mode` as in vqb_token_linear.  out_gelu = 0 stores out = bf16(x) (mode 0) / bf16(h) (modes 1, 2) instead
 * of its GELU -- the last block's output feeds PatchEmbeddingInverse (:19-57), which starts with a (transposed)
 * convolution, not a GELU.  w: (n, 3 * k_in) bf16 row-major, the three taps side by side (Conv1d weight (n, k_in, 3)
 * permuted to (n, 3, k_in)); taps = 1 is vqb_token_linear with the out_gelu switch.  A transposed convolution with
 * kernel = stride = s (PatchEmbeddingInverse.proj[0]) is the one-tap layer with n = s * channels.
 */
int vqb_token_conv(int device, const void *a_bf16, const void *w_bf16, const float *bias, float *h, void *out_bf16,
                   int64_t n_tokens, int k_in, int n, unsigned mode, int taps, int tokens_per_cycle, int out_gelu,
                   void *stream);

/*
 * vqb_token_linear in its fp32-faithful form (the same reference layers, model/vq_vae_patch_embedd.py:60-74, :103-111, and
 * -- as a mode-2 call on zero-padded rows -- the final projection SepCNNBlock, :83-91): every operand is a PAIR of bf16
 * values, hi = bf16(x), lo = bf16(x - hi) (x to 2^-17), stored side by side:
 *   a_pair (n_tokens, 2 k) = [a_hi | a_lo],  w_pair (n, 2 k) = [w_hi | w_lo],  out_pair (n_tokens, 2 n) = [g_hi | g_lo].
 * The tensor cores accumulate a_hi w_hi + a_hi w_lo + a_lo w_hi in fp32 (operand precision 2^-16 instead of 2^-8, three
 * times the tensor work); the GELU is the erf form torch evaluates (erff); modes as in vqb_token_linear, out_gelu as in
 * vqb_token_conv (0: the pair of x / h itself).  bias (n) fp32, h (n_tokens, n) fp32.  Same shape / alignment rules.
 */
int vqb_token_linear_split(int device, const void *a_pair, const void *w_pair, const float *bias, float *h, void *out_pair,
                           int64_t n_tokens, int k, int n, unsigned mode, int out_gelu, void *stream);

/* vqb_token_conv in the fp32-faithful form (the decoder's Conv1d(k = 3, pad = 1) layers, its 1x1 input convolution and the first
 * transposed convolution of PatchEmbeddingInverse; model/vq_vae_patch_embedd.py:19-57,60-74,142-147): operands as bf16 pairs like
 * vqb_token_linear_split -- a_pair (n_tokens, 2 k_in) = [a_hi | a_lo], w_pair (n, taps * 2 k_in) = per tap [w_hi | w_lo],
 * out_pair (n_tokens, 2 n).  Cycles, taps, modes and out_gelu as in vqb_token_conv. */
int vqb_token_conv_split(int device, const void *a_pair, const void *w_pair, const float *bias, float *h, void *out_pair,
                         int64_t n_tokens, int k_in, int n, unsigned mode, int taps, int tokens_per_cycle, int out_gelu,
                         void *stream);

/* vqb_token_out_proj on a bf16 pair: a_pair (n_tokens, 2 * group * hidden) = [hi | lo], each half `group` runs of `hidden`
 * channels (the first transposed convolution's output: group = its kernel size); out (n_tokens * group, p) fp32 =
 * (hi + lo) w^T + bias, the hi half first, the lo half added in a second pass. */
int vqb_token_out_proj_pair(int device, const void *a_pair, const float *w, float bias, float *out, int64_t n_tokens, int group,
                            int hidden, int p, void *stream);

/* out_pair (n_tokens, 2 n) = [bf16(g) | bf16(g - bf16(g))] with g = gelu(h) (erf form; apply_gelu != 0) or g = h:
 * the operand pair of the first vqb_token_linear_split call, from the patch embedding's fp32 rows.  n a multiple of 4. */
int vqb_token_pair(int device, const float *h, void *out_pair, int64_t n_tokens, int n, int apply_gelu, void *stream);

/*
 * PatchEmbeddingInverse.proj[3], ConvTranspose1d(hidden, 1, kernel = stride = p) (model/vq_vae_patch_embedd.py:24-29):
 * out[r][j] = sum_c a[r][c] w[j][c] + bias for every row r of a (n_rows, hidden) bf16 activation; w (p, hidden) fp32
 * (the weight (hidden, 1, p) transposed), out (n_rows, p) fp32.  hidden 256 or 512, p <= 8.
 */
int vqb_token_out_proj(int device, const void *a_bf16, const float *w, float bias, float *out, int64_t n_rows, int hidden,
                       int p, void *stream);

/*
 * The whole residual-block chain of the patch encoder in one launch (model/vq_vae_patch_embedd.py:60-74 ResBlock,
 * :103-111 CNNBlock with seperate=True; n_layers = 2 * n_resblocks dense layers, the centre taps of the k=3 convs):
 *     for every block b:  h <- h + W[2b+1] gelu(W[2b] gelu(h) + bias[2b]) + bias[2b+1]
 * A CTA keeps a 128-token tile on its SM for all layers (activations alternate between shared and tensor memory, the
 * weights stream from L2), so HBM sees 1 read of a0 and h and 1 write of h per token.  Same arithmetic per layer as
 * vqb_token_linear (bf16 operands, fp32 accumulation, fp32 residual stream).
 *   a0_bf16 (n_tokens, hidden) bf16 = bf16(gelu(h)) of the incoming h (vqb_patch_embed writes both)
 *   h       (n_tokens, hidden) fp32, updated in place
 *   w_bf16  (n_layers, hidden, hidden) bf16 (out x in per layer), bias (n_layers, hidden) fp32
 *   scratch vqb_encoder_chain_scratch_bytes(device, hidden) bytes of device memory (contents private)
 * proj_dim > 0 fuses the encoder's final per-token projection hidden -> proj_dim (SepCNNBlock, :83-91) as one more
 * narrow GEMM on the resident tile, both operands as exact bf16 hi + lo pairs (fp32-accurate to 2^-16): w_bf16 then
 * holds 128 further rows behind the layers -- rows [0, proj_dim) = bf16(Wp), rows [64, 64 + proj_dim) =
 * bf16(Wp - bf16(Wp)), zeros elsewhere -- and z_e (n_tokens, proj_dim) fp32 = h Wp^T + proj_bias is written; h is
 * then NOT written back.  proj_dim a multiple of 4, <= 64.
 * pre_bias != NULL fuses the patch embedding (PatchEmbedding, :7-17) as a first short GEMM: a0_bf16 is then the
 * (n_tokens, 64) operand vqb_patch_split writes, w_bf16 holds `hidden` further rows at its end (columns [0, 32) =
 * bf16(Wpe), [32, 64) = bf16(Wpe - bf16(Wpe)), zero beyond the patch size; Wpe = Conv1d weight[:, 0, :]), pre_bias
 * (hidden) is the Conv1d bias, and h is not read.  With both fusions h may be NULL: raw samples in, z_e out.
 * hidden in {256, 512}, n_layers even; all pointers 16-byte aligned.
 */
size_t vqb_encoder_chain_scratch_bytes(int device, int hidden);
int vqb_encoder_chain(int device, const void *a0_bf16, float *h, const void *w_bf16, const float *bias, int64_t n_tokens,
                      int hidden, int n_layers, void *scratch, size_t scratch_bytes, const float *proj_bias, float *z_e,
                      int proj_dim, const float *pre_bias, void *stream);

/* Operand of the fused patch embedding: x (n_cycles, seq_len, channels) fp32 contiguous -> out (n_tokens, 64) bf16,
 * token order as in vqb_patch_embed (channel-major), row = [bf16(x_k), k < patch, zeros to 32 | bf16(x_k - bf16(x_k)),
 * zeros to 32].  patch <= 32, seq_len a multiple of patch. */
int vqb_patch_split(int device, const float *x, int64_t n_cycles, int seq_len, int channels, int patch, void *out_bf16,
                    void *stream);

/* h += bias (fp32 (n_tokens, n), in place); out = bf16(gelu(h)): the element-wise step between the fp32 patch
 * embedding (model/vq_vae_patch_embedd.py:13-17) and the first fused layer, in one pass.  n a multiple of 4. */
int vqb_token_bias_gelu(int device, float *h, const float *bias, void *out_bf16, int64_t n_tokens, int n, void *stream);

/* Patch embedding fused with the first block's leading GELU (model/vq_vae_patch_embedd.py:7-17, the
 * PatchEmbedding.forward the reference calls at :158 and dataloader/latentspace_dataloader.py:147,157):
 * x (n_cycles, seq_len, channels) fp32 contiguous -> tokens in the reference's channel-major order (all patches of
 * channel 0 first), h (n_cycles * T, hidden) fp32 = patches @ w^T + bias with T = channels * seq_len / patch,
 * out = bf16(gelu(h)) (may be NULL).  w: (hidden, patch) fp32 row-major = Conv1d weight[:, 0, :].  fp32 FMA arithmetic.
 * hidden == 512, patch <= 64, seq_len a multiple of patch; h 16-byte aligned. */
int vqb_patch_embed(int device, const float *x, int64_t n_cycles, int seq_len, int channels, int patch, const float *w,
                    const float *bias, float *h, void *out_bf16, int hidden, void *stream);

/* Packing copy for the reference encoder's output layout (model/vq_vae_patch_embedd.py:91 returns z_e as a permuted
 * view, logical (B, T, D) over physical (B, D, T); model/vector_quantizer.py:88 reshapes it into rows): vector (b, t),
 * component j at z[b * stride_outer + t * stride_inner + j * stride_d] -> out (n_outer * n_inner, d) contiguous.
 * Supported: stride_inner == 1, stride_d == n_inner (<= 64), d <= 128; anything else returns VQB_E_UNSUPPORTED and the
 * caller makes the copy itself. */
int vqb_pack_rows(int device, const float *z, int64_t n_outer, int64_t n_inner, int d, int64_t stride_outer,
                  int64_t stride_inner, int64_t stride_d, float *out, void *stream);

/* Input / target rows of the transformer's next-token task from token ids (dataloader/base_dataloader.py:74-110,
 * MyLatentAutoregressiveDataset): ids (n_windows, n_tokens) int64 -> x = [start_token, ids...], y = [ids..., end_token],
 * both (n_windows, n_tokens + 1) int64.  The reference takes start / end = max id of the data set + 1 / + 2. */
int vqb_ar_pairs(int device, const int64_t *ids, int64_t n_windows, int n_tokens, int64_t start_token, int64_t end_token,
                 int64_t *x, int64_t *y, void *stream);

/* De-duplicating data-set builder (the reference encodes every cycle of every overlapping
 * window, dataloader/latentspace_dataloader.py:225-238 over dataloader/asimow_dataloader.py:185-206).
 * vqb_row_keys: two 64-bit fingerprints per row of `words` 32-bit words: keys[2 i + s] = sum_j uint32(rows[i][j]) *
 * mult[s * words + j] (mod 2^64; mult = 2 x words multipliers).  One pass over the rows.
 * vqb_dedupe_first: first[i] = the smallest row index whose two keys equal row i's (i itself if it is the first; also i
 * itself in the 2^-64 case of a row that shares key 0, but not key 1, with an earlier row) -- deterministic.  scratch: vqb_dedupe_scratch_bytes(n)
 * of device memory, contents irrelevant.  The caller verifies rows against `first` word for word. */
int vqb_row_keys(int device, const void *rows, int64_t n_rows, int words, const unsigned long long *mult,
                 unsigned long long *keys, void *stream);
size_t vqb_dedupe_scratch_bytes(int64_t n_rows);
int vqb_dedupe_first(int device, const unsigned long long *keys, int64_t n_rows, void *scratch, size_t scratch_bytes,
                     int64_t *first, void *stream);

/* out[i] = codebook[idx[i]] (n, d).  Out-of-range indices yield NaN rows and set
 * *bad_index (device int, may be NULL) to 1. */
int vqb_gather(int device, const int64_t *idx, int64_t n, const float *codebook, int k, int d,
               float *out, int *bad_index, void *stream);

/* onehot (n, k) fp32 <- one_hot(idx).  4*k bytes per vector: not part of the
 * hot path's roofline (no caller of the reference reads it). */
int vqb_one_hot(int device, const int64_t *idx, int64_t n, int k, float *onehot, void *stream);

/* ---- measurement hooks (bench.py) ---------------------------------------- */
/* Cumulative number of kernels this library has launched in this process. */
long long vqb_launch_counter(void);
/* While enabled, vqb_forward brackets its dominant kernel (the fused distance/argmin/gather
 * kernel) with CUDA events on the caller's stream; vqb_profile_collect waits for them and
 * returns the summed duration and the number of bracketed launches, then forgets them.
 * Brackets are kept per device: collect returns (and forgets) those of launches whose `device`
 * argument equals the calling thread's current device, so a process that drives several GPUs
 * does not mix them. */
int vqb_profile_enable(int on);
int vqb_profile_collect(double *ms_sum, int *launches);

/* Debug only: while a device buffer of vqb_debug_tc_trace_words() uint64 is set, the tcgen05
 * kernel records clock64() of eight pipeline events per tile there (tools/tc_trace.py). */
int vqb_debug_set_tc_trace(unsigned long long *buf);
size_t vqb_debug_tc_trace_words(void);
/* Debug / A-B only: which filter the tcgen05 forward kernel (K <= 256, D <= 32) uses -- -1 automatic (TF32 single product for
 * 16 < D <= 32, three bf16 products otherwise; the VQB_TF32 environment variable overrides), 0 bf16, 1 TF32.  The decision
 * is the oracle's with either; a cached operand image (VQB_KEEP_TC_IMAGE) belongs to the filter it was built for. */
int vqb_debug_set_filter(int mode);

/* ---- host-buffer path (copies inside the call) --------------------------- */
typedef struct vqb_host_ctx vqb_host_ctx;

/* Creates device staging for chunks of `chunk_rows` vectors of dimension d against
 * a k-entry codebook: `depth` (>=2) in-flight chunks, each with its own stream.
 * This is the only place the library owns device memory. */
int vqb_host_create(int device, int64_t chunk_rows, int d, int k, int depth, vqb_host_ctx **out);
int vqb_host_destroy(vqb_host_ctx *ctx);
/* Upload the codebook (host pointer, (k, d) fp32). */
int vqb_host_set_codebook(vqb_host_ctx *ctx, const float *codebook_host);
/*
 * Quantise n host vectors (contiguous (n, d) fp32; pinned memory gives async copies).
 * H2D copy, kernels and D2H copy of successive chunks overlap.  zq_host / idx_host may
 * be NULL (ids-only is what dataloader/latentspace_dataloader.py:160-161 consumes).
 * idx_host receives int64 ids (what the reference's loop stores, :220) unless flags carry VQB_IDS_U8 (k <= 256) or
 * VQB_IDS_U16 (k <= 65536): the same ids narrowed on the device, 8x / 4x fewer bytes over PCIe.
 * flags = VQB_PATH_* | VQB_IDS_*.  The codebook kernels run once per vqb_host_set_codebook, not once per call.
 * Scalars are written to host on return (the call synchronises its own streams; on an error return all of its
 * streams have been drained, so no copy still targets the caller's buffers).
 * launches_out (optional) receives the number of kernels launched.
 */
int vqb_encode_host(vqb_host_ctx *ctx, const float *z_host, int64_t n, float beta,
                    float *zq_host, void *idx_host, float *loss_host, float *perplexity_host,
                    unsigned long long *counts_host, unsigned flags, int *launches_out);
/* Device-timed duration (CUDA events, first H2D to last D2H) of the last vqb_encode_host. */
int vqb_host_last_ms(vqb_host_ctx *ctx, float *ms);

#ifdef __cplusplus
}
#endif
#endif /* VQB200_H */
