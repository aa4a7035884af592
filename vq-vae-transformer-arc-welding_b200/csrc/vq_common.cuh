// vq_common.cuh -- shared declarations of the B200 vector-quantisation kernels.
//
// Reference path being replaced: tmdt-buw/VQ-VAE-Transformer-Arc-Welding,
// model/vector_quantizer.py:76-131.  All kernels are written for sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vqb200.h"

namespace vqb {

// ---------------------------------------------------------------------------------------
// Workspace layout (caller-owned device scratch, see vqb_workspace_bytes()).
// ---------------------------------------------------------------------------------------
constexpr int kMaxPartials = 4096;       // loss partial sums, one per CTA of the forward kernel
constexpr int kHeaderBytes = 256;

struct WsHeader {                        // first kHeaderBytes of the workspace
    unsigned long long stats[4];         // filter-only rows, refined rows, non-finite rows, reserved
    int poisoned_columns;                // number of codebook columns holding a non-finite entry
    int bad_index;
    int pad[2];
};

struct WsLayout {
    size_t off_ee;        // float [kpad]    ||E_k||^2 in oracle order, pads = +inf
    size_t off_colcnt;    // int   [d]       non-finite entries per codebook column
    size_t off_colwhich;  // int   [d]       (largest) code index + 1 with a non-finite entry there
    size_t off_counts;    // u64   [k]       histogram when the caller passes counts == NULL
    size_t off_partials;  // double[kMaxPartials]
    size_t off_sq;        // double[1]       total of squares (host path keeps it across chunks)
    size_t off_tc;        // float [...]     tensor-core path: prepared codebook operand + bounds
    size_t total;
    int kpad;
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Extra floats the tcgen05 path keeps per codebook: B operand tile(s) + per-code data.
__host__ inline size_t tc_scratch_floats(int k, int d)
{
    // tcgen05 path (vq_fwd_tc.cu, tc::IMG_BYTES): image of the constant operands for 256 codes,
    // the scalar constants, and the per-CTA queues of vectors deferred to the fix-up kernel
    const size_t base = 32768 + 8192 + 32768 + 1024 + 64 + 192 * 4 + 192 * 2048 * 8;
    // tile-stationary path (vq_fwd_tcs.cu, tcs::img_bytes): one 40 KB operand block per (256-code chunk, 32-component
    // D-chunk), the scalar constants and queue counters (1 KB) and the per-CTA queues (16 bytes per entry)
    const size_t nc = (size_t)(k + 255) / 256, nd = (size_t)(d + 31) / 32;
    const size_t tcs = (k > 256 || d > 32) ? nc * nd * 40960 + 1024 + 192 * 2048 * 16 : 0;
    return (base > tcs ? base : tcs) / sizeof(float) + 64;
}

__host__ inline WsLayout ws_layout(int k, int d)
{
    WsLayout L;
    L.kpad = (int)align_up((size_t)k, 16);
    size_t o = kHeaderBytes;
    L.off_ee = o;        o = align_up(o + sizeof(float) * L.kpad, 256);
    L.off_colcnt = o;    o = align_up(o + sizeof(int) * d, 256);
    L.off_colwhich = o;  o = align_up(o + sizeof(int) * d, 256);
    L.off_counts = o;    o = align_up(o + sizeof(unsigned long long) * k, 256);
    L.off_partials = o;  o = align_up(o + sizeof(double) * kMaxPartials, 256);
    L.off_sq = o;        o = align_up(o + sizeof(double) * 4, 256);
    L.off_tc = o;        o = align_up(o + sizeof(float) * tc_scratch_floats(k, d), 256);
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------
// Strided view of the input vectors: vector (b, t), component j at
// base[b * s_outer + t * s_inner + j * s_d]  (element strides).
// ---------------------------------------------------------------------------------------
struct ZView {
    const float *base;
    int64_t n_rows;     // n_outer * n_inner
    int64_t n_inner;
    int64_t s_outer, s_inner, s_d;
    __device__ __forceinline__ const float *row(int64_t r) const
    {
        const int64_t b = r / n_inner, t = r - b * n_inner;
        return base + b * s_outer + t * s_inner;
    }
    __host__ __device__ bool rows_contiguous(int d) const
    {   // plain (N, d) row-major
        return s_d == 1 && ((n_inner == 1 && s_outer == d) ||
                            (s_inner == d && (s_outer == n_inner * (int64_t)d || n_rows == n_inner)));
    }
};

struct FwdParams {
    ZView z;
    const float *E;       // (K, D) row-major
    int K, D;
    const float *ee;      // [kpad] from the prep kernel
    const int *colcnt;    // [D]
    const int *colwhich;  // [D]
    const WsHeader *hdr_in;
    float *zq;            // (N, D) contiguous or nullptr
    int64_t *idx;         // (N) or nullptr
    unsigned long long *counts;  // (K), accumulated atomically
    double *partials;     // [gridDim.x] sum of squared residuals per CTA
    int accumulate;       // partials[b] += instead of =
    int kt;               // codes per shared-memory tile (multiple of 4)
    int need_sq;          // the caller wants the loss: residual sums are needed even when zq == nullptr
    unsigned long long *stats;
    // tcgen05 path, codebooks of more than 256 codes (one pass per 256-code chunk, vq_fwd_tc.cu):
    unsigned long long *run;     // running best per vector: float bits of the exact distance << 32 | code
    int code_base;               // first code of this pass's chunk
    int chunk_mode;              // 0: single pass with all outputs; 1: first chunk; 2: later chunk
};

// launchers implemented in the .cu files -------------------------------------------------
cudaError_t launch_prep(const float *E, int K, int D, int kpad, float *ee, int *colcnt, int *colwhich,
                        WsHeader *hdr, cudaStream_t st);
cudaError_t launch_fwd_fma(const FwdParams &p, int sm_count, int max_smem, int *n_ctas, cudaStream_t st);
cudaError_t launch_finalize(const unsigned long long *counts, int K, const double *partials, int n_partials,
                            double *sq_total_io, int sq_mode, int64_t n_rows, int D, float beta,
                            float *loss, float *perplexity, cudaStream_t st);
cudaError_t launch_gather(const int64_t *idx, int64_t n, const float *E, int K, int D, float *out,
                          int *bad_index, cudaStream_t st);
cudaError_t launch_one_hot(const int64_t *idx, int64_t n, int K, float *onehot, cudaStream_t st);
cudaError_t launch_bwd(const float *g_zq, const float *g_loss, const ZView &z, const int64_t *idx,
                       const float *E, int K, int D, float beta, float *grad_z, float *grad_E,
                       int sm_count, int max_smem, cudaStream_t st);

// tcgen05 path (vq_fwd_tc.cu)
bool tc_shape_supported(int K, int D);
cudaError_t launch_fwd_tc(const FwdParams &p, float *tc_scratch, int sm_count, int max_smem, int *n_ctas,
                          int *n_launches, cudaStream_t st, cudaEvent_t ev_begin = nullptr,
                          cudaEvent_t ev_end = nullptr, bool image_ready = false);
bool tc_chunked_supported(int K, int D);
cudaError_t launch_fwd_tc_chunked(const FwdParams &p, float *tc_scratch, int sm_count, int max_smem, int *n_ctas,
                                  int *n_launches, cudaStream_t st, cudaEvent_t ev_begin = nullptr,
                                  cudaEvent_t ev_end = nullptr);
bool tok_linear_supported(int K, int N);
cudaError_t launch_tok_linear(const void *a, const void *w, const float *bias, float *h, void *out, int64_t n_tokens,
                              int K, int N, int mode, int sm_count, int max_smem, cudaStream_t st, int taps = 1,
                              int cyc_len = 1, int out_gelu = 1, int split = 0);
cudaError_t launch_tok_pair(const float *h, void *out, int64_t n_tokens, int N, int apply_gelu, int sm_count, cudaStream_t st);
bool tok_out_proj_supported(int H, int P);
cudaError_t launch_tok_out_proj(const void *a, const float *w, float bias, float *out, int64_t n_rows, int H, int P, int sm_count,
                                cudaStream_t st, int group = 1, int64_t ld_group = 0, int accumulate = 0);
cudaError_t launch_tok_bias_gelu(float *h, const float *bias, void *out, int64_t n_tokens, int N, int sm_count,
                                 cudaStream_t st);
bool enc_chain_supported(int H, int L);
size_t enc_chain_scratch_bytes(int H, int sm_count);
cudaError_t launch_enc_chain(const void *a0, float *h, const void *w, const float *bias, int64_t n_tokens, int H, int L,
                             float *scratch, size_t scratch_bytes, const float *proj_bias, float *z_e, int proj_d,
                             const float *pre_bias, int sm_count, int max_smem, cudaStream_t st);
bool patch_split_supported(int L, int C, int P);
cudaError_t launch_patch_split(const float *x, void *out, int64_t n_cycles, int L, int C, int P, int sm_count, cudaStream_t st);
bool patch_embed_supported(int L, int C, int P, int H);
cudaError_t launch_patch_embed(const float *x, const float *w, const float *bias, float *h, void *a, int64_t n_cycles,
                               int L, int C, int P, int H, int sm_count, int max_smem, cudaStream_t st);
bool pack_rows_supported(int64_t n_inner, int d, int64_t s_outer, int64_t s_inner, int64_t s_d);
cudaError_t launch_pack_rows(const float *src, float *dst, int64_t n_outer, int64_t n_inner, int d, int64_t s_outer,
                             int64_t s_inner, int64_t s_d, int sm_count, cudaStream_t st);
// vq_dedupe.cu
cudaError_t launch_row_keys(const void *rows, int64_t n, int words, const unsigned long long *mult, unsigned long long *keys,
                            int sm_count, cudaStream_t st);
size_t dedupe_scratch_bytes(int64_t n);
cudaError_t launch_dedupe_first(const unsigned long long *keys, int64_t n, void *scratch, int64_t *first, int sm_count,
                                cudaStream_t st);
cudaError_t launch_ar_pairs(const int64_t *ids, int64_t n_windows, int n, int64_t start_token, int64_t end_token, int64_t *x,
                            int64_t *y, int sm_count, cudaStream_t st);
void count_launches(int n);
void set_tc_trace(unsigned long long *buf);
void set_tc_filter(int mode);
size_t tc_trace_words();

// ---------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------
// NaN-propagating minimum: a NaN distance must survive so that the row can be
// re-scanned with torch.argmin's "first NaN wins" rule.
__device__ __forceinline__ float min_nan(float a, float b)
{
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// The reference's distance expression with its own association and rounding:
// fl(fl(zz + ee) - fl(2*dot)); never contracted into an FMA.
__device__ __forceinline__ float ref_distance(float zz, float ee, float dot)
{
    return __fsub_rn(__fadd_rn(zz, ee), __fmul_rn(2.0f, dot));
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Adds one to counts[code] for every active lane with code >= 0, one atomic per
// distinct code in the warp.
__device__ __forceinline__ void warp_histogram_add(unsigned long long *counts, int code)
{
    const unsigned peers = __match_any_sync(0xffffffffu, code);
    const int leader = __ffs(peers) - 1;
    if (code >= 0 && (int)(threadIdx.x & 31) == leader)
        atomicAdd(counts + code, (unsigned long long)__popc(peers));
}

}  // namespace vqb
