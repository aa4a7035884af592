// enc_chain.cu -- the whole residual-block chain of the patch encoder in ONE kernel (sm_100a, tcgen05 + TMA + TMEM).
//
// Reference: model/vq_vae_patch_embedd.py:60-74 (ResBlock), :103-111 (CNNBlock with seperate=True): to every token
// independently, n_resblocks times   h <- h + W2 gelu(W1 gelu(h) + b1) + b2   with H x H weights (the centre tap of
// Conv1d(k=3, pad=1) on a length-1 slice).  The layer-at-a-time kernel (tok_linear.cu) round-trips 6-8 KB per token
// and block through HBM; here a CTA keeps a 128-token tile on the SM for all 2 * n_resblocks GEMMs:
//
//   * the bf16 activation tile (128 x H) ALTERNATES between shared memory and tensor memory: GEMM g reads its A
//     operand from one of the two (tcgen05.mma with an smem descriptor, or with A in TMEM) while its epilogue writes
//     the next GEMM's A operand into the other (st.shared in the UMMA K-major SW128 layout, or tcgen05.st of packed
//     bf16 pairs).  One 128 KB activation buffer in shared memory therefore suffices, which leaves room for a
//     6-stage weight ring (two smem buffers would not fit beside any ring at H = 512);
//   * accumulators are N = 128 column quarters, two of them in flight (TMEM columns: H/2 for the activation operand,
//     2 x 128 for the accumulators): the epilogue of quarter q overlaps the MMAs of quarter q + 1, and the next GEMM
//     starts on the K-chunks the epilogue has already produced;
//   * weights stream from L2 through a TMA ring that runs ahead across GEMM boundaries (they do not depend on the
//     activations): 128 x 64 bf16 boxes of the stacked (L * H, H) weight tensor;
//   * the fp32 residual stream lives in a per-CTA scratch tile (L2-resident: 148 x 256 KB) laid out so that the
//     epilogue's thread-per-row accesses coalesce; the first block reads it from, the last block writes it to the
//     caller's row-major h.
//
//   * optionally the encoder's final per-token projection H -> D (SepCNNBlock, model/vq_vae_patch_embedd.py:83-91) runs
//     as one more, narrow GEMM on the same tile.  Both of its operands are exact bf16 hi + lo pairs: the weights in two
//     column groups of one N = 128 accumulator, the residual stream h as two K passes -- bf16(h) written by the last
//     layer's epilogue, bf16(h - bf16(h)) written into the OTHER activation buffer once the last layer's MMAs are done
//     (rounding h to a single bf16 costs ids: 99.70 % instead of 99.86 % of them equal to the fp32 encoder's).
//     z_e = acc_hi + acc_lo + bias leaves as fp32 rows -- the residual stream then never goes back to HBM at all;
//   * optionally the patch embedding (PatchEmbedding, :7-17) runs as a first, short GEMM on the tile (K = patch <= 32,
//     samples and weights as bf16 hi + lo pairs, three products like the quantiser's filter: fp32-accurate to 2^-16),
//     fed by a 128-byte-per-token operand the gather kernel below writes -- instead of 3 KB per token of h and a.
//
// bf16 operands, fp32 accumulation, fp32 residual stream, GELU as in tok_linear.cu (gelu_fast): the same arithmetic
// as the layer-at-a-time path, layer for layer.
#include "vq_common.cuh"
#include "vq_ptx.cuh"

namespace vqb {

namespace ec {

constexpr int BM = 128;                    // tokens per tile (TMEM lanes)
constexpr int NQ = 128;                    // accumulator width (columns of one MMA)
constexpr int BK = 64;                     // K-chunk of one TMA box (128 bytes of bf16: one SW128 row)
constexpr int KS = 128;                    // K per weight stage = two boxes = one accumulator quarter of the previous GEMM
constexpr int W_STAGE = NQ * KS * 2;       // 32 KB: eight MMAs per barrier round trip of the issuing warp
constexpr int W_STAGES = 3;
constexpr int EPI_WARPS = 16;              // lane quarter x 32-column slab of the 128-column accumulator
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = 128 + EPI_THREADS; // 4 service warps + 16 epilogue warps (4 per SM sub-partition hide each other's latencies)
#ifndef EC_ABLATE
#define EC_ABLATE 0                        // timing experiments: 1 = epilogue without arithmetic / stores (results are wrong)
#endif

template <int H> struct Plan {
    static constexpr int A_BYTES = BM * H * 2;                 // activation tile, H/64 K-chunks of 16 KB
    static constexpr int OFF_A = 0;
    static constexpr int OFF_W = OFF_A + A_BYTES;
    static constexpr int OFF_BIAS = OFF_W + W_STAGES * W_STAGE;  // the current layer's bias (H floats)
    static constexpr int OFF_BARS = OFF_BIAS + H * 4;
    static constexpr int SMEM_BYTES = OFF_BARS + 256;
    static constexpr int TMEM_A_COLS = H / 2;                  // packed bf16 pairs
    static constexpr int TMEM_ACC0 = 256;                      // accumulators at columns 256 .. 511
};

__device__ __forceinline__ float gelu(float x)
{   // identical to tl::gelu_fast (tok_linear.cu): 0.5 x (1 + tanh(x (c1 + c3 x^2 + c5 x^4))), minimax fit of the erf form
    const float u = fminf(x * x, 36.0f);
    float p = fmaf(-3.51517534e-4f, u, 3.70056510e-2f);
    p = fmaf(p, u, 7.97507878e-1f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// The same GELU on two values at once with sm_100's packed fp32 instructions (FMUL2 / FFMA2: IEEE results per lane, so
// bit-identical to the scalar form): 6 packed + 2 FMNMX + 2 MUFU per pair instead of 14 + 2 scalar instructions -- the
// epilogue's issue slots, not the tensor pipe, bound the chain (profiles/README.md, round 2).
__device__ __forceinline__ float2 gelu2(float2 x)
{
    float2 u = __fmul2_rn(x, x);
    u.x = fminf(u.x, 36.0f);
    u.y = fminf(u.y, 36.0f);
    float2 p = __ffma2_rn(make_float2(-3.51517534e-4f, -3.51517534e-4f), u, make_float2(3.70056510e-2f, 3.70056510e-2f));
    p = __ffma2_rn(p, u, make_float2(7.97507878e-1f, 7.97507878e-1f));
    const float2 y = __fmul2_rn(x, p);
    float2 t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(y.x));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(y.y));
    const float2 hx = __fmul2_rn(x, make_float2(0.5f, 0.5f));
    return __ffma2_rn(hx, t, hx);
}

// tcgen05.mma with the A operand in tensor memory (lanes = rows, 32-bit columns = K pairs)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16_one(uint32_t (&a)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace ec

// GEMM sequence of a tile: [PRE] LAYER x L [PROJ].  GEMM j reads its A operand from shared memory if j is even and from
// tensor memory if j is odd; its epilogue writes the next operand into the other one.
//
// a0     : without PRE: (n_tokens, H) bf16 row-major = bf16(gelu(h0)), the first GEMM's operand (TMA);
//          with PRE:    (n_tokens, 64) bf16 = [bf16(x) | bf16(x - bf16(x))] of the token's patch, 32 + 32 columns
//                       (zero beyond the patch size) -- vq_patch_split_kernel
// h      : (n_tokens, H) fp32 row-major.  Without PRE it holds h0 on entry; without PROJ it receives the residual
//          stream after the last block; with both it is not touched (may be NULL)
// w      : (L * H [+ 128] [+ H], H) bf16 row-major: layer l = rows [l*H, (l+1)*H) (out x in), then PROJ's 128 rows
//          (rows [0, D) = bf16(Wp), rows [64, 64 + D) = bf16(Wp - bf16(Wp)), zeros elsewhere), then PRE's H rows
//          (columns [0, 32) = bf16(Wpe), [32, 64) = bf16(Wpe - bf16(Wpe)), zero beyond the patch size)
// bias   : (L * H) fp32;  proj_bias (D);  pre_bias (H)
// scratch: gridDim.x tiles of BM * H fp32 (the residual stream between the blocks)
template <int H>
__global__ void __launch_bounds__(ec::THREADS, 1)
enc_chain_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_w,
                 const float *__restrict__ bias, float *__restrict__ h, float *__restrict__ scratch,
                 int64_t n_tokens, int L, const float *__restrict__ proj_bias, float *__restrict__ z_e, int proj_d,
                 const float *__restrict__ pre_bias)
{
    using namespace tc;
    using namespace ec;
    using P = Plan<H>;
    constexpr int NQT = H / NQ;                // accumulator quarters per GEMM (4 at H = 512)
    constexpr int NKC = H / BK;                // 64-wide K-chunks of the activation tile (8 at H = 512)
    constexpr int NKS = H / KS;                // weight stages per accumulator quarter (4 at H = 512)
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    if ((sbase & 1023u) != 0)
        __trap();
    enum { W_FULL = 0, W_EMPTY = W_FULL + W_STAGES, ACC_FULL = W_EMPTY + W_STAGES, ACC_EMPTY = ACC_FULL + 2,
           A_RDY = ACC_EMPTY + 2, A0_FULL = A_RDY + 4, A_FREE = A0_FULL + 1, LO_RDY = A_FREE + 1, N_BARS = LO_RDY + 1 };
    static_assert(8 * N_BARS + 8 <= 256, "barrier area");
    static_assert(P::SMEM_BYTES <= 232448, "shared memory plan exceeds 227 KB");
    auto bar = [&](int i) { return sbase + P::OFF_BARS + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + P::OFF_BARS + 8 * N_BARS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const bool pre = pre_bias != nullptr, proj = proj_d > 0;
    const int j_layer0 = pre ? 1 : 0;          // first LAYER GEMM
    const int j_proj = j_layer0 + L;           // PROJ GEMM (if any)
    const int NG = j_proj + (proj ? 1 : 0);    // GEMMs per tile
    const int w_row_proj = L * H, w_row_pre = L * H + (proj ? NQ : 0);
    // the last GEMM that READS the smem tile (A_FREE follows it): PROJ reads both buffers; else the last even GEMM
    const int j_free = proj ? j_proj : ((NG - 1) & 1 ? NG - 2 : NG - 1);
    const int64_t n_tiles = (n_tokens + BM - 1) / BM;
    const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < W_STAGES; ++s) {
            mbar_init(bar(W_FULL + s), 1);
            mbar_init(bar(W_EMPTY + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(ACC_FULL + b), 1);
            mbar_init(bar(ACC_EMPTY + b), EPI_THREADS);
        }
        for (int q = 0; q < 4; ++q)
            mbar_init(bar(A_RDY + q), EPI_THREADS);
        mbar_init(bar(A0_FULL), 1);
        mbar_init(bar(A_FREE), 1);
        mbar_init(bar(LO_RDY), EPI_THREADS);
        fence_barrier_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= weight producer: runs ahead of the MMAs across GEMM and tile boundaries =================
        int s = 0;
        uint32_t ph = 0;
        // a stage = two 16 KB boxes: weights [row, row + 128) x [col, col + 128); for PRE the second box is the tile's
        // patch operand instead (tok_row >= 0), so that the first GEMM of the NEXT tile never waits for the smem tile
        auto stage = [&](int row, int col, int64_t tok_row) {
            mbar_wait<32>(bar(W_EMPTY + s), ph ^ 1u);
            if (elect_one()) {
                mbar_expect_tx(bar(W_FULL + s), W_STAGE);
                tma_load_2d(sbase + P::OFF_W + s * W_STAGE, &map_w, bar(W_FULL + s), col, row);
                if (tok_row < 0)
                    tma_load_2d(sbase + P::OFF_W + s * W_STAGE + NQ * BK * 2, &map_w, bar(W_FULL + s), col + BK, row);
                else
                    tma_load_2d(sbase + P::OFF_W + s * W_STAGE + NQ * BK * 2, &map_a0, bar(W_FULL + s), 0, (int)tok_row);
            }
            __syncwarp();
            if (++s == W_STAGES) {
                s = 0;
                ph ^= 1u;
            }
        };
        for (int t = 0; t < my_tiles; ++t)
            for (int j = 0; j < NG; ++j) {
                if (pre && j == 0) {
                    const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
                    for (int q = 0; q < NQT; ++q)                         // [Wpe_hi | Wpe_lo] rows of the quarter + the patches
                        stage(w_row_pre + q * NQ, 0, tile * BM);
                } else if (proj && j == j_proj) {
                    for (int pass = 0; pass < 2; ++pass)                  // the hi and the lo pass use the same rows
                        for (int ks = 0; ks < NKS; ++ks)
                            stage(w_row_proj, ks * KS, -1);
                } else {
                    for (int q = 0; q < NQT; ++q)
                        for (int ks = 0; ks < NKS; ++ks)
                            stage((j - j_layer0) * H + q * NQ, ks * KS, -1);
                }
            }
    } else if (warp == 3) {
        // ================= without PRE: the first operand of every tile, bf16(gelu(h0)) rows, by TMA =================
        for (int t = 0; t < my_tiles && !pre; ++t) {
            if (t > 0)
                mbar_wait<64>(bar(A_FREE), (uint32_t)((t - 1) & 1));    // the last GEMM that read the smem tile is done
            const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
            if (elect_one()) {
                mbar_expect_tx(bar(A0_FULL), P::A_BYTES);
#pragma unroll
                for (int kc = 0; kc < NKC; ++kc)
                    tma_load_2d(sbase + P::OFF_A + kc * (BM * BK * 2), &map_a0, bar(A0_FULL), kc * BK, (int)(tile * BM));
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp runs the loop (uniform control flow), one elected lane issues =====
        const uint32_t idesc = idesc_bf16(NQ);
        const uint64_t a_base = desc_sw128(sbase + P::OFF_A);    // + kc * 1024 in the address field (16 KB K-chunks)
        const uint64_t w_base = desc_sw128(sbase + P::OFF_W);    // + s * 2048 (32 KB stages), + 1024 for the second box
        int s = 0;
        uint32_t wph = 0;
        uint32_t acc_n = 0;                                      // accumulators issued so far
        uint32_t ardy_n = 0;                                     // completed A_RDY phases (per quarter barrier) so far
        auto next_stage = [&]() {
            if (++s == W_STAGES) {
                s = 0;
                wph ^= 1u;
            }
        };
        // one K pass (H wide) of an accumulator: operand from shared memory (K-chunks of the tile) or tensor memory
        auto k_pass = [&](uint32_t d, bool from_tmem, bool wait_a, bool fresh) {
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks) {
                if (wait_a)                                      // K range of stage ks = quarter ks of the previous epilogue
                    mbar_wait<32>(bar(A_RDY + ks), ardy_n & 1u);
                mbar_wait<32>(bar(W_FULL + s), wph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t wd = w_base + (uint64_t)(s * (W_STAGE >> 4));
#pragma unroll
                    for (int hk = 0; hk < KS / BK; ++hk) {       // the stage's two 64-wide boxes
                        const int kc = ks * (KS / BK) + hk;
                        if (!from_tmem) {
                            const uint64_t ad = a_base + (uint64_t)(kc * (BM * BK * 2 >> 4));
#pragma unroll
                            for (int i = 0; i < BK / 16; ++i)
                                umma_bf16(d, ad + 2 * i, wd + hk * (NQ * BK * 2 >> 4) + 2 * i, idesc, !fresh || (kc | i) != 0);
                        } else {
                            const uint32_t at = tmem_base + kc * (BK / 2);
#pragma unroll
                            for (int i = 0; i < BK / 16; ++i)
                                umma_bf16_ts(d, at + 8 * i, wd + hk * (NQ * BK * 2 >> 4) + 2 * i, idesc, !fresh || (kc | i) != 0);
                        }
                    }
                    umma_commit(bar(W_EMPTY + s));
                }
                __syncwarp();
                next_stage();
            }
        };
        for (int t = 0; t < my_tiles; ++t) {
            if (!pre)
                mbar_wait<32>(bar(A0_FULL), (uint32_t)(t & 1));
            for (int j = 0; j < NG; ++j) {
                const bool from_tmem = (j & 1) != 0;             // even GEMMs read shared memory, odd ones tensor memory
                const bool is_pre = pre && j == 0, is_proj = proj && j == j_proj;
                const int nq = is_proj ? 1 : NQT;
                for (int q = 0; q < nq; ++q, ++acc_n) {
                    const int ab = (int)(acc_n & 1u);
                    mbar_wait<32>(bar(ACC_EMPTY + ab), ((acc_n >> 1) & 1u) ^ 1u);
                    const uint32_t d = tmem_base + P::TMEM_ACC0 + ab * NQ;
                    if (is_pre) {
                        // h0 = x_hi Wpe_hi + x_hi Wpe_lo + x_lo Wpe_hi: K-slices 0, 1 = hi and 2, 3 = lo of the 64-column chunk
                        mbar_wait<32>(bar(W_FULL + s), wph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t wd = w_base + (uint64_t)(s * (W_STAGE >> 4));
                            const uint64_t pd = wd + (NQ * BK * 2 >> 4);    // the patch operand: second box of the stage
                            umma_bf16(d, pd + 0, wd + 0, idesc, 0);
                            umma_bf16(d, pd + 2, wd + 2, idesc, 1);
                            umma_bf16(d, pd + 0, wd + 4, idesc, 1);
                            umma_bf16(d, pd + 2, wd + 6, idesc, 1);
                            umma_bf16(d, pd + 4, wd + 0, idesc, 1);
                            umma_bf16(d, pd + 6, wd + 2, idesc, 1);
                            umma_commit(bar(W_EMPTY + s));
                        }
                        __syncwarp();
                        next_stage();
                    } else if (is_proj) {
                        k_pass(d, from_tmem, true, true);        // bf16(h) against [Wp_hi | Wp_lo]
                        mbar_wait<32>(bar(LO_RDY), (uint32_t)(t & 1));
                        k_pass(d, !from_tmem, false, false);     // bf16(h - bf16(h)) from the other buffer, same accumulator
                    } else {
                        k_pass(d, from_tmem, j > 0 && q == 0, true);
                    }
                    if (elect_one())
                        umma_commit(bar(ACC_FULL + ab));
                    __syncwarp();
                }
                if (j > 0)
                    ++ardy_n;
                if (j == j_free && !pre) {
                    if (elect_one())
                        umma_commit(bar(A_FREE));                // every MMA that reads the smem tile has been issued
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> bias (+ residual) -> GELU -> bf16 -> next operand =================
        // warp -> (TMEM lane quarter, 32-column slab of the accumulator): one slab per thread and accumulator quarter
        const int q4 = warp & 3;                                 // TMEM lane quarter
        const int sl = (warp - 4) >> 2;                          // slab 0..3 of the 128-column accumulator
        const int r = q4 * 32 + lane;                            // row of the tile
        const int et = tid - 128;                                // 0 .. EPI_THREADS-1
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        float *scr = scratch + (size_t)blockIdx.x * (BM * H);
        float *bias_s = reinterpret_cast<float *>(smem + P::OFF_BIAS);
        // 16 packed bf16 pairs of this thread's row, columns [col, col + 16), into the operand buffer GEMM j + 1 reads
        auto put_operand = [&](bool to_tmem, int col, const uint32_t (&pk)[8]) {
            if (to_tmem) {
                tmem_st8(tmem_base + (col >> 1) + lane_addr, pk);
            } else {
                unsigned char *arow = smem + P::OFF_A + (col >> 6) * (BM * BK * 2) + r * 128;
                const int c16 = (col & 63) >> 3;                 // first 16-byte chunk of the 128-byte row
#pragma unroll
                for (int c = 0; c < 2; ++c)
                    *reinterpret_cast<uint4 *>(arow + (((c16 + c) ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
        };
        auto publish = [&](bool to_tmem, int bar_id) {           // make the operand visible to the MMAs, then arrive
            if (to_tmem) {
                tmem_wait_st();
                tc_fence_before();
            } else {
                fence_proxy_async();
            }
            mbar_arrive(bar(bar_id));
        };
        uint32_t acc_n = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
            const int64_t row = tile * BM + r;
            const bool row_ok = row < n_tokens;
            for (int j = 0; j < NG; ++j) {
                const bool is_pre = pre && j == 0, is_proj = proj && j == j_proj;
                const int l = j - j_layer0;                      // layer index of a LAYER GEMM
                const bool resid = !is_pre && !is_proj && (l & 1) != 0;      // second GEMM of a block: + residual
                const bool last = j == NG - 1;
                const bool to_tmem = (j & 1) == 0;               // where the NEXT operand goes
                const bool plain = proj && j == j_proj - 1;      // PROJ's operand is bf16(h), not bf16(gelu(h))
                const bool h_in_std = !pre && l == 1;            // without PRE the caller's h holds h0
                const bool h_out_std = !proj && l == L - 1;      // without PROJ the caller's h receives the result
                // this GEMM's bias: fetched now, parked in shared memory behind the first quarter's barrier (every warp
                // has then finished the previous GEMM), published by a second barrier
                float bias_reg = 0.0f;
                if (is_proj) {
                    if (et < proj_d)
                        bias_reg = __ldg(proj_bias + et);
                } else if (et < H) {
                    bias_reg = __ldg((is_pre ? pre_bias : bias + (size_t)l * H) + et);
                }
                for (int q = 0; q < (is_proj ? 1 : NQT); ++q, ++acc_n) {
                    const int ab = (int)(acc_n & 1u);
                    const int col0 = q * NQ + sl * 32;           // this thread's 32 columns of the GEMM's output
                    // the residual slab is requested before the accumulator is waited for
                    float hv[32];
                    if (resid) {
                        if (h_in_std) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float4 v = row_ok ? __ldcs(reinterpret_cast<const float4 *>(h + row * H + col0) + c)
                                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                                hv[4 * c] = v.x; hv[4 * c + 1] = v.y; hv[4 * c + 2] = v.z; hv[4 * c + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float4 v = *reinterpret_cast<const float4 *>(scr + ((size_t)((col0 >> 2) + c) * BM + r) * 4);
                                hv[4 * c] = v.x; hv[4 * c + 1] = v.y; hv[4 * c + 2] = v.z; hv[4 * c + 3] = v.w;
                            }
                        }
                    }
                    if (warp == 4)
                        mbar_wait<32>(bar(ACC_FULL + ab), (acc_n >> 1) & 1u);
                    named_bar_sync(1, EPI_THREADS);
                    if (q == 0) {
                        if (et < H)
                            bias_s[et] = bias_reg;
                        named_bar_sync(2, EPI_THREADS);
                    }
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + P::TMEM_ACC0 + ab * NQ + sl * 32 + lane_addr;
                    if (is_proj) {
                        // z_e = (h . Wp_hi) + (h . Wp_lo) + bias: accumulator columns [0, 64) and [64, 128);
                        // slabs 0 and 1 own output columns [0, 32) and [32, 64), slabs 2 and 3 only hand the buffer back
                        if (sl < 2 && sl * 32 < proj_d) {
#pragma unroll
                            for (int hf = 0; hf < 2; ++hf) {          // 16 output columns at a time keeps 32 registers live
                                uint32_t vh[16], vl[16];
                                tmem_ld16(taddr + 16 * hf, vh);
                                tmem_ld16(taddr + 64 + 16 * hf, vl);
                                tmem_wait_ld16(vh, vl);
                                if (hf == 1) {
                                    tc_fence_before();
                                    mbar_arrive(bar(ACC_EMPTY + ab));
                                }
                                if (row_ok) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) {
                                        const int col = sl * 32 + 16 * hf + 4 * c;
                                        if (col < proj_d) {              // proj_d is a multiple of 4
                                            const float4 b = *reinterpret_cast<const float4 *>(bias_s + col);
                                            float4 o;
                                            o.x = (__uint_as_float(vh[4 * c]) + __uint_as_float(vl[4 * c])) + b.x;
                                            o.y = (__uint_as_float(vh[4 * c + 1]) + __uint_as_float(vl[4 * c + 1])) + b.y;
                                            o.z = (__uint_as_float(vh[4 * c + 2]) + __uint_as_float(vl[4 * c + 2])) + b.z;
                                            o.w = (__uint_as_float(vh[4 * c + 3]) + __uint_as_float(vl[4 * c + 3])) + b.w;
                                            *reinterpret_cast<float4 *>(z_e + row * proj_d + col) = o;
                                        }
                                    }
                                }
                            }
                        } else {
                            tc_fence_before();
                            mbar_arrive(bar(ACC_EMPTY + ab));
                        }
                        continue;
                    }
                    // the slab in two halves of 16 columns (keeps the packed-pair registers within the 96 of a 640-thread CTA)
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        uint32_t v[16];
                        tmem_ld16(taddr + 16 * hf, v);
                        tmem_wait_ld16_one(v);
                        if (hf == 1) {
                            tc_fence_before();
                            mbar_arrive(bar(ACC_EMPTY + ab));    // the slab sits in registers: this thread is done with the accumulator
                        }
                        const int colh = col0 + 16 * hf;
#if EC_ABLATE == 1
                        if (!last) {
                            uint32_t pk[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                pk[c] = v[c] ^ v[c + 8];
                            put_operand(to_tmem, colh, pk);
                        }
                        continue;
#endif
                        float2 x[8];                             // 8 pairs (packed fp32 arithmetic)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float4 b = *reinterpret_cast<const float4 *>(bias_s + colh + 4 * c);
                            x[2 * c] = __fadd2_rn(make_float2(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1])),
                                                  make_float2(b.x, b.y));
                            x[2 * c + 1] = __fadd2_rn(make_float2(__uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])),
                                                      make_float2(b.z, b.w));
                        }
                        if (resid) {
#pragma unroll
                            for (int c = 0; c < 8; ++c)
                                x[c] = __fadd2_rn(x[c], make_float2(hv[16 * hf + 2 * c], hv[16 * hf + 2 * c + 1]));
                        }
                        if (resid || is_pre) {                   // the residual stream: h0 after PRE, h after every block
                            if (h_out_std) {
                                if (row_ok) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c)
                                        __stcs(reinterpret_cast<float4 *>(h + row * H + colh) + c,
                                               make_float4(x[2 * c].x, x[2 * c].y, x[2 * c + 1].x, x[2 * c + 1].y));
                                }
                            } else {
#pragma unroll
                                for (int c = 0; c < 4; ++c)
                                    *reinterpret_cast<float4 *>(scr + ((size_t)((colh >> 2) + c) * BM + r) * 4) =
                                        make_float4(x[2 * c].x, x[2 * c].y, x[2 * c + 1].x, x[2 * c + 1].y);
                            }
                        }
                        if (!last) {
                            uint32_t pk[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) {
                                const float2 y = plain ? x[c] : gelu2(x[c]);
                                const __nv_bfloat162 p2 = __floats2bfloat162_rn(y.x, y.y);
                                pk[c] = *reinterpret_cast<const uint32_t *>(&p2);
                            }
                            put_operand(to_tmem, colh, pk);
                        }
                    }
                    if (!last)
                        publish(to_tmem, A_RDY + q);
                }
                if (plain) {
                    // ---- PROJ's second operand: bf16(h - bf16(h)) of the whole row into the buffer the last layer's MMAs
                    // have finished reading (its last accumulator was waited for above), h from the scratch tile
#pragma unroll 1
                    for (int q = 0; q < NQT; ++q) {
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            const int colh = q * NQ + sl * 32 + 16 * hf;
                            uint32_t pk[8];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const float4 v = *reinterpret_cast<const float4 *>(scr + ((size_t)((colh >> 2) + c) * BM + r) * 4);
                                const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
                                const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
                                const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
                                pk[2 * c] = *reinterpret_cast<const uint32_t *>(&l01);
                                pk[2 * c + 1] = *reinterpret_cast<const uint32_t *>(&l23);
                            }
                            put_operand(!to_tmem, colh, pk);
                        }
                    }
                    publish(!to_tmem, LO_RDY);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------------------
// Operand of the PRE GEMM: token t of cycle b is channel c = t / (L/P), patch p = t % (L/P) (the reference's
// channel-major order, model/vq_vae_patch_embedd.py:14-15) with samples x[b][p*P + k][c]; out[token] =
// [bf16(x_k) k < P, 0 ... | bf16(x_k - bf16(x_k)), 0 ...] as 32 + 32 bf16 (128 bytes per token).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vq_patch_split_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ out,
                                                             int64_t n_tokens, int L, int C, int P)
{
    const int ppc = L / P, T = ppc * C;
    // one thread per (token, k): 32 threads per token write the hi and the lo half of one row
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_tokens * 32; i += (int64_t)gridDim.x * 256) {
        const int64_t t = i >> 5;
        const int k = (int)(i & 31);
        float v = 0.0f;
        if (k < P) {
            const int64_t b = t / T;
            const int ct = (int)(t - b * T), c = ct / ppc, pp = ct - c * ppc;
            v = __ldg(x + (b * L + (int64_t)pp * P + k) * C + c);
        }
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        out[t * 64 + k] = hi;
        out[t * 64 + 32 + k] = __float2bfloat16_rn(v - __bfloat162float(hi));
    }
}

bool enc_chain_supported(int H, int L) { return (H == 512 || H == 256) && L >= 2 && L % 2 == 0 && L <= 64; }
bool patch_split_supported(int L, int C, int P) { return P >= 1 && P <= 32 && C >= 1 && L >= P && L % P == 0; }

size_t enc_chain_scratch_bytes(int H, int sm_count) { return (size_t)sm_count * ec::BM * H * sizeof(float); }

cudaError_t launch_patch_split(const float *x, void *out, int64_t n_cycles, int L, int C, int P, int sm_count, cudaStream_t st)
{
    if (!patch_split_supported(L, C, P))
        return cudaErrorNotSupported;
    const int64_t n_tokens = n_cycles * (L / P) * C;
    if (n_tokens == 0)
        return cudaSuccess;
    const int64_t blocks = (n_tokens * 32 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)sm_count * 16 ? blocks : (int64_t)sm_count * 16);
    vq_patch_split_kernel<<<grid, 256, 0, st>>>(x, (__nv_bfloat16 *)out, n_tokens, L, C, P);
    return cudaGetLastError();
}

cudaError_t launch_enc_chain(const void *a0, float *h, const void *w, const float *bias, int64_t n_tokens, int H, int L,
                             float *scratch, size_t scratch_bytes, const float *proj_bias, float *z_e, int proj_d,
                             const float *pre_bias, int sm_count, int max_smem, cudaStream_t st)
{
    using namespace ec;
    if (!enc_chain_supported(H, L))
        return cudaErrorNotSupported;
    if (proj_d != 0 && (proj_d < 4 || proj_d > 64 || proj_d % 4 != 0 || !proj_bias || !z_e))
        return cudaErrorNotSupported;
    if ((!pre_bias || !proj_d) && !h)
        return cudaErrorInvalidValue;
    if (n_tokens == 0)
        return cudaSuccess;
    if (n_tokens >= (1ll << 31))
        return cudaErrorNotSupported;
    const int64_t tiles = (n_tokens + BM - 1) / BM;
    const int grid = (int)(tiles < sm_count ? tiles : sm_count);
    if (scratch_bytes < (size_t)grid * BM * H * sizeof(float))
        return cudaErrorInvalidValue;
    const int smem_bytes = H == 512 ? Plan<512>::SMEM_BYTES : Plan<256>::SMEM_BYTES;
    if (smem_bytes > max_smem)
        return cudaErrorNotSupported;
    CUtensorMap map_a0, map_w;
    const int64_t w_rows = (int64_t)L * H + (proj_d ? NQ : 0) + (pre_bias ? H : 0);
    if (!tc::make_tensor_map_2d(&map_a0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a0, n_tokens, pre_bias ? BK : H, BM, BK,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) ||
        !tc::make_tensor_map_2d(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, w_rows, H, NQ, BK,
                                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
        return cudaErrorNotSupported;
    auto kern = H == 512 ? enc_chain_kernel<512> : enc_chain_kernel<256>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (err != cudaSuccess)
        return err;
    kern<<<grid, THREADS, smem_bytes, st>>>(map_a0, map_w, bias, h, scratch, n_tokens, L, proj_bias, z_e, proj_d, pre_bias);
    return cudaGetLastError();
}

}  // namespace vqb
