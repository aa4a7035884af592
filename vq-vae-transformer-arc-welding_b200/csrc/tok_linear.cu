// tok_linear.cu -- per-token linear layer of the patch encoder with its element-wise neighbours fused
// (sm_100a, tcgen05 + TMA + TMEM).
//
// The reference's encoder (model/vq_vae_patch_embedd.py:60-74, 103-111) applies, to every token
// independently, eight residual blocks  h <- h + W2 gelu(W1 gelu(h) + b1) + b2  with 512 x 512 weights
// (the centre tap of Conv1d(k=3, pad=1) on a length-1 slice).  Run as stock PyTorch ops that is 16 GEMMs
// plus ~40 element-wise passes over the (tokens x 512) activations, and the passes -- not the GEMMs --
// bound it (profiles/README.md: bulk encoding).  This kernel is one of those GEMMs with the passes around it
// folded into its epilogue:
//     mode 0:  out = bf16( gelu( A W^T + b ) )                      (first GEMM of a block)
//     mode 1:  h  += A W^T + b  (fp32, in place);  out = bf16( gelu(h) )   (second GEMM + residual + the
//              next block's leading GELU; out may be NULL after the last block)
//     mode 2:  h   = A W^T + b  (fp32, written);   out = bf16( gelu(h) )   (patch embedding, K zero-padded to 64,
//              + the first block's leading GELU)
// Three-tap form (taps = 3; the DECODER's Conv1d(k = 3, pad = 1) along the 16 positions of a cycle,
// model/vq_vae_patch_embedd.py:60-74,142-147): out[t] = sum_k W_k a[t + k - 1] with zeros beyond the ends of the
// token's cycle.  Same kernel, K = 3 * K_in: the A operand of K-chunk (tap, c) is the TMA box at token offset tap - 1
// of a 3-D map (K_in, tokens per cycle, cycles), whose out-of-range positions arrive as zeros -- no im2col copy.
// out_gelu = 0 writes out = bf16(x) instead of bf16(gelu(x)) (the last block feeds the transposed convolutions).
// A: (T, K) bf16 row-major activations, W: (N, K) bf16 row-major, fp32 accumulation in TMEM, erf-form GELU
// evaluated through a fitted tanh argument (gelu_fast below, |error| <= 2.5e-5 + 2.5e-4 |x|, then rounded to bf16).  bf16 operands make this the reduced-precision encoder mode
// (same operand precision as the reference under its own torch.set_float32_matmul_precision('medium'));
// the quantiser behind it stays exact.
//
// Split form (SPLIT = true; vqb_token_linear_split, the fp32-faithful encoder mode): every operand is an exact-to-2^-17 PAIR
// of bf16 values, A = [a_hi | a_lo] (T, 2 K_in) and W = [w_hi | w_lo] (N, 2 K_in) with hi = bf16(x), lo = bf16(x - hi); the
// kernel accumulates the three products a_hi w_hi + a_hi w_lo + a_lo w_hi in the fp32 accumulator (K-chunk 3c + p of the
// K = 3 K_in sweep reads the A box at column 64c + (p == 2) K_in and the W box at column 64c + (p == 1) K_in; the dropped
// a_lo w_lo term is below 2^-16 of |a||w|), evaluates GELU in its erf form (gelu_erf: erff, ~1e-7), and writes `out` as the
// same kind of pair, (T, 2 N) = [bf16(g) | bf16(g - bf16(g))].  Three times the tensor work of the bf16 form for 2^-16
// instead of 2^-8 operand precision: ids equal to the fp32 encoder's except on ~1e-5 of the tokens.
//
// Tile 128 tokens x 256 outputs, K streamed in 64-element chunks through a 3-stage TMA ring (A 16 KB + W 32 KB per
// stage, SWIZZLE_128B; tl::Plan), 4 tcgen05.mma (128 x 256 x 16) per chunk into one of two 256-column TMEM accumulators,
// eight epilogue warps (TMEM lane quarter x column half) that overlap with the next tile's MMAs.
#include <cstdlib>
#include "vq_common.cuh"
#include "vq_ptx.cuh"

namespace vqb {

namespace tl {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
constexpr int THREADS = 128 + 256;                       // 4 service warps + 8 epilogue warps (EW = 16: 128 + 512)
// Shared-memory plan per mode: three operand stages; modes 1 / 2 give every epilogue warp two 4 KB transposing buffers
// (residual slab prefetch), mode 0 needs one.  (A fourth stage for mode 0 fits in 227 KB and was measured: 0.63 ms against
// 0.58-0.61 ms -- with all of the SM's memory carved out as shared memory there is no L1 left for the bias / descriptor loads.)
// EW = 16 epilogue warps (TMEM lane quarter x column QUARTER, two slabs each): one 4 KB buffer per warp in every mode.
template <int MODE, int EW = 8> struct Plan {
    static constexpr int STAGES = 3;
    static constexpr int XPOSE_PER_WARP = (MODE == 0 || EW == 16) ? 4096 : 8192;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
    static constexpr int OFF_XPOSE = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_BARS = OFF_XPOSE + EW * XPOSE_PER_WARP;
    static constexpr int SMEM_BYTES = OFF_BARS + 256;
};

// GELU(x) = x Phi(x) as 0.5 x (1 + tanh(x (c1 + c3 x^2 + c5 x^4))): the odd polynomial is a minimax fit of
// atanh(erf(x / sqrt 2)) (max |error| of the formula 2.5e-5, against 4.7e-4 for the textbook two-term "tanh GELU");
// MUFU.TANH adds up to 2^-11 relative on the tanh, i.e. <= 2.5e-4 |x| on the result -- below a tenth of the bf16
// spacing of the value it is rounded to; x^2 is clamped to 36 because the fitted polynomial turns over beyond |x| ~ 10
// (tanh is +-1 there to fp32 accuracy).  7 FMA-pipe instructions + 1 FMNMX + 1 MUFU per element: the A&S erf form (2 MUFU, ~30
// instructions with IEEE reciprocal / exp) made the epilogue 2.6x longer than the tile's MMAs (profiles/README.md).
__device__ __forceinline__ float gelu_fast(float x)
{
    const float u = fminf(x * x, 36.0f);                 // the fit covers |x| <= 6; beyond, tanh(1.67 x) is +-1 anyway
    float p = fmaf(-3.51517534e-4f, u, 3.70056510e-2f);
    p = fmaf(p, u, 7.97507878e-1f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * p));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}

// GELU to fp32 accuracy for the split (fp32-faithful) layers: x Phi(x) = 0.5 x (1 + erf(x / sqrt 2)) with
// erf(|x| / sqrt 2) = 1 - 2^(-u P(u)), u = min(|x|, 6.1): ONE branch-free formula -- P is a degree-7 fit of
// -log2(erfc(u / sqrt 2)) / u, weighted for the absolute error of the erfc value (Lawson iterations on 6000 Chebyshev nodes,
// fit error 1e-9; beyond u = 6.1 erfc is below 2^-29).  In fp32 arithmetic the formula is within 1.1e-7 |x| of the exact GELU --
// the same as torch's own fp32 erf form (1.0e-7 |x|, dominated by the rounding of 1 + erf) -- plus <= 1.2e-7 |x| from
// ex2.approx: 15x below the 2^-18 precision of the bf16 pair the value is stored as.  15 instructions per element; erff
// (two coefficient sets chosen by FSEL, ~40 instructions) left the 8 epilogue warps behind the split layer's MMAs.
__device__ __forceinline__ float gelu_erf(float x)
{
    const float u = fminf(fabsf(x), 6.1f);
    float p = fmaf(2.834913857e-06f, u, -3.937768997e-05f);
    p = fmaf(p, u, 1.861801720e-04f);
    p = fmaf(p, u, 1.369366655e-04f);
    p = fmaf(p, u, -7.063420489e-03f);
    p = fmaf(p, u, 5.249617994e-02f);
    p = fmaf(p, u, 4.592081904e-01f);
    p = fmaf(p, u, 1.151105165e+00f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * p));
    const float erf_abs = 1.0f - e;
    const float erf_x = __uint_as_float(__float_as_uint(erf_abs) | (__float_as_uint(x) & 0x80000000u));
    const float hx = 0.5f * x;
    return fmaf(hx, erf_x, hx);
}

}  // namespace tl

template <int MODE, bool SPLIT, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
tok_linear_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const float *__restrict__ bias, float *__restrict__ h, __nv_bfloat16 *__restrict__ out,
                  int64_t n_tokens, int K, int N, int taps, int cyc_len, int out_gelu)
{
    using namespace tc;
    using namespace tl;
    using P = Plan<MODE, EW>;
    constexpr int STAGES = P::STAGES, OFF_A = P::OFF_A, OFF_B = P::OFF_B, OFF_XPOSE = P::OFF_XPOSE, OFF_BARS = P::OFF_BARS;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    if ((sbase & 1023u) != 0)
        __trap();
    enum { FULL = 0, EMPTY = FULL + STAGES, T_FULL = EMPTY + STAGES, T_EMPTY = T_FULL + 2, N_BARS = T_EMPTY + 2 };
    auto bar = [&](int i) { return sbase + OFF_BARS + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + 8 * N_BARS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int n_ntiles = N / BN;
    const int64_t n_mtiles = (n_tokens + BM - 1) / BM;
    const int64_t n_items = n_mtiles * n_ntiles;            // item = (m tile, n tile), n fastest: A is re-read from L2
    const int64_t my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_k = K / BK;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(FULL + s), 1);
            mbar_init(bar(EMPTY + s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(T_FULL + b), 1);
            mbar_init(bar(T_EMPTY + b), 32 * EW);
        }
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
        // ================= TMA producer (warp-uniform loop, the loads under elect_one) =================
        {
            int64_t step = 0;
            for (int64_t it = 0; it < my_items; ++it) {
                const int64_t item = blockIdx.x + it * gridDim.x;
                const int64_t mt = item / n_ntiles;
                const int nt = (int)(item % n_ntiles);
                for (int k = 0; k < n_k; ++k, ++step) {
                    const int s = (int)(step % STAGES);
                    mbar_wait<32>(bar(EMPTY + s), (uint32_t)(((step / STAGES) & 1) ^ 1));
                    if (elect_one()) {
                        mbar_expect_tx(bar(FULL + s), A_BYTES + B_BYTES);
                        if (SPLIT) {      // K-chunk 3c + p: (a_hi, w_hi), (a_hi, w_lo), (a_lo, w_hi) of input columns 64c ..
                            const int c3 = k / 3, p = k - 3 * c3, k_in = K / (3 * taps);
                            if (taps == 1) {
                                tma_load_2d(sbase + OFF_A + s * A_BYTES, &map_a, bar(FULL + s), c3 * BK + (p == 2 ? k_in : 0),
                                            (int)(mt * BM));
                                tma_load_2d(sbase + OFF_B + s * B_BYTES, &map_w, bar(FULL + s), c3 * BK + (p == 1 ? k_in : 0),
                                            nt * BN);
                            } else {      // three taps: W rows are [tap][w_hi | w_lo], the A box is shifted by tap - 1 positions
                                const int kc = k_in / BK, tap = c3 / kc, c = c3 - tap * kc;
                                tma_load_3d(sbase + OFF_A + s * A_BYTES, &map_a, bar(FULL + s), c * BK + (p == 2 ? k_in : 0),
                                            tap - (taps >> 1), (int)(mt * (BM / cyc_len)));
                                tma_load_2d(sbase + OFF_B + s * B_BYTES, &map_w, bar(FULL + s),
                                            tap * 2 * k_in + c * BK + (p == 1 ? k_in : 0), nt * BN);
                            }
                        } else if (taps == 1) {
                            tma_load_2d(sbase + OFF_A + s * A_BYTES, &map_a, bar(FULL + s), k * BK, (int)(mt * BM));
                        } else {          // K-chunk (tap, c): the tile's cycles, shifted by tap - 1 positions (zeros outside)
                            const int kin = n_k / taps, tap = k / kin, c = k - tap * kin;
                            tma_load_3d(sbase + OFF_A + s * A_BYTES, &map_a, bar(FULL + s), c * BK, tap - (taps >> 1),
                                        (int)(mt * (BM / cyc_len)));
                        }
                        if (!SPLIT)
                            tma_load_2d(sbase + OFF_B + s * B_BYTES, &map_w, bar(FULL + s), k * BK, nt * BN);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp runs the loop, one elected lane issues (vq_ptx.cuh: elect_one) ====
        {
            const uint32_t idesc = idesc_bf16(BN);
            const uint64_t a_base = desc_sw128(sbase + OFF_A), w_base = desc_sw128(sbase + OFF_B);
            int64_t step = 0;
            for (int64_t it = 0; it < my_items; ++it) {
                const int b = (int)(it & 1);
                mbar_wait<32>(bar(T_EMPTY + b), (uint32_t)(((it >> 1) & 1) ^ 1));
                tc_fence_after();
                const uint32_t d = tmem_base + b * BN;
                for (int k = 0; k < n_k; ++k, ++step) {
                    const int s = (int)(step % STAGES);
                    mbar_wait<32>(bar(FULL + s), (uint32_t)((step / STAGES) & 1));
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a = a_base + (uint64_t)(s * (A_BYTES >> 4));
                        const uint64_t w = w_base + (uint64_t)(s * (B_BYTES >> 4));
#pragma unroll
                        for (int j = 0; j < BK / 16; ++j)    // K-slices of 16 bf16 = 32 bytes = +2 in the address field
                            umma_bf16(d, a + 2 * j, w + 2 * j, idesc, (k | j) != 0);
                        umma_commit(bar(EMPTY + s));
                    }
                    __syncwarp();
                }
                if (elect_one())
                    umma_commit(bar(T_FULL + b));
                __syncwarp();
            }
        }
    } else if (EW == 16 && warp >= 4) {
        // ================= epilogue, 16 warps: TMEM lane quarter x column quarter (64 columns = two 32-column slabs) ==========
        // Four warps per scheduler instead of two hide the epilogue's dependent chains (the 8-warp epilogue issued at ~46 % with
        // both of a scheduler's warps stalled on their own previous instruction most of the time).  640 threads leave 96
        // registers per thread: a slab is worked on in two halves of 16 columns, and a warp has one 4 KB buffer (the residual
        // slab of the NEXT slab is fetched once this slab's stores have left the buffer).
        const int q = warp & 3;
        const int cq = (warp - 4) >> 2;
        unsigned char *xp = smem + OFF_XPOSE + (warp - 4) * 4096;
        auto prefetch_h = [&](int64_t row0, int col0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + (lane >> 3), cc = lane & 7;
                const bool in = row0 + rr < n_tokens;
                const float *src = in ? h + (row0 + rr) * N + col0 + 4 * cc : h;
                const uint32_t dst = smem_u32(xp + rr * 128 + ((cc ^ (rr & 7)) << 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int64_t it = 0; it < my_items; ++it) {
            const int b = (int)(it & 1);
            const int64_t item = blockIdx.x + it * gridDim.x;
            const int64_t mt = item / n_ntiles;
            const int nt = (int)(item % n_ntiles);
            const int64_t row0 = mt * BM + q * 32;
            if (MODE == 1)
                prefetch_h(row0, nt * BN + cq * 64);
            if (warp == 4)
                mbar_wait<32>(bar(T_FULL + b), (uint32_t)((it >> 1) & 1));
            asm volatile("bar.sync 1, 512;" ::: "memory");
            tc_fence_after();
            const uint32_t taddr = tmem_base + b * BN + cq * 64 + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int sl = 0; sl < 2; ++sl) {
                const int col0 = nt * BN + cq * 64 + sl * 32;
                uint32_t packed[16];
                uint32_t packed_lo[SPLIT ? 16 : 1];
                if (MODE == 1) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t v[16];
                    tmem_ld16(taddr + sl * 32 + half * 16, v);
                    float hv[16], bv[16];
                    if (MODE == 1) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int ck = half * 4 + c;
                            const float4 t = *reinterpret_cast<const float4 *>(xp + lane * 128 + ((ck ^ (lane & 7)) << 4));
                            hv[4 * c] = t.x; hv[4 * c + 1] = t.y; hv[4 * c + 2] = t.z; hv[4 * c + 3] = t.w;
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(bias + col0 + half * 16) + c);
                        bv[4 * c] = t.x; bv[4 * c + 1] = t.y; bv[4 * c + 2] = t.z; bv[4 * c + 3] = t.w;
                    }
                    tmem_wait_ld_fence16(v);
#pragma unroll
                    for (int c = 0; c < 16; c += 2) {
                        float x0 = __uint_as_float(v[c]) + bv[c];
                        float x1 = __uint_as_float(v[c + 1]) + bv[c + 1];
                        if (MODE == 1) {
                            x0 += hv[c];
                            x1 += hv[c + 1];
                        }
                        if (MODE != 0) {
                            hv[c] = x0;
                            hv[c + 1] = x1;
                        }
                        if (SPLIT) {
                            const float g0 = out_gelu ? gelu_erf(x0) : x0, g1 = out_gelu ? gelu_erf(x1) : x1;
                            const __nv_bfloat162 pk = __floats2bfloat162_rn(g0, g1);
                            const __nv_bfloat162 pl = __floats2bfloat162_rn(g0 - __low2float(pk), g1 - __high2float(pk));
                            packed[half * 8 + (c >> 1)] = *reinterpret_cast<const uint32_t *>(&pk);
                            packed_lo[half * 8 + (c >> 1)] = *reinterpret_cast<const uint32_t *>(&pl);
                        } else {
                            const __nv_bfloat162 pk = out_gelu ? __floats2bfloat162_rn(gelu_fast(x0), gelu_fast(x1))
                                                               : __floats2bfloat162_rn(x0, x1);
                            packed[half * 8 + (c >> 1)] = *reinterpret_cast<const uint32_t *>(&pk);
                        }
                    }
                    if (MODE != 0) {        // (own row, own chunks: no other lane reads or writes them in this phase)
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int ck = half * 4 + c;
                            *reinterpret_cast<float4 *>(xp + lane * 128 + ((ck ^ (lane & 7)) << 4)) =
                                make_float4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
                        }
                    }
                }
                if (MODE != 0) {
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + (lane >> 3), cc = lane & 7;
                        const float4 t = *reinterpret_cast<const float4 *>(xp + rr * 128 + ((cc ^ (rr & 7)) << 4));
                        if (row0 + rr < n_tokens)
                            *reinterpret_cast<float4 *>(h + (row0 + rr) * N + col0 + 4 * cc) = t;
                    }
                }
                if (out) {
                    const int64_t ld_out = SPLIT ? 2 * (int64_t)N : (int64_t)N;
#pragma unroll
                    for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
                        const uint32_t *pk = part == 0 ? packed : packed_lo;
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<uint4 *>(xp + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                                make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = 8 * i + (lane >> 2), cc = lane & 3;
                            const uint4 t = *reinterpret_cast<const uint4 *>(xp + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
                            if (row0 + rr < n_tokens)
                                *reinterpret_cast<uint4 *>(out + (row0 + rr) * ld_out + part * N + col0 + 8 * cc) = t;
                        }
                    }
                }
                if (MODE == 1 && sl == 0) {
                    __syncwarp();                         // every lane has taken its stores' data out of the buffer
                    prefetch_h(row0, col0 + 32);
                }
            }
            tc_fence_before();
            mbar_arrive(bar(T_EMPTY + b));
        }
    } else if (EW == 8 && warp >= 4) {
        // ================= epilogue: TMEM -> bias (+ residual) -> GELU -> bf16 =================
        const int q = warp & 3;                           // TMEM lane quarter
        const int ch = (warp - 4) >> 2;                   // column half of the 256-column accumulator
        unsigned char *xp0 = smem + OFF_XPOSE + (warp - 4) * P::XPOSE_PER_WARP;
        // cp.async of the residual slab `sl` of the current item into buffer `buf` (lane -> row 4i + lane/8, chunk lane%8;
        // rows beyond the tensor are zero-filled): in flight while the tile's MMAs / the previous slab are worked on
        auto prefetch_h = [&](int64_t row0, int col0, int buf) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = 4 * i + (lane >> 3), cc = lane & 7;
                const bool in = row0 + rr < n_tokens;
                const float *src = in ? h + (row0 + rr) * N + col0 + 4 * cc : h;
                const uint32_t dst = smem_u32(xp0 + buf * 4096 + rr * 128 + ((cc ^ (rr & 7)) << 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(in ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int64_t it = 0; it < my_items; ++it) {
            const int b = (int)(it & 1);
            const int64_t item = blockIdx.x + it * gridDim.x;
            const int64_t mt = item / n_ntiles;
            const int nt = (int)(item % n_ntiles);
            if (MODE == 1)
                prefetch_h(mt * BM + q * 32, nt * BN + ch * 128, 0);
            if (warp == 4)                                // one warp polls, the other seven block on a named barrier
                mbar_wait<32>(bar(T_FULL + b), (uint32_t)((it >> 1) & 1));
            asm volatile("bar.sync 1, 256;" ::: "memory");
            tc_fence_after();
            const uint32_t taddr = tmem_base + b * BN + ch * 128 + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
            for (int sl = 0; sl < 4; ++sl) {
                uint32_t v[32];
                tmem_ld32(taddr + sl * 32, v);
                tmem_wait_ld_fence(v);
                const int col0 = nt * BN + ch * 128 + sl * 32;
                // TMEM hands every thread one ROW of the slab (32 columns); global memory wants every warp instruction
                // to cover whole 128-byte row segments.  The slab is therefore turned through a warp-private 4 KB
                // buffer (16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4): conflict-free both ways):
                // global <-> buffer moves use lane -> (row 4i + lane/8, chunk lane%8), i.e. 4 full lines per instruction.
                float hv[32];
                const int64_t row0 = mt * BM + q * 32;
                unsigned char *xp = xp0 + (MODE == 1 ? (sl & 1) * 4096 : 0);   // (EW == 8: P::XPOSE_PER_WARP is 8 KB in mode 1)
                if (MODE == 1) {
                    __syncwarp();                         // every lane is done with the buffer the next slab lands in
                    if (sl < 3) {
                        prefetch_h(row0, col0 + 32, (sl + 1) & 1);
                        asm volatile("cp.async.wait_group 1;" ::: "memory");
                    } else {
                        asm volatile("cp.async.wait_group 0;" ::: "memory");
                    }
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 t = *reinterpret_cast<const float4 *>(xp + lane * 128 + ((c ^ (lane & 7)) << 4));
                        hv[4 * c] = t.x; hv[4 * c + 1] = t.y; hv[4 * c + 2] = t.z; hv[4 * c + 3] = t.w;
                    }
                }
                uint32_t packed[16];
                uint32_t packed_lo[SPLIT ? 16 : 1];
                float bv[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 t = __ldg(reinterpret_cast<const float4 *>(bias + col0) + c);
                    bv[4 * c] = t.x; bv[4 * c + 1] = t.y; bv[4 * c + 2] = t.z; bv[4 * c + 3] = t.w;
                }
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    float x0 = __uint_as_float(v[c]) + bv[c];
                    float x1 = __uint_as_float(v[c + 1]) + bv[c + 1];
                    if (MODE == 1) {
                        x0 += hv[c];
                        x1 += hv[c + 1];
                    }
                    if (MODE != 0) {
                        hv[c] = x0;
                        hv[c + 1] = x1;
                    }
                    if (SPLIT) {
                        const float g0 = out_gelu ? gelu_erf(x0) : x0, g1 = out_gelu ? gelu_erf(x1) : x1;
                        const __nv_bfloat162 pk = __floats2bfloat162_rn(g0, g1);
                        const __nv_bfloat162 pl = __floats2bfloat162_rn(g0 - __low2float(pk), g1 - __high2float(pk));
                        packed[c >> 1] = *reinterpret_cast<const uint32_t *>(&pk);
                        packed_lo[c >> 1] = *reinterpret_cast<const uint32_t *>(&pl);
                    } else {
                        const __nv_bfloat162 pk = out_gelu ? __floats2bfloat162_rn(gelu_fast(x0), gelu_fast(x1))
                                                           : __floats2bfloat162_rn(x0, x1);
                        packed[c >> 1] = *reinterpret_cast<const uint32_t *>(&pk);
                    }
                }
                if (MODE != 0) {
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4 *>(xp + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                            make_float4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = 4 * i + (lane >> 3), cc = lane & 7;
                        const float4 t = *reinterpret_cast<const float4 *>(xp + rr * 128 + ((cc ^ (rr & 7)) << 4));
                        if (row0 + rr < n_tokens)
                            *reinterpret_cast<float4 *>(h + (row0 + rr) * N + col0 + 4 * cc) = t;
                    }
                }
                if (out) {
                    // bf16 rows are 64 bytes: chunk c of row r at r*64 + ((c ^ ((r >> 1) & 3)) << 4); global moves use
                    // lane -> (row 8i + lane/4, chunk lane%4): 8 row segments of 64 bytes per instruction.
                    // Split form: rows of `out` are 2 N wide, the hi slab goes to column col0, the lo slab to N + col0.
                    const int64_t ld_out = SPLIT ? 2 * (int64_t)N : (int64_t)N;
#pragma unroll
                    for (int part = 0; part < (SPLIT ? 2 : 1); ++part) {
                        const uint32_t *pk = part == 0 ? packed : packed_lo;
                        __syncwarp();
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<uint4 *>(xp + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
                                make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int rr = 8 * i + (lane >> 2), cc = lane & 3;
                            const uint4 t = *reinterpret_cast<const uint4 *>(xp + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
                            if (row0 + rr < n_tokens)
                                *reinterpret_cast<uint4 *>(out + (row0 + rr) * ld_out + part * N + col0 + 8 * cc) = t;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar(T_EMPTY + b));
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// h += bias (fp32, in place); out = bf16(gelu(h)) -- the element-wise step between the fp32 patch embedding GEMM
// (K = 25, stock cuBLAS without bias) and the first fused layer: one pass instead of bias + GELU + cast (three).
__global__ void __launch_bounds__(256) tok_bias_gelu_kernel(float *__restrict__ h, const float *__restrict__ bias,
                                                            __nv_bfloat16 *__restrict__ out, int64_t n_vec4, int n4)
{
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < n_vec4; t += (int64_t)gridDim.x * 256) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + (int)(t % n4));
        float4 v = reinterpret_cast<float4 *>(h)[t];
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        reinterpret_cast<float4 *>(h)[t] = v;
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(tl::gelu_fast(v.x), tl::gelu_fast(v.y));
        const __nv_bfloat162 p1 = __floats2bfloat162_rn(tl::gelu_fast(v.z), tl::gelu_fast(v.w));
        reinterpret_cast<uint2 *>(out)[t] = make_uint2(*reinterpret_cast<const uint32_t *>(&p0),
                                                       *reinterpret_cast<const uint32_t *>(&p1));
    }
}

cudaError_t launch_tok_bias_gelu(float *h, const float *bias, void *out, int64_t n_tokens, int N, int sm_count,
                                 cudaStream_t st)
{
    if (N % 4 != 0)
        return cudaErrorNotSupported;
    const int64_t n_vec4 = n_tokens * (N / 4);
    if (n_vec4 == 0)
        return cudaSuccess;
    const int64_t blocks = (n_vec4 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)sm_count * 16 ? blocks : (int64_t)sm_count * 16);
    tok_bias_gelu_kernel<<<grid, 256, 0, st>>>(h, bias, (__nv_bfloat16 *)out, n_vec4, N / 4);
    return cudaGetLastError();
}

// out = [bf16(g) | bf16(g - bf16(g))] (T, 2 N) with g = gelu_erf(h) or h itself: the operand pair the split layers read, made
// from a fp32 (T, N) tensor (the patch embedding's output in front of the first block).  4 N bytes in, 4 N out per token.
__global__ void __launch_bounds__(256) tok_pair_kernel(const float *__restrict__ h, __nv_bfloat16 *__restrict__ out,
                                                       int64_t n_vec4, int n4, int apply_gelu)
{
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < n_vec4; t += (int64_t)gridDim.x * 256) {
        float4 v = __ldcs(reinterpret_cast<const float4 *>(h) + t);
        if (apply_gelu) {
            v.x = tl::gelu_erf(v.x); v.y = tl::gelu_erf(v.y); v.z = tl::gelu_erf(v.z); v.w = tl::gelu_erf(v.w);
        }
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
        const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
        const int64_t row = t / n4;
        const int c4 = (int)(t - row * n4);
        uint2 *o = reinterpret_cast<uint2 *>(out) + row * (2 * n4) + c4;
        o[0] = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
        o[n4] = make_uint2(*reinterpret_cast<const uint32_t *>(&l0), *reinterpret_cast<const uint32_t *>(&l1));
    }
}

cudaError_t launch_tok_pair(const float *h, void *out, int64_t n_tokens, int N, int apply_gelu, int sm_count, cudaStream_t st)
{
    if (N % 4 != 0)
        return cudaErrorNotSupported;
    const int64_t n_vec4 = n_tokens * (N / 4);
    if (n_vec4 == 0)
        return cudaSuccess;
    const int64_t blocks = (n_vec4 + 255) / 256;
    const int grid = (int)(blocks < (int64_t)sm_count * 16 ? blocks : (int64_t)sm_count * 16);
    tok_pair_kernel<<<grid, 256, 0, st>>>(h, (__nv_bfloat16 *)out, n_vec4, N / 4, apply_gelu);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Patch embedding (model/vq_vae_patch_embedd.py:13-17): channel-major patchify + linear P -> H + bias, fused with the
// first residual block's leading GELU / bf16 cast.  fp32 FMA arithmetic (K = P = 25 is 0.3 % of the encoder's FLOPs
// and the raw signal in bf16 costs index matches, see encode_fused_bf16): the point is the memory traffic -- the
// stock sequence (permute copy, SGEMM, bias, GELU, cast) moves ~9 KB per token, this kernel reads 4P bytes and
// writes 4H + 2H.
//   x   (n_cycles, L, C) fp32 contiguous; token t of cycle b is channel c = t / (L/P), patch p = t % (L/P):
//       components x[b][p*P + k][c], k = 0..P-1
//   w   (H, P) fp32 row-major (Conv1d weight[:, 0, :]), bias (H)
//   h   (n_cycles * T, H) fp32, T = C * L / P;   a = bf16(gelu(h)) (same shape, may be NULL)
// CTA: 256 threads, 32 tokens x H = 512 outputs per step, W^T and the bias resident in shared memory; thread
// (tx = tid % 64, ty = tid / 64) owns tokens 8 ty .. 8 ty + 7 and columns 4 tx .. 4 tx + 3, 256 + 4 tx .. + 3.
// ---------------------------------------------------------------------------------------
namespace pe {
constexpr int TOK = 32, H = 512, PMAX = 64;
}

// PT: patch size fixed at compile time (the repo's 25: the k loop unrolls and the operand loads of the next steps are
// issued ahead of the FMAs), 0 = run-time P.  The next tile's samples are fetched into registers before the FMA loop and
// parked in the other half of a double-buffered tile afterwards: one __syncthreads per tile, global latency hidden.
template <int PT>
__global__ void __launch_bounds__(256) patch_embed_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                          const float *__restrict__ bias, float *__restrict__ h,
                                                          __nv_bfloat16 *__restrict__ a, int64_t n_tokens, int L, int C, int p_rt)
{
    using namespace pe;
    const int P = PT ? PT : p_rt;
    extern __shared__ __align__(16) float pe_smem[];
    float *wt = pe_smem;                       // [P][H]   W transposed
    float *bs = wt + P * H;                    // [H]
    float *xt = bs + H;                        // 2 x [P][TOK] the tile's patches, component-major (double-buffered)
    const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
    for (int i = tid; i < H * P; i += 256) {
        const int col = i / P, k = i - col * P;
        wt[k * H + col] = __ldg(w + i);
    }
    for (int i = tid; i < H; i += 256)
        bs[i] = __ldg(bias + i);
    const int ppc = L / P;                     // patches per channel
    const int T = ppc * C;                     // tokens per cycle
    const int64_t n_tiles = (n_tokens + TOK - 1) / TOK;
    constexpr int XPT = (TOK * PMAX + 255) / 256;            // samples a thread fetches per tile, at most
    float xr[XPT];
    auto fetch = [&](int64_t tile) {           // this thread's samples of `tile` (zeros beyond the tensor)
        const int64_t t0 = tile * TOK;
#pragma unroll
        for (int q = 0; q < XPT; ++q) {
            const int i = tid + 256 * q;
            float v = 0.0f;
            if (i < TOK * P) {
                const int tok = i / P, k = i - tok * P;
                const int64_t t = t0 + tok;
                if (t < n_tokens) {
                    const int64_t b = t / T;
                    const int ct = (int)(t - b * T), c = ct / ppc, pp = ct - c * ppc;
                    v = __ldg(x + (b * L + (int64_t)pp * P + k) * C + c);
                }
            }
            xr[q] = v;
        }
    };
    auto park = [&](float *dst) {
#pragma unroll
        for (int q = 0; q < XPT; ++q) {
            const int i = tid + 256 * q;
            if (i < TOK * P) {
                const int tok = i / P, k = i - tok * P;
                dst[k * TOK + tok] = xr[q];
            }
        }
    };
    int buf = 0;
    if ((int64_t)blockIdx.x < n_tiles) {
        fetch(blockIdx.x);
        park(xt);
    }
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * TOK;
        const float *xc = xt + buf * (PMAX * TOK);
        const bool more = tile + gridDim.x < n_tiles;
        if (more)
            fetch(tile + gridDim.x);           // in flight during the FMA loop
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j)
                acc[i][j] = 0.0f;
#pragma unroll
        for (int k = 0; k < (PT ? PT : 1); ++k) {
            for (int kk = PT ? k : 0; kk < (PT ? k + 1 : P); ++kk) {
                const float4 x0 = *reinterpret_cast<const float4 *>(xc + kk * TOK + 8 * ty);
                const float4 x1 = *reinterpret_cast<const float4 *>(xc + kk * TOK + 8 * ty + 4);
                const float4 w0 = *reinterpret_cast<const float4 *>(wt + kk * H + 4 * tx);
                const float4 w1 = *reinterpret_cast<const float4 *>(wt + kk * H + 256 + 4 * tx);
                const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const float ws[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        acc[i][j] = fmaf(xs[i], ws[j], acc[i][j]);
            }
        }
        if (more)
            park(xt + (buf ^ 1) * (PMAX * TOK));
        const float4 b0 = *reinterpret_cast<const float4 *>(bs + 4 * tx);
        const float4 b1 = *reinterpret_cast<const float4 *>(bs + 256 + 4 * tx);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t t = t0 + 8 * ty + i;
            if (t >= n_tokens)
                break;
            const float4 o0 = make_float4(acc[i][0] + b0.x, acc[i][1] + b0.y, acc[i][2] + b0.z, acc[i][3] + b0.w);
            const float4 o1 = make_float4(acc[i][4] + b1.x, acc[i][5] + b1.y, acc[i][6] + b1.z, acc[i][7] + b1.w);
            __stcs(reinterpret_cast<float4 *>(h + t * H + 4 * tx), o0);
            __stcs(reinterpret_cast<float4 *>(h + t * H + 256 + 4 * tx), o1);
            if (a) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(tl::gelu_fast(o0.x), tl::gelu_fast(o0.y));
                const __nv_bfloat162 p1 = __floats2bfloat162_rn(tl::gelu_fast(o0.z), tl::gelu_fast(o0.w));
                const __nv_bfloat162 p2 = __floats2bfloat162_rn(tl::gelu_fast(o1.x), tl::gelu_fast(o1.y));
                const __nv_bfloat162 p3 = __floats2bfloat162_rn(tl::gelu_fast(o1.z), tl::gelu_fast(o1.w));
                *reinterpret_cast<uint2 *>(a + t * H + 4 * tx) =
                    make_uint2(*reinterpret_cast<const uint32_t *>(&p0), *reinterpret_cast<const uint32_t *>(&p1));
                *reinterpret_cast<uint2 *>(a + t * H + 256 + 4 * tx) =
                    make_uint2(*reinterpret_cast<const uint32_t *>(&p2), *reinterpret_cast<const uint32_t *>(&p3));
            }
        }
        __syncthreads();                       // the other half is complete, this one may be overwritten next time round
        buf ^= 1;
    }
}

bool patch_embed_supported(int L, int C, int P, int H)
{
    return H == pe::H && P >= 1 && P <= pe::PMAX && C >= 1 && L >= P && L % P == 0;
}

cudaError_t launch_patch_embed(const float *x, const float *w, const float *bias, float *h, void *a, int64_t n_cycles,
                               int L, int C, int P, int H, int sm_count, int max_smem, cudaStream_t st)
{
    if (!patch_embed_supported(L, C, P, H))
        return cudaErrorNotSupported;
    const int64_t n_tokens = n_cycles * (L / P) * C;
    if (n_tokens == 0)
        return cudaSuccess;
    const int smem = (int)sizeof(float) * (P * pe::H + pe::H + 2 * pe::PMAX * pe::TOK);
    if (smem > max_smem)
        return cudaErrorNotSupported;
    auto kern = P == 25 ? patch_embed_kernel<25> : patch_embed_kernel<0>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess)
        return err;
    const int64_t tiles = (n_tokens + pe::TOK - 1) / pe::TOK;
    const int per_sm = max_smem / (smem + 1024) < 3 ? (max_smem / (smem + 1024) < 1 ? 1 : max_smem / (smem + 1024)) : 3;
    const int grid = (int)(tiles < (int64_t)sm_count * per_sm ? tiles : (int64_t)sm_count * per_sm);
    kern<<<grid, 256, smem, st>>>(x, w, bias, h, (__nv_bfloat16 *)a, n_tokens, L, C, P);
    return cudaGetLastError();
}

namespace {

// (rows, cols) bf16 row-major tensor under boxes of `box_rows` x 64 columns (128 bytes, SWIZZLE_128B)
bool make_bf16_map(CUtensorMap *map, const void *base, int64_t rows, int cols, int box_rows)
{
    return tc::make_tensor_map_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, cols, box_rows, tl::BK,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
}

}  // namespace

// which launches take the 16-warp epilogue by default (measured)
static bool tok_linear_default_ew16(int mode, int split, int taps)
{
    // T = 2^20 tokens, 512 x 512: bf16 form 0.614 / 1.291 ms (modes 0 / 1) against 0.660 / 1.389 with 8 warps; split form 1.547 /
    // 2.251 against 1.528 / 2.049 -- its mode 0 is MMA-bound either way, and its mode 1 loses more to the single residual buffer
    // (no prefetch across slabs, 104 bytes of spills at 96 registers) than it gains from the extra warps.  (Rejected for mode 1 of both
    // forms: a cp.async.bulk.prefetch.L2 of the next item's residual slab one tile ahead -- 1.378 / 2.201 ms instead of 1.291 / 2.049.)
    (void)mode; (void)taps;
    return !split;
}

bool tok_linear_supported(int K, int N) { return K >= tl::BK && K % tl::BK == 0 && N >= tl::BN && N % tl::BN == 0; }

cudaError_t launch_tok_linear(const void *a, const void *w, const float *bias, float *h, void *out, int64_t n_tokens,
                              int K, int N, int mode, int sm_count, int max_smem, cudaStream_t st, int taps, int cyc_len,
                              int out_gelu, int split)
{
    using namespace tl;
    // 16 epilogue warps where measured faster (tools/tok_linear_time.py); VQB_TOK_EW=8 / 16 forces one for A/B runs
    static const int ew_env = [] { const char *e = getenv("VQB_TOK_EW"); return e ? atoi(e) : 0; }();
    const bool ew16 = ew_env == 16 || (ew_env != 8 && tok_linear_default_ew16(mode, split, taps));
    const int smem_bytes = ew16 ? (mode == 0 ? Plan<0, 16>::SMEM_BYTES : Plan<1, 16>::SMEM_BYTES)
                                : (mode == 0 ? Plan<0>::SMEM_BYTES : Plan<1>::SMEM_BYTES);
    if (!tok_linear_supported(K, N) || smem_bytes > max_smem || (mode != 0 && !h) || (mode == 0 && !out))
        return cudaErrorNotSupported;
    // three taps: whole cycles of cyc_len tokens, a 128-token tile holds whole cycles
    if (taps != 1 && (taps != 3 || cyc_len < 1 || BM % cyc_len != 0 || n_tokens % cyc_len != 0 || cyc_len > 256))
        return cudaErrorNotSupported;
    if (n_tokens == 0)
        return cudaSuccess;
    CUtensorMap map_a, map_w;
    if (taps == 1) {
        if (!make_bf16_map(&map_a, a, n_tokens, split ? 2 * K : K, BM))
            return cudaErrorNotSupported;
    } else if (!tc::make_tensor_map_3d(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, a, n_tokens / cyc_len, cyc_len, split ? 2 * K : K, BM / cyc_len,
                                       cyc_len, BK, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) {
        return cudaErrorNotSupported;
    }
    if (!make_bf16_map(&map_w, w, N, (split ? 2 : 1) * taps * K, BN))
        return cudaErrorNotSupported;
    const int64_t items = (n_tokens + BM - 1) / BM * (N / BN);
    const int grid = (int)(items < sm_count ? items : sm_count);
    using KernT = void (*)(const CUtensorMap, const CUtensorMap, const float *, float *, __nv_bfloat16 *, int64_t, int, int, int, int, int);
    static const KernT table[2][2][3] = {
        {{tok_linear_kernel<0, false, 8>, tok_linear_kernel<1, false, 8>, tok_linear_kernel<2, false, 8>},
         {tok_linear_kernel<0, true, 8>, tok_linear_kernel<1, true, 8>, tok_linear_kernel<2, true, 8>}},
        {{tok_linear_kernel<0, false, 16>, tok_linear_kernel<1, false, 16>, tok_linear_kernel<2, false, 16>},
         {tok_linear_kernel<0, true, 16>, tok_linear_kernel<1, true, 16>, tok_linear_kernel<2, true, 16>}}};
    KernT kern = table[ew16 ? 1 : 0][split ? 1 : 0][mode];
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (err != cudaSuccess)
        return err;
    kern<<<grid, 128 + 32 * (ew16 ? 16 : 8), smem_bytes, st>>>(map_a, map_w, bias, h, (__nv_bfloat16 *)out, n_tokens, (split ? 3 : 1) * taps * K, N, taps,
                                            cyc_len, out_gelu);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Last layer of the decoder: ConvTranspose1d(H, 1, kernel = stride = P) (model/vq_vae_patch_embedd.py:24-29,52-56) --
// every row of a (rows, H) bf16 activation gives P fp32 samples: out[r][j] = sum_c a[r][c] w[j][c] + bias.
// One warp per row (a row is 2 H bytes, read once), fp32 accumulation.  HBM-bound: 2 H bytes in, 4 P out per row.
// ---------------------------------------------------------------------------------------
constexpr int kOutProjMaxP = 8;
// NS = H / 256 steps of 8 columns per lane; the lane's P x 8 NS weights live in registers (the first version read them from
// shared memory with scalar loads: 8.2 ms per 5.2 M rows, LDS-bound; this one is bound by the row reads)
template <int NS, int P>
__global__ void __launch_bounds__(256) tok_out_proj_kernel(const __nv_bfloat16 *__restrict__ a, const float *__restrict__ w,
                                                           float bias, float *__restrict__ out, int64_t n_rows, int group,
                                                           int64_t ld_group, int accumulate)
{
    constexpr int H = NS * 256;
    const int lane = threadIdx.x & 31;
    float wr[NS][P][8];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int j = 0; j < P; ++j)
#pragma unroll
            for (int t = 0; t < 8; ++t)
                wr[s][j][t] = __ldg(w + j * H + s * 256 + lane * 8 + t);
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * 8) {
        uint4 v[NS];
#pragma unroll
        for (int s = 0; s < NS; ++s)        // 16 bytes per lane and step: a warp reads 512 contiguous bytes
            v[s] = __ldcs(reinterpret_cast<const uint4 *>(a + (r / group) * ld_group + (r % group) * H + s * 256 + lane * 8));
        float acc[P];
#pragma unroll
        for (int j = 0; j < P; ++j)
            acc[j] = 0.0f;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const uint32_t vv[4] = {v[s].x, v[s].y, v[s].z, v[s].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float x0 = __uint_as_float(vv[t] << 16), x1 = __uint_as_float(vv[t] & 0xffff0000u);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    acc[j] = fmaf(x0, wr[s][j][2 * t], acc[j]);
                    acc[j] = fmaf(x1, wr[s][j][2 * t + 1], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < P; ++j) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
        }
        if (lane < P) {
            float res = acc[0];
#pragma unroll
            for (int j = 1; j < P; ++j)
                res = lane == j ? acc[j] : res;
            out[r * P + lane] = res + bias + (accumulate ? out[r * P + lane] : 0.0f);
        }
    }
}

bool tok_out_proj_supported(int H, int P) { return (H == 256 || H == 512) && P >= 1 && P <= kOutProjMaxP; }

// group / ld_group: row r is the (r % group)-th run of H values of the (r / group)-th input row of ld_group elements (group = 1,
// ld_group = H: plain rows); accumulate: out += instead of out = (the lo half of a bf16 pair after its hi half)
cudaError_t launch_tok_out_proj(const void *a, const float *w, float bias, float *out, int64_t n_rows, int H, int P, int sm_count,
                                cudaStream_t st, int group, int64_t ld_group, int accumulate)
{
    if (!tok_out_proj_supported(H, P))
        return cudaErrorNotSupported;
    if (n_rows == 0)
        return cudaSuccess;
    if (group < 1)
        group = 1;
    if (ld_group <= 0)
        ld_group = (int64_t)group * H;
    const int64_t blocks = (n_rows + 7) / 8;
    const int grid = (int)(blocks < (int64_t)sm_count * 8 ? blocks : (int64_t)sm_count * 8);
    const __nv_bfloat16 *ab = (const __nv_bfloat16 *)a;
#define VQB_OUT_PROJ(PV)                                                                                     \
    case PV:                                                                                                 \
        if (H == 256) tok_out_proj_kernel<1, PV><<<grid, 256, 0, st>>>(ab, w, bias, out, n_rows, group, ld_group, accumulate); \
        else tok_out_proj_kernel<2, PV><<<grid, 256, 0, st>>>(ab, w, bias, out, n_rows, group, ld_group, accumulate);          \
        break;
    switch (P) {
        VQB_OUT_PROJ(1) VQB_OUT_PROJ(2) VQB_OUT_PROJ(3) VQB_OUT_PROJ(4) VQB_OUT_PROJ(5) VQB_OUT_PROJ(6) VQB_OUT_PROJ(7)
        VQB_OUT_PROJ(8)
    }
#undef VQB_OUT_PROJ
    return cudaGetLastError();
}

}  // namespace vqb
