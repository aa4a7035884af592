// vq_fwd_tc.cu -- tcgen05 "filter, then decide exactly" forward kernel (sm_100a).
//
// Replaces model/vector_quantizer.py:88-119 for contiguous (N, 32) inputs and K <= 256.
//
// Why tensor cores at all: K=256, D=32 needs 16 384 flop per 264 bytes; on the fp32 FMA
// pipe that is 18 % of the HBM roofline at best (measured: 8.8 %, profiles/README.md).
// Why not tensor cores alone: TF32/BF16 products flip indices.  So the
// tensor cores only FILTER:
//
//   1. split  z = z1 + z2 (+r), E = E1 + E2 (+r) into bf16 pairs (|r| <= 2^-18 |x|) and let
//      tcgen05.mma accumulate   s~[i][k] = ee_k - 2*(z1.E1 + z1.E2 + z2.E1)   in TMEM
//      (7 MMAs of 128 x 256 x 16 per 128-vector tile; ee_k enters through a 7th K-slice of
//      ones times its exact three-way bf16 split, E is pre-scaled by -2);
//   2. per vector, the 256 approximate scores are reduced along TWO orthogonal groupings:
//      A-groups of 16 consecutive codes {16a .. 16a+15} and B-groups of 16 strided codes
//      {b, b+16, b+32, ...}; each grouping keeps a packed (value | group id) top-2 (FMNMX3
//      trees: 1.5 ALU ops per score in total).  Two codes that share an A-group never share
//      a B-group, so min(runner-up A, runner-up B) bounds the second-smallest SCORE;
//   3. |s~ - s| <= eps  =>  if that runner-up is more than delta = 2*eps + rounding slack
//      above the best score, the oracle's argmin is the code (best A-group, best B-group)
//      -- decided by the tensor cores alone (~99.8 % of vectors);
//   4. otherwise (~0.2 % of vectors: near-ties) the vector's own thread evaluates, with the oracle-order
//      expression (ascending fmaf chain, fl(fl(zz+ee) - 2dot), lowest index wins), only the codes the
//      filter could not rule out: the cross product of the A-groups and B-groups whose minimum lies within
//      delta of the best score (2-4 codes as a rule); non-finite vectors / codebooks, NaN distances and
//      crowded candidate sets fall back to a warp-wide exact scan of all K codes (first NaN wins).
//
// The decision is therefore bit-identical to the FMA kernel and the CPU oracle, and everything -- filter,
// exact decision, gather, loss, histogram, z_q -- happens in this one kernel.
//
// Pipeline per CTA (persistent, one CTA per SM, 768 threads):
//   warp 3      ring owner: TMA loads of z tiles (128 x 32 fp32, SWIZZLE_128B) into a 6-deep ring,
//               z_q TMA store out of the same slot, refill of the slot it just released
//   warp 1      MMA issuer (one thread): 7 x tcgen05.mma.kind::f16 -> TMEM (2 x 256 columns)
//   warp 2      TMEM allocator
//   warps 4-7   converters: fp32 tile -> bf16 [z1|z2] tile in the UMMA K-major SW128 layout (Veltkamp split
//               on the FMA pipe), ||z||^2 bound on the side
//   warps 8-23  four epilogue groups (tile i -> group i % 4, TMEM buffer i & 1): TMEM -> A/B group minima ->
//               decision -> z_q written in place into the ring slot -> TMA store; idx, loss, histogram
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "vq_common.cuh"
#include "vq_ptx.cuh"

namespace vqb {

namespace tc {

constexpr int D = 32;
constexpr int KMAX = 256;
constexpr int STAGES = 6;                 // (7 stages fit but were measured slower: no L1 left beside 227 KB of smem)
#ifndef VQB_GROUPS
#define VQB_GROUPS 3
#endif
constexpr int GROUPS = VQB_GROUPS;        // epilogue groups (4 warps each) rotating over the 2 TMEM buffers
constexpr int THREADS = 256 + 128 * GROUPS;
constexpr int MAX_CAND = 32;              // candidate codes a queued vector may have (one lane of the fix-up warp each)
// Role -> warp map.  The SM's warp arbiter prefers the highest warp id among eligible warps (B300_MICROARCH.md,
// "Multi-warp arbiter"), so the single-threaded roles on the accumulator's critical path (MMA issuer, ring owner)
// sit in the last warpgroup, the converters below them and the 16 epilogue warps at the bottom.
#ifndef VQB_ROLEMAP
#define VQB_ROLEMAP 1
#endif
constexpr int W_EPI = VQB_ROLEMAP ? 0 : 8;                   // 4 * GROUPS epilogue warps
constexpr int W_CONV = VQB_ROLEMAP ? 4 * GROUPS : 4;         // 4 converter warps
constexpr int W_SVC = VQB_ROLEMAP ? 4 * GROUPS + 4 : 0;      // service warpgroup: +0 barrier init, +1 MMA issuer, +2 TMEM allocator, +3 ring owner
#ifndef VQB_LDPIPE
#define VQB_LDPIPE 1
#endif

// Register pool of the CTA = THREADS x (registers at launch); setmaxnreg moves it between the warpgroups:
// 4 groups: 768 x 80 = 61440 = 128 x (40 + 56 + 4 x 96);  3 groups: 640 x 96 = 61440 = 128 x (40 + 56 + 3 x 128).
// (An increase beyond what the other warpgroups released blocks forever.)
#ifndef VQB_EPI_REGS
#define VQB_EPI_REGS (VQB_GROUPS == 3 ? 128 : 96)
#endif
static_assert(128 * (40 + 56 + VQB_GROUPS * VQB_EPI_REGS) <= (256 + 128 * VQB_GROUPS) * (65536 / (256 + 128 * VQB_GROUPS) / 8 * 8),
              "setmaxnreg budget exceeds the CTA's register pool");

// shared-memory map (bytes); SW128 operands need 1024-byte alignment
constexpr int OFF_ZRING = 0;                              // STAGES x 16384  fp32 z tiles (TMA, SW128)
constexpr int OFF_ARING = OFF_ZRING + STAGES * 16384;     // 2 x 16384       bf16 [z1|z2] tiles (SW128)
constexpr int OFF_BMAIN = OFF_ARING + 2 * 16384;          // 32768           bf16 -2*[E1|E2] (SW128)
constexpr int OFF_BAUG = OFF_BMAIN + 32768;               // 8192            bf16 ee split (SW32)
constexpr int OFF_AAUG = OFF_BAUG + 8192;                 // 4096            bf16 ones (SW32)
constexpr int OFF_EF32 = OFF_AAUG + 4096;                 // 32768           fp32 codebook, XOR-swizzled rows
constexpr int OFF_EE = OFF_EF32 + 32768;                  // 1024            fp32 ||E_k||^2 (oracle order)
constexpr int OFF_HIST = OFF_EE + 1024;                   // 1024            u32 histogram
constexpr int OFF_ZZ = OFF_HIST + 1024;                   // STAGES x 512    ||z||^2 bound per row (from the converters)
constexpr int OFF_BARS = OFF_ZZ + STAGES * 512;           // 512             mbarriers + tmem base
constexpr int SMEM_BYTES = OFF_BARS + 512;
constexpr int SMEM_ALLOC = SMEM_BYTES;
constexpr int SMEM_ALLOC_TF32 = SMEM_BYTES + (STAGES + 2) * TILE_M * 8 + 64;   // + the refiner lists (vq_fwd_tc_kernel, TF32)

// image of the constant operands prepared once per call in global scratch
constexpr int IMG_BMAIN = 0;
constexpr int IMG_BAUG = IMG_BMAIN + 32768;
constexpr int IMG_EF32 = IMG_BAUG + 8192;
constexpr int IMG_EE = IMG_EF32 + 32768;
constexpr int IMG_CONST = IMG_EE + 1024;                  // Consts
constexpr int IMG_WLCOUNT = IMG_CONST + 64;               // u32 [WL_CTAS]  vectors queued per CTA
constexpr int WL_CTAS = 192;                              // >= number of CTAs (one per SM)
constexpr int WL_CAP = 2048;                              // vectors a CTA can queue for the fix-up kernel
constexpr int IMG_WL = IMG_WLCOUNT + WL_CTAS * 4;         // uint2 [WL_CTAS][WL_CAP]: row, A-mask | B-mask << 16
constexpr int IMG_BYTES = IMG_WL + WL_CTAS * WL_CAP * 8;
#ifndef VQB_FIX_SPLIT
#define VQB_FIX_SPLIT 8
#endif
constexpr int FIX_SPLIT = VQB_FIX_SPLIT;                  // fix-up CTAs per queue (8 x 148 CTAs of 256 threads = one resident wave)

struct Consts {
    unsigned emax2_bits;   // max_k ee_k (finite ones), as float bits
    unsigned nonfinite;    // some ee_k is not finite
    unsigned pad[14];
};

// byte offset of 16-byte chunk `c` of row `r` in a 128-byte-row SW128 tile
__device__ __forceinline__ int sw128(int r, int c) { return r * 128 + ((c ^ (r & 7)) << 4); }
// fp32 codebook rows: chunk swizzle that also separates rows 8 apart (best-chunk gathers)
__device__ __forceinline__ int ef32_off(int k, int c) { return k * 128 + ((c ^ ((k ^ (k >> 3)) & 7)) << 4); }

}  // namespace tc

// ---------------------------------------------------------------------------------------
// prep: one thread per (padded) code builds the constant operand image in global scratch
// ---------------------------------------------------------------------------------------
// tf32: the operand image of the TF32 variant of the kernel (same regions, same sizes): BMAIN row k = -2 * tf32(E_k) as
// 32 fp32 words, BAUG row k = ee_k as an exact three-way tf32 split (8 fp32 words)
__global__ void vq_tc_prep_kernel(const float *__restrict__ E, const float *__restrict__ ee, int K, int d, int kp,
                                  unsigned char *__restrict__ img, int tf32)
{
    using namespace tc;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= KMAX)
        return;
    Consts *cst = reinterpret_cast<Consts *>(img + IMG_CONST);
    const bool real = k < K;
    // components [32*dc, 32*dc + 32) per D-chunk dc; columns beyond d are zero (they change no sum)
    float e[D];
    for (int dc = 0; dc < (d > D ? 2 : 1); ++dc) {
#pragma unroll
        for (int j = 0; j < D; ++j)
            e[j] = (real && dc * D + j < d) ? __ldg(E + (size_t)k * d + dc * D + j) : 0.0f;
        // main B operand: row k = [-2*E1 (32 bf16) | -2*E2 (32 bf16)], SW128; the second D-chunk of a wide
        // codebook (d > 32) takes the place of the fp32 copy, which wide codebooks read from global memory
        unsigned char *bm = img + (dc == 0 ? IMG_BMAIN : IMG_EF32);
        if (tf32) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4 *>(bm + sw128(k, c)) =
                    make_float4(-2.0f * tc::round_tf32(e[4 * c]), -2.0f * tc::round_tf32(e[4 * c + 1]),
                                -2.0f * tc::round_tf32(e[4 * c + 2]), -2.0f * tc::round_tf32(e[4 * c + 3]));
            continue;
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            __nv_bfloat16 out[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int j = (c & 3) * 8 + t;
                const __nv_bfloat16 hi = __float2bfloat16_rn(e[j]);
                const __nv_bfloat16 lo = __float2bfloat16_rn(e[j] - __bfloat162float(hi));
                out[t] = __float2bfloat16_rn(-2.0f * __bfloat162float(c < 4 ? hi : lo));   // exact scaling
            }
            *reinterpret_cast<uint4 *>(bm + sw128(k, c)) = *reinterpret_cast<uint4 *>(out);
        }
    }
    // augmentation B operand: ee_k = a1 + a2 + a3 exactly (3 x 8 bits); pads get a huge score
    const float eek = real ? ee[k] : 3.0e38f;
    const bool fin = isfinite(eek);
    const float eef = fin ? eek : 3.0e38f;
    const __nv_bfloat16 a1 = __float2bfloat16_rn(eef);
    const float r1 = eef - __bfloat162float(a1);
    const __nv_bfloat16 a2 = __float2bfloat16_rn(r1);
    const __nv_bfloat16 a3 = __float2bfloat16_rn(r1 - __bfloat162float(a2));
    if (tf32) {
        const float t1 = tc::round_tf32(eef), t2 = tc::round_tf32(eef - t1), t3 = tc::round_tf32((eef - t1) - t2);
        const int sw = (k >> 2) & 1;   // SW32: 16-byte chunk index ^= bit 7 of the byte offset
        *reinterpret_cast<float4 *>(img + IMG_BAUG + k * 32 + ((0 ^ sw) << 4)) = make_float4(t1, t2, t3, 0.0f);
        *reinterpret_cast<float4 *>(img + IMG_BAUG + k * 32 + ((1 ^ sw) << 4)) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        __nv_bfloat16 out[8] = {a1, a2, a3, __float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f),
                                __float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f), __float2bfloat16_rn(0.f)};
        const int sw = (k >> 2) & 1;   // SW32: 16-byte chunk index ^= bit 7 of the byte offset
        *reinterpret_cast<uint4 *>(img + IMG_BAUG + k * 32 + ((0 ^ sw) << 4)) = *reinterpret_cast<uint4 *>(out);
        *reinterpret_cast<uint4 *>(img + IMG_BAUG + k * 32 + ((1 ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
    }
    // fp32 codebook for the exact decision + gather (d <= 32)
    if (d <= D) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4 *>(img + IMG_EF32 + ef32_off(k, c)) =
                make_float4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
    }
    reinterpret_cast<float *>(img + IMG_EE)[k] = real ? ee[k] : __int_as_float(0x7f800000);
    if (real) {
        if (fin)
            atomicMax(&cst->emax2_bits, __float_as_uint(eek));   // non-negative floats order like uints
        else
            atomicOr(&cst->nonfinite, 1u);
    }
    (void)kp;
}

// ---------------------------------------------------------------------------------------
// exact pieces shared by the epilogue
// ---------------------------------------------------------------------------------------
namespace tc {

// Warp-cooperative exact scan of all K codes for one vector (row `row_in_tile` of the ring
// slot `ztile`, SW128 layout): the oracle-order expression, lowest index on ties, first NaN
// wins.  Every lane returns the same code.
__device__ __noinline__ int warp_full_scan(const unsigned char *ztile, int row_in_tile,
                                           const unsigned char *ef32, const float *ees, int K)
{
    const int lane = threadIdx.x & 31;
    float z[D];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(ztile + sw128(row_in_tile, c));
        z[4 * c] = v.x; z[4 * c + 1] = v.y; z[4 * c + 2] = v.z; z[4 * c + 3] = v.w;
    }
    float zz = 0.0f;
#pragma unroll
    for (int j = 0; j < D; ++j)
        zz = fmaf(z[j], z[j], zz);          // oracle-order ||z||^2
    float best = __int_as_float(0x7f800000);
    int bidx = 0x7fffffff;
    unsigned first_nan = 0xffffffffu;
    for (int k = lane; k < K; k += 32) {
        float acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 e = *reinterpret_cast<const float4 *>(ef32 + ef32_off(k, c));
            acc = fmaf(z[4 * c], e.x, acc);
            acc = fmaf(z[4 * c + 1], e.y, acc);
            acc = fmaf(z[4 * c + 2], e.z, acc);
            acc = fmaf(z[4 * c + 3], e.w, acc);
        }
        const float dist = ref_distance(zz, ees[k], acc);
        if (dist != dist)
            first_nan = min(first_nan, (unsigned)k);
        if (dist < best) {                  // k ascends per lane: strict < keeps the lowest index
            best = dist;
            bidx = k;
        }
    }
    const unsigned nan_k = __reduce_min_sync(0xffffffffu, first_nan);
    if (nan_k != 0xffffffffu)
        return (int)nan_k;                  // torch.argmin: the first NaN wins
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob < best || (ob == best && oi < bidx)) {
            best = ob;
            bidx = oi;
        }
    }
    return bidx == 0x7fffffff ? 0 : bidx;
}

// minimum of 16 registers with 3-input FMNMX: 8 instructions
__device__ __forceinline__ float min16(const uint32_t *v)
{
    const float t0 = min3(__uint_as_float(v[0]), __uint_as_float(v[1]), __uint_as_float(v[2]));
    const float t1 = min3(__uint_as_float(v[3]), __uint_as_float(v[4]), __uint_as_float(v[5]));
    const float t2 = min3(__uint_as_float(v[6]), __uint_as_float(v[7]), __uint_as_float(v[8]));
    const float t3 = min3(__uint_as_float(v[9]), __uint_as_float(v[10]), __uint_as_float(v[11]));
    const float t4 = min3(__uint_as_float(v[12]), __uint_as_float(v[13]), __uint_as_float(v[14]));
    const float t5 = min3(__uint_as_float(v[15]), t0, t1);
    const float t6 = min3(t2, t3, t4);
    return fminf(t5, t6);
}

__device__ __forceinline__ float min16f(const float *v)
{
    const float t0 = min3(v[0], v[1], v[2]), t1 = min3(v[3], v[4], v[5]), t2 = min3(v[6], v[7], v[8]);
    const float t3 = min3(v[9], v[10], v[11]), t4 = min3(v[12], v[13], v[14]);
    return fminf(min3(v[15], t0, t1), min3(t2, t3, t4));
}

// z_q row = z + (e - z) written in place over the z row (ring slot, SW128), returns the row's sum of
// squared residuals.  POISON: gather-by-GEMM semantics for a non-finite codebook (oracle column_poison);
// kept out of the common instantiation so that its loads and branches do not split the hot loop.
template <bool POISON>
__device__ __forceinline__ float emit_row(unsigned char *zrow, int x, const unsigned char *ef32, int code,
                                          bool write_zq, const int *colcnt, const int *colwhich)
{
    // (Walking the row in physical chunk order would save the swizzle XORs, but then the 8 lanes of a
    // quarter-warp hit the same bank group: measured 2x slower.  Logical order is conflict-free.)
    const unsigned char *erow = ef32 + code * 128;
    const int xe = ((code ^ (code >> 3)) & 7) << 4;
    float2 rs01 = make_float2(0.f, 0.f), rs23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 *zp4 = reinterpret_cast<float4 *>(zrow + ((c << 4) ^ x));
        const float4 zv = *zp4;
        float4 e = *reinterpret_cast<const float4 *>(erow + ((c << 4) ^ xe));
        if (POISON) {
            float *ev = reinterpret_cast<float *>(&e);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = 4 * c + t, cc = colcnt[j];
                if (!(cc == 0 || (cc == 1 && colwhich[j] == code + 1)))
                    ev[t] = __int_as_float(0x7fc00000);
            }
        }
        // packed fp32 pairs (FADD2 / FFMA2, IEEE per lane: the same values as the scalar form, half the instructions --
        // the kernel's time tracks the energy of what it executes, profiles/README.md round 2)
        const float2 d01 = __fadd2_rn(make_float2(e.x, e.y), make_float2(-zv.x, -zv.y));     // fl(e - z)
        const float2 d23 = __fadd2_rn(make_float2(e.z, e.w), make_float2(-zv.z, -zv.w));
        rs01 = __ffma2_rn(d01, d01, rs01);
        rs23 = __ffma2_rn(d23, d23, rs23);
        if (write_zq) {
            const float2 o01 = __fadd2_rn(make_float2(zv.x, zv.y), d01);                     // fl(z + fl(e - z))
            const float2 o23 = __fadd2_rn(make_float2(zv.z, zv.w), d23);
            *zp4 = make_float4(o01.x, o01.y, o23.x, o23.y);
        }
    }
    return (rs01.x + rs01.y) + (rs23.x + rs23.y);
}


// The same for the 32 rows of a warp (rows 32 q .. 32 q + 31 of the tile), EIGHT LANES PER ROW: lane = (row % 4, chunk) walks
// rows 4 j + lane / 8, so that every quarter-warp reads one whole 128-byte row of z and one whole codebook row per
// access -- conflict-free whatever codes the rows have.  (With one thread per row the codebook gather costs ~10
// shared-memory wavefronts per LDS.128 instead of 4: 32 random rows, profiles/README.md r02b.)  `code_or_neg`: this
// lane's OWN row's code, or -1 when that row is written elsewhere (queued / refined / beyond the tensor); the return
// value is this lane's share of the squared residuals of the rows it walked (their sum over the warp is what counts).
__device__ __forceinline__ float emit_rows_coop(unsigned char *ztile, int q, int lane, const unsigned char *ef32, int code_or_neg,
                                                bool write_zq)
{
    const int c = lane & 7, sub = lane >> 3;
    float2 rs01 = make_float2(0.f, 0.f), rs23 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int rr = 4 * j + sub;
        const int cj = __shfl_sync(0xffffffffu, code_or_neg, rr);
        // (branch-free: a skipped row is read like the others -- against code 0 -- and neither stored nor counted; with a
        // branch per row the eight iterations serialise on their load latencies)
        const bool on = cj >= 0;
        float4 *zp4 = reinterpret_cast<float4 *>(ztile + (q * 32 + rr) * 128 + ((c ^ (rr & 7)) << 4));
        const float4 zv = *zp4;
        const float4 e = *reinterpret_cast<const float4 *>(ef32 + ef32_off(on ? cj : 0, c));
        float2 d01 = __fadd2_rn(make_float2(e.x, e.y), make_float2(-zv.x, -zv.y));               // fl(e - z)
        float2 d23 = __fadd2_rn(make_float2(e.z, e.w), make_float2(-zv.z, -zv.w));
        if (write_zq && on) {
            const float2 o01 = __fadd2_rn(make_float2(zv.x, zv.y), d01);                         // fl(z + fl(e - z))
            const float2 o23 = __fadd2_rn(make_float2(zv.z, zv.w), d23);
            *zp4 = make_float4(o01.x, o01.y, o23.x, o23.y);
        }
        if (!on) {
            d01 = make_float2(0.f, 0.f);
            d23 = make_float2(0.f, 0.f);
        }
        rs01 = __ffma2_rn(d01, d01, rs01);
        rs23 = __ffma2_rn(d23, d23, rs23);
    }
    return (rs01.x + rs01.y) + (rs23.x + rs23.y);
}

// Oracle-order distance of one code for the vector in ring-slot row `zrow` (chunk passes of large codebooks:
// the winners of the 256-code chunks are compared by their exact distances).
__device__ __forceinline__ float exact_distance(const unsigned char *zrow, int x, const unsigned char *ef32,
                                                const float *ees, int code)
{
    float zz = 0.0f, acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>(zrow + ((c << 4) ^ x));
        const float4 e = *reinterpret_cast<const float4 *>(ef32 + ef32_off(code, c));
        zz = fmaf(v.x, v.x, zz); zz = fmaf(v.y, v.y, zz); zz = fmaf(v.z, v.z, zz); zz = fmaf(v.w, v.w, zz);
        acc = fmaf(v.x, e.x, acc); acc = fmaf(v.y, e.y, acc); acc = fmaf(v.z, e.z, acc); acc = fmaf(v.w, e.w, acc);
    }
    return ref_distance(zz, ees[code], acc);
}

// ---- wide vectors (32 < d <= 64): a tile occupies two ring slots (components 0..31 and 32..63, the second one
// zero-filled beyond d) and the fp32 codebook is read from global memory (L2) instead of shared memory ----
__device__ __forceinline__ float4 ld_e4(const float *E, int code, int d, int j)
{   // E[code][j .. j+3], zero beyond the row (d is a multiple of 4)
    return j < d ? __ldg(reinterpret_cast<const float4 *>(E + (size_t)code * d + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__device__ __forceinline__ float exact_distance_wide(const unsigned char *zrow0, const unsigned char *zrow1, int x,
                                                     const float *E, int d, const float *ees, int code)
{
    float zz = 0.0f, acc = 0.0f;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float4 v = *reinterpret_cast<const float4 *>((c < 8 ? zrow0 : zrow1) + (((c & 7) << 4) ^ x));
        const float4 e = ld_e4(E, code, d, 4 * c);
        zz = fmaf(v.x, v.x, zz); zz = fmaf(v.y, v.y, zz); zz = fmaf(v.z, v.z, zz); zz = fmaf(v.w, v.w, zz);
        acc = fmaf(v.x, e.x, acc); acc = fmaf(v.y, e.y, acc); acc = fmaf(v.z, e.z, acc); acc = fmaf(v.w, e.w, acc);
    }
    return ref_distance(zz, ees[code], acc);
}

// warp-wide exact scan of all K codes for one wide vector (rare path: codebook rows come from L2)
__device__ __noinline__ int warp_full_scan_wide(const unsigned char *zt0, const unsigned char *zt1, int row_in_tile,
                                                const float *E, int d, const float *ees, int K)
{
    const int lane = threadIdx.x & 31;
    const int x = (row_in_tile & 7) << 4;
    const unsigned char *zrow0 = zt0 + row_in_tile * 128, *zrow1 = zt1 + row_in_tile * 128;
    float best = __int_as_float(0x7f800000);
    int bidx = 0x7fffffff;
    unsigned first_nan = 0xffffffffu;
    for (int k = lane; k < K; k += 32) {
        const float dist = exact_distance_wide(zrow0, zrow1, x, E, d, ees, k);
        if (dist != dist)
            first_nan = min(first_nan, (unsigned)k);
        if (dist < best) {
            best = dist;
            bidx = k;
        }
    }
    const unsigned nan_k = __reduce_min_sync(0xffffffffu, first_nan);
    if (nan_k != 0xffffffffu)
        return (int)nan_k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ob < best || (ob == best && oi < bidx)) {
            best = ob;
            bidx = oi;
        }
    }
    return bidx == 0x7fffffff ? 0 : bidx;
}

// z_q for one 32-component half of a wide row, codebook row from global memory
template <bool POISON>
__device__ __forceinline__ float emit_row_wide(unsigned char *zrow, int x, const float *E, int d, int j0, int code,
                                               bool write_zq, const int *colcnt, const int *colwhich)
{
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    float4 e[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
        e[c] = ld_e4(E, code, d, j0 + 4 * c);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 *zp4 = reinterpret_cast<float4 *>(zrow + ((c << 4) ^ x));
        const float4 zv = *zp4;
        if (POISON) {
            float *ev = reinterpret_cast<float *>(&e[c]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int j = j0 + 4 * c + t;
                if (j < d) {
                    const int cc = colcnt[j];
                    if (!(cc == 0 || (cc == 1 && colwhich[j] == code + 1)))
                        ev[t] = __int_as_float(0x7fc00000);
                }
            }
        }
        float4 o;
        float dj;
        dj = __fsub_rn(e[c].x, zv.x); rs[0] = fmaf(dj, dj, rs[0]); o.x = __fadd_rn(zv.x, dj);
        dj = __fsub_rn(e[c].y, zv.y); rs[1] = fmaf(dj, dj, rs[1]); o.y = __fadd_rn(zv.y, dj);
        dj = __fsub_rn(e[c].z, zv.z); rs[2] = fmaf(dj, dj, rs[2]); o.z = __fadd_rn(zv.z, dj);
        dj = __fsub_rn(e[c].w, zv.w); rs[3] = fmaf(dj, dj, rs[3]); o.w = __fadd_rn(zv.w, dj);
        if (write_zq)
            *zp4 = o;
    }
    return (rs[0] + rs[1]) + (rs[2] + rs[3]);
}

// Running argmin over the chunks of a large codebook, torch.argmin semantics: the first NaN wins, a tie keeps
// the earlier chunk (= the lower index).
__device__ __forceinline__ void merge_running(unsigned long long *run, int64_t row, float dist, int code, int chunk_mode)
{
    const unsigned long long cur = ((unsigned long long)__float_as_uint(dist) << 32) | (unsigned)code;
    if (chunk_mode == 2) {
        const float dp = __uint_as_float((unsigned)(run[row] >> 32));
        if ((dp != dp) || !((dist != dist) || dist < dp))
            return;
    }
    run[row] = cur;
}

}  // namespace tc

// ---------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------
// TRACE: debug build of the same kernel that records clock64() of eight pipeline events per tile
// (first kTraceTiles tiles of the first kTraceCtas CTAs) -- tools/tc_trace.py turns them into a timeline.
constexpr int kTraceCtas = 4, kTraceTiles = 1024, kTraceEvents = 8;
// WIDE: 32 < D <= 64 (two pipeline items per tile); a separate instantiation keeps the D <= 32 kernel's hot loops free
// of the wide-vector code.
// PAIR: two CTAs of a cluster (one TPC) run every tcgen05.mma together (cta_group::2, M = 256): each CTA converts and
// filters its own 128-vector tile, but holds only HALF of the codebook operand (codes 128 r .. 128 r + 127 for cluster
// rank r), so the tensor core of each SM reads 8 KB instead of 12 KB of shared memory per MMA.  Only the leader issues;
// the peer's converters and epilogue warps arrive on the leader's mbarriers, completions are multicast to both CTAs.
// TF32 (experiment, VQB_TF32=1): ONE tf32 product straight off the TMA-written fp32 tile instead of three bf16 products
// of a converted tile -- 4 MMAs of K = 8 + the norm slice = 5 tensor-pipe slots per tile instead of 7, and no converter
// warps at all.  The price is a filter radius ~32x larger (the tensor core uses 11 significant bits of z and of E), i.e.
// ~8 % of the vectors instead of ~0.25 % leave the filter uncertified; up to four candidate codes are then decided by
// the vector's own thread right away (oracle-order distances against the fp32 codebook in shared memory), larger
// candidate sets take the queue as before.  The decision stays the oracle's in every case.
template <bool TRACE, bool WIDE, bool PAIR, bool TF32>
__global__ void __launch_bounds__(tc::THREADS, 1)
vq_fwd_tc_kernel(const FwdParams p, const unsigned char *__restrict__ img, const __grid_constant__ CUtensorMap map_z,
                 const __grid_constant__ CUtensorMap map_zq, int kp, unsigned long long *trace)
{
    using namespace tc;
    // ring depth: the TF32 variant has no bf16 operand ring, its 32 KB hold two more fp32 slots (the refiner warps hold a
    // slot ~2000 clocks longer than the epilogue alone)
    constexpr int NST = TF32 ? STAGES + 2 : STAGES;
    auto stamp = [&](int i, int ev) {
        if (TRACE && blockIdx.x < kTraceCtas && i < kTraceTiles)
            trace[((size_t)blockIdx.x * kTraceTiles + i) * kTraceEvents + ev] = (unsigned long long)clock64();
    };
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    if ((sbase & 1023u) != 0)
        __trap();                  // SW128 operands and TMA boxes need a 1024-byte aligned base
    // barrier indices
    enum { Z_FULL = 0, Q_DONE = Z_FULL + NST, A_FULL = Q_DONE + NST,
           A_EMPTY = A_FULL + 2, T_FULL = A_EMPTY + 2, T_EMPTY = T_FULL + GROUPS, R_FULL = T_EMPTY + 2,
           N_BARS = R_FULL + NST };
    // R_FULL (TF32): the tile's list of uncertified rows is complete -- the refiner warps may take it
    // T_FULL is per epilogue GROUP (a waiter must see every phase of its barrier), T_EMPTY per TMEM buffer
    static_assert(8 * N_BARS + 8 <= 512, "barrier area");
    auto bar = [&](int i) { return sbase + OFF_BARS + 8 * i; };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + OFF_BARS + 8 * N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long clk_begin = clock64();
    const int64_t n_rows = p.z.n_rows;
    const int64_t n_tiles = (n_rows + TILE_M - 1) / TILE_M;
    // PAIR: cluster c of n_cl works on the tile pairs 2 (c + i n_cl), + 1; its CTA of rank r takes the tile 2 (c + i n_cl) + r
    // (a pair's second tile may lie beyond the tensor: its rows are masked like the rows of a ragged last tile)
    const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
    const uint32_t tile_first = PAIR ? 2u * (blockIdx.x >> 1) + crank : blockIdx.x;
    const uint32_t tile_step = PAIR ? (gridDim.x & ~1u) : gridDim.x;
    const int64_t n_units = PAIR ? (n_tiles + 1) / 2 : n_tiles, unit0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x,
                  n_workers = PAIR ? (gridDim.x >> 1) : gridDim.x;
    const int64_t my_tiles = unit0 < n_units ? (n_units - unit0 + n_workers - 1) / n_workers : 0;
    // wide vectors (32 < D <= 64): every tile is two pipeline items (D-chunks of 32 components), each with its own
    // ring slot and converter pass; their products accumulate in the same TMEM buffer
    constexpr int nd = WIDE ? 2 : 1;
    const int64_t my_items = my_tiles * nd;
    const int K = p.K;

    // ---- one-time setup -----------------------------------------------------------------
    if (warp == W_SVC && lane == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bar(Z_FULL + s), 1);
            mbar_init(bar(Q_DONE + s), TF32 ? 160 : 128);       // TF32: + the lanes of the tile's refiner warp
            mbar_init(bar(R_FULL + s), 128);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(A_FULL + b), PAIR ? 8 : 128);       // PAIR: one arrival per warp, the leader's barriers count both CTAs
            mbar_init(bar(A_EMPTY + b), 1);
            mbar_init(bar(T_EMPTY + b), PAIR ? 8 : 128);
        }
        for (int g = 0; g < GROUPS; ++g)
            mbar_init(bar(T_FULL + g), 1);
        fence_barrier_init();
    }
    if (warp == W_SVC + 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    // constant operands: copy the prepared image, build the ones tile, clear the histogram
    for (int i = tid; i < (IMG_EE + 1024) / 16; i += THREADS) {
        // image order = smem order for BMAIN | BAUG, then EF32 | EE live after the AAUG tile
        const int off = i * 16;
        int src = off;
        if (PAIR && off < IMG_EF32) {
            // a CTA of a pair holds the operand rows of ITS 128 codes at the start of BMAIN / BAUG
            if (off < IMG_BAUG) {
                if (off >= IMG_BAUG / 2) continue;
                src = off + (int)crank * (IMG_BAUG / 2);
            } else {
                if (off - IMG_BAUG >= (IMG_EF32 - IMG_BAUG) / 2) continue;
                src = off + (int)crank * ((IMG_EF32 - IMG_BAUG) / 2);
            }
        }
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(img + src));
        const int dst = off < IMG_EF32 ? OFF_BMAIN + off : OFF_EF32 + (off - IMG_EF32);
        *reinterpret_cast<uint4 *>(smem + dst) = v;
    }
    if (tid < TILE_M) {
        const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f);
        __nv_bfloat16 out[8] = {one, one, one, zero, zero, zero, zero, zero};
        const int sw = (tid >> 2) & 1;
        if (TF32)       // K = 8 tf32 per row: [1, 1, 1, 0 | 0, 0, 0, 0]
            *reinterpret_cast<float4 *>(smem + OFF_AAUG + tid * 32 + ((0 ^ sw) << 4)) = make_float4(1.f, 1.f, 1.f, 0.f);
        else
            *reinterpret_cast<uint4 *>(smem + OFF_AAUG + tid * 32 + ((0 ^ sw) << 4)) = *reinterpret_cast<uint4 *>(out);
        *reinterpret_cast<uint4 *>(smem + OFF_AAUG + tid * 32 + ((1 ^ sw) << 4)) = make_uint4(0, 0, 0, 0);
    }
    if (tid < KMAX)
        reinterpret_cast<unsigned *>(smem + OFF_HIST)[tid] = 0u;
    unsigned *wl_count_s = reinterpret_cast<unsigned *>(smem + OFF_BARS + 8 * N_BARS + 4);
    if (tid == 0)
        *wl_count_s = 0u;
    // TF32: per ring slot, the rows the filter left uncertified with at most four candidates (row in tile, masks)
    uint2 *rlist = reinterpret_cast<uint2 *>(smem + SMEM_BYTES);                          // [NST][TILE_M]
    unsigned *rcnt = reinterpret_cast<unsigned *>(smem + SMEM_BYTES + NST * TILE_M * 8);  // [NST]
    if (TF32 && tid < NST)
        rcnt[tid] = 0u;
    fence_proxy_async();          // generic-proxy writes of the operands -> visible to tcgen05/TMA
    tc_fence_before();
    if (PAIR)
        cluster_sync_all();       // the peer's barriers and operands are ready before anything is signalled across
    else
        __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const Consts *cst = reinterpret_cast<const Consts *>(img + IMG_CONST);
    double sq = 0.0;

    // register budget (setmaxnreg acts on warpgroups): 4 x 40 + 4 x 56 + 16 x 96 (104) <= 24 warps x 85
    if (warp >= W_SVC && warp < W_SVC + 4)
        reg_dec<40>();
    if (warp == W_SVC) {
        // (idle: the ring is refilled by the store warp the moment it has released a slot)
    } else if (warp == W_SVC + 1) {
        // ================= MMA issuer: the whole warp runs the loop, one elected lane issues (vq_ptx.cuh: elect_one) ====
        if (TF32) {
            // the ring slot the TMA wrote IS the A operand (fp32 words, SW128 K-major): no conversion, no operand ring
            const uint32_t idesc = idesc_tf32(kp);
            const uint64_t bmain = desc_sw128(sbase + OFF_BMAIN);
            const uint64_t baug = desc_sw32(sbase + OFF_BAUG);
            const uint64_t aaug = desc_sw32(sbase + OFF_AAUG);
            const uint64_t z0 = desc_sw128(sbase + OFF_ZRING);
            const int n_items = (int)my_items;
            const int n_ks = (p.D + 7) >> 3;                         // K-slices of 8 components that hold data
            int g = 0, s = 0;
            uint32_t zph = 0;
            for (int i = 0; i < n_items; ++i) {
                const int b = i & 1;
                mbar_wait<32>(bar(Z_FULL + s), zph);
                mbar_wait<32>(bar(T_EMPTY + b), (uint32_t)(((i >> 1) & 1) ^ 1));
                tc_fence_after();
                if (lane == 0) stamp(i, 3);
                if (elect_one()) {
                    const uint64_t a = z0 + (uint64_t)(s * (16384 >> 4));
                    const uint32_t d = tmem_base + b * KMAX;
                    umma_tf32(d, a + 0, bmain + 0, idesc, 0);
                    if (n_ks > 1) umma_tf32(d, a + 2, bmain + 2, idesc, 1);
                    if (n_ks > 2) umma_tf32(d, a + 4, bmain + 4, idesc, 1);
                    if (n_ks > 3) umma_tf32(d, a + 6, bmain + 6, idesc, 1);
                    umma_tf32(d, aaug, baug, idesc, 1);                // + ee_k
                    umma_commit(bar(T_FULL + g));
                }
                __syncwarp();
                g = g + 1 == GROUPS ? 0 : g + 1;
                if (++s == NST) {
                    s = 0;
                    zph ^= 1u;
                }
            }
        } else if (!PAIR || crank == 0) {
            // (Issuing the codebook as two N = kp/2 halves with an early commit was measured: the 14
            // half-width MMAs take ~2x the tensor time of 7 full-width ones, a net loss.)
            const uint32_t idesc = PAIR ? idesc_bf16_mn(256, KMAX) : idesc_bf16(kp);
            const uint64_t bmain = desc_sw128(sbase + OFF_BMAIN);
            const uint64_t baug = desc_sw32(sbase + OFF_BAUG);
            const uint64_t aaug = desc_sw32(sbase + OFF_AAUG);
            const uint64_t bmain1 = desc_sw128(sbase + OFF_EF32);      // second D-chunk of a wide codebook
            const uint64_t a0 = desc_sw128(sbase + OFF_ARING);         // + ba * 1024 in the address field (16 KB buffers)
            // (32-bit counters throughout the per-tile loops: the tcgen05 path takes at most 2^31 - 1 rows)
            const int n_items = (int)my_items;
            int g = 0;                                                  // epilogue group of the tile: i % GROUPS
            for (int it = 0; it < n_items; ++it) {
                const int i = nd == 2 ? it >> 1 : it;                  // tile
                const int dc = nd == 2 ? (it & 1) : 0;                 // D-chunk
                const int ba = it & 1;                                 // A buffer
                const int b = i & 1;                                   // TMEM buffer
                mbar_wait<32>(bar(A_FULL + ba), (uint32_t)((it >> 1) & 1));
                if (dc == 0)
                    mbar_wait<32>(bar(T_EMPTY + b), (uint32_t)(((i >> 1) & 1) ^ 1));
                tc_fence_after();
                if (dc == 0 && lane == 0) stamp(i, 3);
                const bool wide = p.D - 32 * dc > 16;                  // else components 16..31 of this chunk are padding
                if (PAIR) {
                    if (elect_one()) {
                        const uint64_t a = a0 + (uint64_t)(ba * (16384 >> 4));
                        const uint32_t d = tmem_base + b * KMAX;
                        umma_bf16_2cta(d, a + 0, bmain + 0, idesc, 0);
                        if (wide) umma_bf16_2cta(d, a + 2, bmain + 2, idesc, 1);
                        umma_bf16_2cta(d, a + 0, bmain + 4, idesc, 1);
                        if (wide) umma_bf16_2cta(d, a + 2, bmain + 6, idesc, 1);
                        umma_bf16_2cta(d, a + 4, bmain + 0, idesc, 1);
                        if (wide) umma_bf16_2cta(d, a + 6, bmain + 2, idesc, 1);
                        umma_commit_2cta(bar(A_EMPTY + ba));
                        umma_bf16_2cta(d, aaug, baug, idesc, 1);
                        umma_commit_2cta(bar(T_FULL + g));
                    }
                } else if (elect_one()) {
                    const uint64_t a = a0 + (uint64_t)(ba * (16384 >> 4));
                    const uint64_t bm = dc ? bmain1 : bmain;
                    const uint32_t d = tmem_base + b * KMAX;
                    // K-slices of 16 bf16 = 32 bytes = +2 in the descriptor's address field
                    umma_bf16(d, a + 0, bm + 0, idesc, dc);                // z1[0:16]  . E1[0:16]
                    if (wide) umma_bf16(d, a + 2, bm + 2, idesc, 1);       // z1[16:32] . E1[16:32]
                    umma_bf16(d, a + 0, bm + 4, idesc, 1);                 // z1[0:16]  . E2[0:16]
                    if (wide) umma_bf16(d, a + 2, bm + 6, idesc, 1);       // z1[16:32] . E2[16:32]
                    umma_bf16(d, a + 4, bm + 0, idesc, 1);                 // z2[0:16]  . E1[0:16]
                    if (wide) umma_bf16(d, a + 6, bm + 2, idesc, 1);       // z2[16:32] . E1[16:32]
                    umma_commit(bar(A_EMPTY + ba));
                    if (dc == nd - 1) {
                        umma_bf16(d, aaug, baug, idesc, 1);                // + ee_k
                        umma_commit(bar(T_FULL + g));
                    }
                }
                __syncwarp();
                if (dc == nd - 1)
                    g = g + 1 == GROUPS ? 0 : g + 1;
            }
        }
    } else if (warp == W_SVC + 3) {
        // ================= ring owner: z_q TMA store, slot release, TMA refill =================
        // (Keeping one store in flight and releasing slot i when store i+1 is issued was measured:
        // the extra tile period of slot hold time costs more than the wait it hides.)
        // (warp-uniform loop, TMA instructions under elect_one: no election loops around UTMALDG / UTMASTG; bulk groups
        // belong to the issuing thread -- elect.sync picks the same lane of a full warp every time)
        {
            const int n_items = (int)my_items;
            auto load_item = [&](int it, int s) {
                const int i = nd == 2 ? it >> 1 : it;
                const int dc = nd == 2 ? (it & 1) : 0;
                const uint32_t tile = tile_first + (uint32_t)i * tile_step;
                if (elect_one()) {
                    mbar_expect_tx(bar(Z_FULL + s), TILE_M * D * 4);
                    tma_load_2d(sbase + OFF_ZRING + s * 16384, &map_z, bar(Z_FULL + s), dc * D, (int)(tile * TILE_M));
                }
                __syncwarp();
                if (dc == 0 && lane == 0) stamp(i, 0);
            };
            for (int it = 0; it < n_items && it < NST; ++it)
                load_item(it, it);
            int s = 0;
            uint32_t qph = 0;                                            // parity of Q_DONE[s]: (it / NST) & 1
            for (int it = 0; it < n_items; ++it) {
                mbar_wait<64>(bar(Q_DONE + s), qph);
                if (p.zq) {
                    const int i = nd == 2 ? it >> 1 : it;
                    const int dc = nd == 2 ? (it & 1) : 0;
                    const uint32_t tile = tile_first + (uint32_t)i * tile_step;
                    if (elect_one()) {
                        tma_store_2d(&map_zq, sbase + OFF_ZRING + s * 16384, dc * D, (int)(tile * TILE_M));
                        tma_store_commit();
                        tma_store_wait_read();     // the slot may be refilled once the store has read it
                    }
                    __syncwarp();
                }
                if ((nd == 1 || (it & 1)) && lane == 0) stamp(nd == 2 ? it >> 1 : it, 7);
                if (it + NST < n_items)
                    load_item(it + NST, s);
                if (++s == NST) {
                    s = 0;
                    qph ^= 1u;
                }
            }
            if (elect_one())
                tma_store_wait_all();
            __syncwarp();
        }
    } else if (warp >= W_CONV && warp < W_CONV + 4) {
        // ================= converters: fp32 -> bf16 hi/lo, thread = row =================
        if (TF32)
            reg_dec<80>();
        else
            reg_dec<56>();
        if (TF32) {
            // ================= refiner warps (TF32): warp w decides the uncertified rows of the tiles w, w + 4, ... ======
            // One lane per listed row: the row's oracle-order ||z||^2, the oracle-order distances of its (at most four)
            // candidate codes side by side (independent fmaf chains), lowest index among the smallest; then the row's
            // outputs exactly as the epilogue writes them -- z_q in place in the ring slot, idx, histogram, residual.
            const int w = warp - W_CONV;
            const unsigned char *ef32 = smem + OFF_EF32;
            const float *ees = reinterpret_cast<const float *>(smem + OFF_EE);
            unsigned *hist = reinterpret_cast<unsigned *>(smem + OFF_HIST);
            const int n_my = (int)my_tiles;
            float sqf = 0.0f;
            for (int i = w; i < n_my; i += 4) {
                const int s = i % NST;
                mbar_wait<512>(bar(R_FULL + s), (uint32_t)((i / NST) & 1));   // (long sleeps: this warp has three tile periods of slack, polling costs issue slots)
                if (lane == 0) stamp(i, 1);
                const unsigned cnt = rcnt[s];
                const uint32_t tile = tile_first + (uint32_t)i * tile_step;
                unsigned char *zt = smem + OFF_ZRING + s * 16384;
                // sixteen rows per pass, TWO LANES PER ROW: lane 2 j + t evaluates candidates 2 t and 2 t + 1 of row j (two fmaf
                // chains side by side), the pair picks the lowest index among the smallest, then writes chunks 4 t .. 4 t + 3 of
                // the row's z_q.  (One lane per row: a warp needs ~3500 clocks per pass whatever the number of active lanes --
                // dependent chains -- and four warps could not keep up with ~10 rows per tile.)
                const int t = lane & 1;
                for (unsigned base = 0; base < cnt; base += 16) {
                    const bool act = base + (lane >> 1) < cnt;
                    const uint2 ent = act ? rlist[s * TILE_M + base + (lane >> 1)] : make_uint2(0u, 0u);
                    const int rr = (int)ent.x, xr = (rr & 7) << 4;
                    const unsigned ma = ent.y & 0xffffu, mb = ent.y >> 16;
                    unsigned char *zrow = zt + rr * 128;
                    float zreg[D];
                    float zz = 0.0f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4 *>(zrow + ((c << 4) ^ xr));
                        zreg[4 * c] = v.x; zreg[4 * c + 1] = v.y; zreg[4 * c + 2] = v.z; zreg[4 * c + 3] = v.w;
                        zz = fmaf(v.x, v.x, zz); zz = fmaf(v.y, v.y, zz); zz = fmaf(v.z, v.z, zz); zz = fmaf(v.w, v.w, zz);
                    }
                    // candidate u = (A-group u / nb, B-group u % nb), u = 0..3 in ascending code order; this lane: 2 t, 2 t + 1
                    const int nb = __popc(mb), ncand = act ? __popc(ma) * nb : 0;
                    int ga[4], gb[4];
                    {
                        unsigned ra = ma, rb = mb;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            ga[u] = ra ? __ffs(ra) - 1 : 0; ra &= ra - 1u;
                            gb[u] = rb ? __ffs(rb) - 1 : 0; rb &= rb - 1u;
                        }
                    }
                    int kc[2];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int u = 2 * t + v;
                        const int ia = nb == 1 ? u : (nb == 2 ? u >> 1 : (nb == 3 ? (u == 3) : 0));
                        const int ib = nb == 1 ? 0 : (nb == 2 ? u & 1 : (nb == 3 ? (u == 3 ? 0 : u) : u));
                        const int a = ia == 0 ? ga[0] : (ia == 1 ? ga[1] : (ia == 2 ? ga[2] : ga[3]));
                        const int b = ib == 0 ? gb[0] : (ib == 1 ? gb[1] : (ib == 2 ? gb[2] : gb[3]));
                        const int k = (a << 4) | b;
                        kc[v] = (u < ncand && k < K) ? k : -1;
                    }
                    float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 e0 = *reinterpret_cast<const float4 *>(ef32 + ef32_off(kc[0] < 0 ? 0 : kc[0], c));
                        const float4 e1 = *reinterpret_cast<const float4 *>(ef32 + ef32_off(kc[1] < 0 ? 0 : kc[1], c));
                        acc0 = fmaf(zreg[4 * c], e0.x, acc0); acc1 = fmaf(zreg[4 * c], e1.x, acc1);
                        acc0 = fmaf(zreg[4 * c + 1], e0.y, acc0); acc1 = fmaf(zreg[4 * c + 1], e1.y, acc1);
                        acc0 = fmaf(zreg[4 * c + 2], e0.z, acc0); acc1 = fmaf(zreg[4 * c + 2], e1.z, acc1);
                        acc0 = fmaf(zreg[4 * c + 3], e0.w, acc0); acc1 = fmaf(zreg[4 * c + 3], e1.w, acc1);
                    }
                    float best = __int_as_float(0x7f800000);
                    int code = 0x7fffffff;
                    if (kc[0] >= 0) { best = ref_distance(zz, ees[kc[0]], acc0); code = kc[0]; }
                    if (kc[1] >= 0) {
                        const float d1 = ref_distance(zz, ees[kc[1]], acc1);
                        if (d1 < best) { best = d1; code = kc[1]; }          // ascending codes, strict <
                    }
                    {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, 1);
                        const int oc = __shfl_xor_sync(0xffffffffu, code, 1);
                        if (ob < best || (ob == best && oc < code)) {
                            best = ob;
                            code = oc;
                        }
                    }
                    if (code == 0x7fffffff)
                        code = 0;
                    // the row's outputs: z_q = z + (e - z) in place (chunks 4 t .. 4 t + 3), residual, idx, histogram
                    float r2 = 0.0f;
                    if (act && (p.zq || p.need_sq)) {
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const int c = 4 * t + h;
                            const float4 e = *reinterpret_cast<const float4 *>(ef32 + ef32_off(code, c));
                            const float z0 = t ? zreg[16 + 4 * h] : zreg[4 * h], z1 = t ? zreg[17 + 4 * h] : zreg[4 * h + 1];
                            const float z2 = t ? zreg[18 + 4 * h] : zreg[4 * h + 2], z3 = t ? zreg[19 + 4 * h] : zreg[4 * h + 3];
                            const float d0 = __fsub_rn(e.x, z0), d1 = __fsub_rn(e.y, z1), d2 = __fsub_rn(e.z, z2), d3 = __fsub_rn(e.w, z3);
                            r2 = fmaf(d0, d0, r2); r2 = fmaf(d1, d1, r2); r2 = fmaf(d2, d2, r2); r2 = fmaf(d3, d3, r2);
                            if (p.zq)
                                *reinterpret_cast<float4 *>(zrow + ((c << 4) ^ xr)) =
                                    make_float4(__fadd_rn(z0, d0), __fadd_rn(z1, d1), __fadd_rn(z2, d2), __fadd_rn(z3, d3));
                        }
                    }
                    r2 += __shfl_xor_sync(0xffffffffu, r2, 1);
                    if (act && t == 0) {
                        const uint32_t row = tile * TILE_M + (uint32_t)rr;
                        if (p.idx)
                            p.idx[row] = code;
                        atomicAdd(hist + code, 1u);
                        sqf += r2;
                    }
                }
                sq += (double)sqf;
                sqf = 0.0f;
                // (the counter is cleared by the epilogue group when it takes the slot for its next tile)
                if (p.zq)
                    fence_proxy_async();               // z_q rows (generic proxy) -> visible to the TMA store
                mbar_arrive(bar(Q_DONE + s));
                if (lane == 0) stamp(i, 2);
            }
        }
        const int r = tid - W_CONV * 32;
        const int x = (r & 7) << 4;
        const int n_items = TF32 ? 0 : (int)my_items;    // (TF32: nothing to convert)
        int s = 0;
        uint32_t zph = 0;                                               // parity of Z_FULL[s]: (i / NST) & 1
        for (int i = 0; i < n_items; ++i) {              // i: pipeline item (= tile, or half a wide tile)
            const int b = i & 1;
            if (warp == W_CONV) {
                mbar_wait<128>(bar(Z_FULL + s), zph);
                if (r == 0 && nd == 1) stamp(i, 1);
                mbar_wait<128>(bar(A_EMPTY + b), (uint32_t)(((i >> 1) & 1) ^ 1));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const unsigned char *zrow = smem + OFF_ZRING + s * 16384 + r * 128;
            unsigned char *arow = smem + OFF_ARING + b * 16384 + r * 128;
            float2 zp[4];                            // ||z||^2 for the epilogue's filter radius (a bound, not the decision)
#pragma unroll
            for (int h = 0; h < 4; ++h)
                zp[h] = make_float2(0.f, 0.f);
#pragma unroll
            for (int cp = 0; cp < 4; ++cp) {        // two 16-byte chunks of z -> one chunk of z1 and one of z2
                const float4 va = *reinterpret_cast<const float4 *>(zrow + (((2 * cp) << 4) ^ x));
                const float4 vb = *reinterpret_cast<const float4 *>(zrow + (((2 * cp + 1) << 4) ^ x));
                const float xs[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float x0 = xs[2 * h], x1 = xs[2 * h + 1];
                    zp[h] = __ffma2_rn(make_float2(x0, x1), make_float2(x0, x1), zp[h]);
                    // z1 = rn_bf16(x), z2 = rn_bf16(x - z1).  (Measured alternatives: truncating instead of
                    // rounding saves two ALU ops per pair but widens the filter radius by 60 %; a Veltkamp split
                    // on the FMA pipe relieves the ALU pipe but issues four more instructions per pair, ~1 % slower.)
                    // (the halves are widened with one shift and one mask; __low2float / __high2float compile to a
                    // PRMT + shift pair each)
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(x0, x1);
                    const uint32_t hb = *reinterpret_cast<const uint32_t *>(&h2);
                    const float2 lo2 = __fadd2_rn(make_float2(x0, x1), make_float2(-__uint_as_float(hb << 16),
                                                                                   -__uint_as_float(hb & 0xffff0000u)));
                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(lo2.x, lo2.y);
                    hi[h] = hb;
                    lo[h] = *reinterpret_cast<const uint32_t *>(&l2);
                }
                *reinterpret_cast<uint4 *>(arow + ((cp << 4) ^ x)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4 *>(arow + (((cp + 4) << 4) ^ x)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            {
                const float2 t = __fadd2_rn(__fadd2_rn(zp[0], zp[1]), __fadd2_rn(zp[2], zp[3]));
                reinterpret_cast<float *>(smem + OFF_ZZ + s * 512)[r] = t.x + t.y;
            }
            fence_proxy_async();
            if (PAIR) {                                       // the leader issues for both tiles: one arrival per warp
                __syncwarp();
                if (lane == 0)
                    mbar_arrive_cluster(bar(A_FULL + b), 0);
            } else {
                mbar_arrive(bar(A_FULL + b));
            }
            if (r == 0 && nd == 1) stamp(i, 2);
            if (++s == NST) {
                s = 0;
                zph ^= 1u;
            }
        }
    } else if (warp >= W_EPI && warp < W_EPI + 4 * GROUPS) {
        // ================= epilogue groups =================
        if (TF32)
            reg_inc<120>();                       // (the refiner warps take 80 each: 128 x (40 + 80 + 3 x 120) = 61440)
        else
            reg_inc<VQB_EPI_REGS>();
        const int g = (warp - W_EPI) >> 2;            // tile i is handled by group i % GROUPS, TMEM buffer i & 1
        const int q = warp & 3;                   // TMEM lane quarter of this warp
        const int r = q * 32 + lane;              // row in tile = TMEM lane
        const int x = (r & 7) << 4;
        const unsigned char *ef32 = smem + OFF_EF32;
        const float *ees = reinterpret_cast<const float *>(smem + OFF_EE);
        unsigned *hist = reinterpret_cast<unsigned *>(smem + OFF_HIST);
        const float eemax = __uint_as_float(cst->emax2_bits);
        const float emax = sqrt_approx(eemax) * 1.00001f;
        const bool poisoned = p.hdr_in->poisoned_columns != 0;
        const bool cb_bad = cst->nonfinite != 0 || poisoned;
        const int n_slab = kp >> 5;
        unsigned long long n_slow_total = 0;
        uint2 *wl = reinterpret_cast<uint2 *>(const_cast<unsigned char *>(img) + IMG_WL) + (size_t)blockIdx.x * WL_CAP;

        // accumulator drained (this thread's share): in a pair the leader's barrier counts both CTAs' epilogue threads
        auto t_empty_arrive = [&](int b) {
            if (PAIR) {
                __syncwarp();
                if (lane == 0)
                    mbar_arrive_cluster(bar(T_EMPTY + b), 0);
            } else {
                mbar_arrive(bar(T_EMPTY + b));
            }
        };
        float sqf = 0.0f;
        const int n_my = (int)my_tiles;
        constexpr int s_step = (GROUPS * nd) % NST;
        int s = (g * nd) % NST;                            // ring slot of the tile (its first half if wide)
        uint32_t ph = 0;                                      // parity of T_FULL[g]: (i / GROUPS) & 1
        int run = 0;
        for (int i = g; i < n_my; i += GROUPS, ph ^= 1u, s = s + s_step >= NST ? s + s_step - NST : s + s_step) {
            const int s1 = s + 1 == NST ? 0 : s + 1;       // second half of a wide tile
            const int b = i & 1;
            const uint32_t tile = tile_first + (uint32_t)i * tile_step;
            const uint32_t row = tile * TILE_M + r;
            const bool ok = row < (uint32_t)n_rows;

            // ---- filter: minima of the approximate scores over the 16 A-groups and the 16 B-groups ----
            if (TF32) {
                // (the accumulator is ready => the slot holds THIS tile => the previous tile's refiner warps are done with
                // its list: the leader clears the counter before anyone of the group can push)
                if (q == 0) {
                    mbar_wait<64>(bar(T_FULL + g), ph);
                    if (lane == 0)
                        rcnt[s] = 0u;
                }
                named_bar_sync(2 + g, 128);
            } else {
                group_wait<64>(q == 0, bar(T_FULL + g), ph, 2 + g);
            }
            tc_fence_after();
            if (r == 0) stamp(i, 4);
            const uint32_t taddr = tmem_base + b * KMAX + ((uint32_t)(q * 32) << 16);
            const float inf = __int_as_float(0x7f800000);
            // (finite sentinel: a group id packed into the low mantissa bits of +inf would make a NaN key)
            const float big = 3.0e38f;
            float amin[16], bmin[16];
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                amin[t] = big;
                bmin[t] = big;
            }
            auto reduce_slab = [&](const uint32_t (&v)[32], int sl) {
                amin[2 * sl] = min16(&v[0]);
                amin[2 * sl + 1] = min16(&v[16]);
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    bmin[t] = min3(bmin[t], __uint_as_float(v[t]), __uint_as_float(v[t + 16]));
            };
#if VQB_LDPIPE
            // Two slabs in flight: the load of slab s + 1 is issued before slab s is reduced, so the TMEM read
            // (64 B/clk per SM sub-partition) overlaps the min trees, and the accumulator is handed back the moment
            // its last slab sits in registers -- before that slab's min trees.  n_slab is CTA-uniform.
            {
                uint32_t va[32], vb[32];
                tmem_ld32(taddr, va);
#pragma unroll
                for (int sl = 0; sl < KMAX / 32; sl += 2) {
                    if (sl < n_slab) {
                        tmem_wait_ld_fence(va);
                        if (sl + 1 < n_slab) {
                            tmem_ld32(taddr + (sl + 1) * 32, vb);
                        } else {
                            tc_fence_before();
                            t_empty_arrive(b);
                        }
                        reduce_slab(va, sl);
                        if (sl + 1 < n_slab) {
                            tmem_wait_ld_fence(vb);
                            if (sl + 2 < n_slab) {
                                tmem_ld32(taddr + (sl + 2) * 32, va);
                            } else {
                                tc_fence_before();
                                t_empty_arrive(b);
                            }
                            reduce_slab(vb, sl + 1);
                        }
                    }
                }
            }
#else
#pragma unroll
            for (int sl = 0; sl < KMAX / 32; ++sl) {
                if (sl < n_slab) {                // CTA-uniform: slabs beyond the padded codebook are never read
                    uint32_t v[32];
                    tmem_ld32(taddr + sl * 32, v);
                    tmem_wait_ld_fence(v);
                    reduce_slab(v, sl);
                }
            }
            tc_fence_before();
            t_empty_arrive(b);        // accumulator drained: the next MMA may overwrite it
#endif
            if (TRACE && lane == 0 && blockIdx.x < kTraceCtas && i < kTraceTiles)    // the last of the group's four warps
                atomicMax(&trace[((size_t)blockIdx.x * kTraceTiles + i) * kTraceEvents + 5], (unsigned long long)clock64());
            // ---- the vector itself (TMA-written ring slot) ----
            // (the tile itself was TMA-written before the converters read it, i.e. long before T_FULL)
            unsigned char *zt = smem + OFF_ZRING + s * 16384;
            unsigned char *zrow = zt + r * 128;
            unsigned char *zt1 = smem + OFF_ZRING + s1 * 16384;
            unsigned char *zrow1 = zt1 + r * 128;
            // ||z||^2 comes from the converter that already had the row in registers (ordered before us
            // by A_FULL -> MMA -> T_FULL); it only bounds the filter radius, it is not part of the decision
            float zz = 0.0f;
            if (TF32) {
                // no converter: the row's own thread sums the squares (it only bounds the filter radius)
                float zp4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4 *>(zrow + ((c << 4) ^ x));
                    zp4[0] = fmaf(v.x, v.x, zp4[0]); zp4[1] = fmaf(v.y, v.y, zp4[1]);
                    zp4[2] = fmaf(v.z, v.z, zp4[2]); zp4[3] = fmaf(v.w, v.w, zp4[3]);
                }
                zz = (zp4[0] + zp4[1]) + (zp4[2] + zp4[3]);
            } else {
                zz = reinterpret_cast<const float *>(smem + OFF_ZZ + s * 512)[r];
                if (nd == 2)
                    zz += reinterpret_cast<const float *>(smem + OFF_ZZ + s1 * 512)[r];
            }
            // (sums of squares below 2^-120 are treated as 2^-120: the fp32 sum loses bits there -- it underflows to 0 for
            // components below 2^-75 -- and sqrt.approx.ftz would flush a sub-normal sum to 0; 2^-60 still bounds the norm)
            const float zn = sqrt_approx(fmaxf(zz, 7.52316385e-37f)) * 1.00001f;
            // Filter radius delta = 2*eps + 2*H + 2*pack (DESIGN.md "Exactness"), u = 2^-8 the bf16 unit
            // roundoff, zn >= |z|, emax >= max|e_k|, eemax = max ee_k, S = sum_j |z_j e_j| <= zn*emax:
            //   eps : terms the three products leave out, z2.E2 + r_z.E + z.r_E <= (u*u + u*u + u*u) S = 3*2^-16 S,
            //         doubled by the -2 scaling, plus tensor-core accumulation slack 2^-19 (2S + ee)
            //   H   : the oracle's own fp32 roundings fl(zz+ee), fl(t-u) and its D-step fmaf chain:
            //         <= 2^-24 [2 (zn+emax)^2 + 2 D S]
            //   pack: group ids in 4 mantissa bits of the keys: <= 2^-19 (2S + ee)
            //   =>  delta <= (12*2^-16 + 2^-17 + 2^-17 + D*2^-22) zn*emax + 2^-17 eemax + 2^-22 (zn+emax)^2
            //   + an absolute floor for sub-normal z entries (lost by the bf16 split and possibly flushed by the
            //     tensor core): each is < 2^-126, so they move a score by < 2*sqrt(D)*2^-126*emax
            // TF32 variant: the tensor core sees z truncated or rounded to 11 significant bits (|dz| <= 2^-10 |z|) and E rounded
            // to nearest (|dE| <= 2^-11 |E|): eps <= 2 (2^-10 + 2^-11 + 2^-21) S + 2^-19 (2S + ee), so
            //   delta <= (6*2^-10 + 2^-17 + D*2^-22) zn*emax + ... = 5.9e-3 zn*emax + (the same other terms)
            const float delta = (TF32 ? 5.9e-3f : 2.1e-4f) * zn * emax + 8.0e-6f * eemax + 3.0e-7f * (zn + emax) * (zn + emax) +
                                (1.0e-35f + 1.0e-36f * emax);
            // Decision on the group minima alone: with m the smallest approximate score and thr = m + delta, every code
            // whose approximate score lies within delta of m sits in an A-group AND a B-group whose minimum is <= thr.
            // If exactly one A-group and exactly one B-group qualify, exactly one code does (two codes of one A-group
            // never share a B-group), it is the smallest one, and every other code is more than delta above it: the
            // oracle's argmin, certified by the tensor cores alone (~99.8 % of vectors).  The masks are built as exact
            // float sums of powers of two -- one FSET per group on the ALU pipe, the accumulation on the FMA pipe.
            const float thr = min16f(amin) + delta;          // (min over the A-groups = min over all scores)
            float maf = 0.0f, mbf = 0.0f;
#pragma unroll
            for (int t = 0; t < 16; ++t) {
                maf = fmaf(amin[t] <= thr ? 1.0f : 0.0f, (float)(1 << t), maf);
                mbf = fmaf(bmin[t] <= thr ? 1.0f : 0.0f, (float)(1 << t), mbf);
            }
            const unsigned ma = __float_as_uint(maf + 8388608.0f) & 0xffffu;     // integer part of an exact sum < 2^16
            const unsigned mb = __float_as_uint(mbf + 8388608.0f) & 0xffffu;
            const bool one = ma != 0u && mb != 0u && (ma & (ma - 1u)) == 0u && (mb & (mb - 1u)) == 0u;
            const bool certain = one && (zz <= 3.0e38f) && !cb_bad;
            int code = (int)((31 - __clz(ma | 1u)) << 4 | (31 - __clz(mb | 1u)));
            if (code >= K)
                code = 0;
            // Vectors the filter cannot certify (~0.2 %: near-ties) are queued, together with the codes the filter
            // could not rule out (the two masks), for the fix-up kernel that follows.  Rows whose exact distances could
            // overflow, non-finite rows / codebooks, crowded candidate sets (exact ties of many codes) and a full queue
            // take the warp-wide exact scan of all K codes right here instead.
            bool slow = false, deferred = false;
            bool decided = certain;
            bool refined = false;
            if (TF32) {
                // up to four candidate codes (the cross product of the qualifying A- and B-groups): the row goes to the tile's
                // list and is decided -- and written -- by a refiner warp; this thread leaves the row alone
                refined = !certain && ok && zz <= 1.0e37f && eemax <= 1.0e37f && !cb_bad &&
                          __popc(ma) * __popc(mb) >= 1 && __popc(ma) * __popc(mb) <= 4;
                if (refined) {
                    const unsigned pos = atomicAdd(rcnt + s, 1u);
                    rlist[s * TILE_M + pos] = make_uint2((unsigned)r, ma | (mb << 16));
                    decided = true;
                }
                mbar_arrive(bar(R_FULL + s));      // this thread's entry (if any) is in the tile's list: the refiner warp
                                                   // works on it while this thread writes its own row's outputs
            }
            if (!decided && ok) {
                slow = true;
                if (zz <= 1.0e37f && eemax <= 1.0e37f && !cb_bad) {
                    const int nc = __popc(ma) * __popc(mb);
                    if (nc >= 1 && nc <= MAX_CAND) {
                        const unsigned pos = atomicAdd(wl_count_s, 1u);
                        if (pos < (unsigned)WL_CAP) {
                            wl[pos] = make_uint2((unsigned)row, ma | (mb << 16));
                            slow = false;
                            deferred = true;
                        }
                    }
                }
            }
            unsigned need = __ballot_sync(0xffffffffu, slow);
            n_slow_total += __popc(__ballot_sync(0xffffffffu, !certain && ok));
            (void)decided;
            while (need) {
                const int src = __ffs(need) - 1;
                need &= need - 1;
                const int res = nd == 2 ? warp_full_scan_wide(zt, zt1, q * 32 + src, p.E, p.D, ees, K)
                                        : warp_full_scan(zt, q * 32 + src, ef32, ees, K);
                if (lane == src)
                    code = res;
            }

            // ---- chunk pass of a large codebook: the chunk's winner and its exact distance join the running best ----
            if (p.chunk_mode) {
                if (ok && !deferred)
                    merge_running(p.run, row,
                                  nd == 2 ? exact_distance_wide(zrow, zrow1, x, p.E, p.D, ees, code)
                                          : exact_distance(zrow, x, ef32, ees, code),
                                  p.code_base + code, p.chunk_mode);
                mbar_arrive(bar(Q_DONE + s));
                if (nd == 2)
                    mbar_arrive(bar(Q_DONE + s1));
                if (r == 0) stamp(i, 6);
                continue;
            }

            // ---- outputs: idx, histogram, loss, z_q (in place in the ring slot) ----
            const bool emit = ok && !deferred && !refined;
            if (emit) {
                if (p.idx)
                    p.idx[row] = code;
                atomicAdd(hist + code, 1u);
            }
            // (ids-only calls -- no z_q, no loss -- skip the gather and the residual altogether)
            float r2 = 0.0f;
            if ((p.zq || p.need_sq) && nd == 1 && !poisoned) {
                // (the whole warp walks its 32 rows together; rows written elsewhere are skipped)
                sqf += emit_rows_coop(zt, q, lane, ef32, emit ? code : -1, p.zq != nullptr);
            } else if ((p.zq || p.need_sq) && !refined) {
                if (nd == 1) {
                    // (keeping the row in registers from the ||z||^2 pass for this write was measured: 108 bytes of spills at
                    // 120 registers, 1.04 ms instead of 0.96)
                    r2 = poisoned ? emit_row<true>(zrow, x, ef32, code, p.zq != nullptr, p.colcnt, p.colwhich)
                                  : emit_row<false>(zrow, x, ef32, code, p.zq != nullptr, nullptr, nullptr);
                } else if (poisoned) {
                    r2 = emit_row_wide<true>(zrow, x, p.E, p.D, 0, code, p.zq != nullptr, p.colcnt, p.colwhich) +
                         emit_row_wide<true>(zrow1, x, p.E, p.D, D, code, p.zq != nullptr, p.colcnt, p.colwhich);
                } else {
                    r2 = emit_row_wide<false>(zrow, x, p.E, p.D, 0, code, p.zq != nullptr, nullptr, nullptr) +
                         emit_row_wide<false>(zrow1, x, p.E, p.D, D, code, p.zq != nullptr, nullptr, nullptr);
                }
            }
            if (emit)
                sqf += r2;
            if (++run == 16) {                    // bounded fp32 run lengths, fp64 across them
                sq += (double)sqf;
                sqf = 0.0f;
                run = 0;
            }
            if (p.zq)
                fence_proxy_async();               // z_q rows (generic proxy) -> visible to the TMA store
            mbar_arrive(bar(Q_DONE + s));          // warp 3 stores the tile and frees the slot
            if (nd == 2)
                mbar_arrive(bar(Q_DONE + s1));
            if (r == 0) stamp(i, 6);
        }
        sq += (double)sqf;
        if (p.stats && n_slow_total && lane == 0)
            atomicAdd(p.stats + 1, n_slow_total);
    }

    // ---- teardown ----------------------------------------------------------------------
    tc_fence_before();
    if (PAIR)
        cluster_sync_all();       // neither CTA of a pair may leave (or free tensor memory) while the other still signals it
    else
        __syncthreads();
    tc_fence_after();
    if (warp == W_SVC + 2) {
        if (PAIR)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
    if (!WIDE && p.chunk_mode == 0) {
        // ---- the CTA's own queue of uncertified vectors, decided here by all its warps (one warp per vector, one lane
        // per candidate code, exact oracle-order distances; codebook rows from shared memory, the vector from global
        // memory / L2) instead of by a second kernel: saves a launch, its ramp and its tail (~3 % of the call).  The
        // tile's TMA store has completed (tma_store_wait_all above the barrier), so the row written here is final.
        const unsigned count = min(*wl_count_s, (unsigned)WL_CAP);
        const uint2 *wl = reinterpret_cast<const uint2 *>(img + IMG_WL) + (size_t)blockIdx.x * WL_CAP;
        const unsigned char *ef32 = smem + OFF_EF32;
        const float *ees = reinterpret_cast<const float *>(smem + OFF_EE);
        unsigned *hist = reinterpret_cast<unsigned *>(smem + OFF_HIST);
        const int d = p.D;
        uint2 ent = warp < (int)count ? wl[warp] : make_uint2(0u, 0u);
        float zj = (warp < (int)count && lane < d) ? __ldg(p.z.base + (size_t)ent.x * d + lane) : 0.0f;
        for (unsigned e = warp; e < count; e += THREADS / 32) {
            const uint2 cur = ent;
            const float zc = zj;
            const unsigned nxt = e + THREADS / 32;
            if (nxt < count) {                            // next entry's loads overlap this entry's arithmetic
                ent = wl[nxt];
                zj = lane < d ? __ldg(p.z.base + (size_t)ent.x * d + lane) : 0.0f;
            }
            const unsigned ma = cur.y & 0xffffu, mb = cur.y >> 16;
            const int nb = __popc(mb), nc = __popc(ma) * nb;
            int k = -1;
            if (lane < nc) {                              // lane j: the j-th candidate in ascending code order
                k = 16 * __fns(ma, 0, lane / nb + 1) + __fns(mb, 0, lane % nb + 1);
                if (k >= K)
                    k = -1;
            }
            const int kk = k < 0 ? 0 : k;
            float zz = 0.0f, acc = 0.0f;
#pragma unroll
            for (int c = 0; c < 8; ++c) {                 // oracle-order chains, ascending j (columns beyond d are zero)
                const float4 e4 = *reinterpret_cast<const float4 *>(ef32 + ef32_off(kk, c));
                const float z0 = __shfl_sync(0xffffffffu, zc, 4 * c), z1 = __shfl_sync(0xffffffffu, zc, 4 * c + 1);
                const float z2 = __shfl_sync(0xffffffffu, zc, 4 * c + 2), z3 = __shfl_sync(0xffffffffu, zc, 4 * c + 3);
                zz = fmaf(z0, z0, zz); zz = fmaf(z1, z1, zz); zz = fmaf(z2, z2, zz); zz = fmaf(z3, z3, zz);
                acc = fmaf(z0, e4.x, acc); acc = fmaf(z1, e4.y, acc); acc = fmaf(z2, e4.z, acc); acc = fmaf(z3, e4.w, acc);
            }
            float best = __int_as_float(0x7f800000);
            int bidx = 0x7fffffff;
            if (k >= 0) {
                best = ref_distance(zz, ees[k], acc);
                bidx = k;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
                if (ob < best || (ob == best && oi < bidx)) {
                    best = ob;
                    bidx = oi;
                }
            }
            const int code = bidx == 0x7fffffff ? 0 : bidx;
            float r2 = 0.0f;
            if (lane < d) {
                const float ej = *reinterpret_cast<const float *>(ef32 + ef32_off(code, lane >> 2) + (lane & 3) * 4);
                const float diff = __fsub_rn(ej, zc);
                if (p.zq)
                    p.zq[(size_t)cur.x * d + lane] = __fadd_rn(zc, diff);
                r2 = diff * diff;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                r2 += __shfl_xor_sync(0xffffffffu, r2, o);
            if (lane == 0) {
                if (p.idx)
                    p.idx[cur.x] = code;
                atomicAdd(hist + code, 1u);
                sq += (double)r2;
            }
        }
        __syncthreads();
    } else if (tid == 0) {
        const unsigned n = *wl_count_s;
        reinterpret_cast<unsigned *>(const_cast<unsigned char *>(img) + IMG_WLCOUNT)[blockIdx.x] =
            n < (unsigned)WL_CAP ? n : (unsigned)WL_CAP;
    }
    if (tid < K) {
        const unsigned c = reinterpret_cast<unsigned *>(smem + OFF_HIST)[tid];
        if (c)
            atomicAdd(p.counts + tid, (unsigned long long)c);
    }
    // per-CTA sum of squared residuals (deterministic order within the CTA)
    __shared__ double red[THREADS / 32];
    sq = warp_sum(sq);
    if (lane == 0)
        red[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < THREADS / 32; ++w)
            t += red[w];
        p.partials[blockIdx.x] = p.accumulate ? p.partials[blockIdx.x] + t : t;
        if (p.stats)              // stats[3]: SM clocks of the longest-running CTA (cycles per tile, SM clock inside the launch)
            atomicMax(p.stats + 3, (unsigned long long)(clock64() - clk_begin));
    }
}

// ---------------------------------------------------------------------------------------
// fix-up: each queued vector is decided among its candidate codes with the exact expression
// (one warp per vector, one lane per candidate code), and its idx / z_q / histogram / loss
// contributions are written here.  The main kernel only queues vectors whose distances
// cannot overflow, so no NaN rule is needed.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vq_tc_fixup_kernel(const FwdParams p, const unsigned char *__restrict__ img,
                                                           double *__restrict__ partial_out)
{
    using namespace tc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int queue = blockIdx.x / FIX_SPLIT, part = blockIdx.x % FIX_SPLIT;
    const unsigned count = reinterpret_cast<const unsigned *>(img + IMG_WLCOUNT)[queue];
    const uint2 *wl = reinterpret_cast<const uint2 *>(img + IMG_WL) + (size_t)queue * WL_CAP;
    const float *ee = reinterpret_cast<const float *>(img + IMG_EE);
    const int K = p.K;
    double sq = 0.0;
    for (unsigned e = part * 8 + warp; e < count; e += 8 * FIX_SPLIT) {
        const uint2 ent = wl[e];
        const int64_t row = ent.x;
        const unsigned ma = ent.y & 0xffffu, mb = ent.y >> 16;
        const int nb = __popc(mb), nc = __popc(ma) * nb;
        // lane j takes the j-th candidate in ascending code order (A-group major)
        int k = -1;
        if (lane < nc) {
            const int a = __fns(ma, 0, lane / nb + 1), b = __fns(mb, 0, lane % nb + 1);
            k = 16 * a + b;
            if (k >= K)
                k = -1;
        }
        const int d = p.D;                            // real row width (<= 64, a multiple of 4)
        const float zj0 = lane < d ? __ldg(p.z.base + row * d + lane) : 0.0f;
        const float zj1 = lane + 32 < d ? __ldg(p.z.base + row * d + lane + 32) : 0.0f;
        const int kk = k < 0 ? 0 : k;
        float zz = 0.0f, acc = 0.0f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {                // oracle-order chains, ascending j
            if (4 * c < d) {                          // warp-uniform
                const float4 e4 = __ldg(reinterpret_cast<const float4 *>(p.E + (size_t)kk * d) + c);
                const float src = c < 8 ? zj0 : zj1;
                const int j = (4 * c) & 31;
                const float z0 = __shfl_sync(0xffffffffu, src, j), z1 = __shfl_sync(0xffffffffu, src, j + 1);
                const float z2 = __shfl_sync(0xffffffffu, src, j + 2), z3 = __shfl_sync(0xffffffffu, src, j + 3);
                zz = fmaf(z0, z0, zz); zz = fmaf(z1, z1, zz); zz = fmaf(z2, z2, zz); zz = fmaf(z3, z3, zz);
                acc = fmaf(z0, e4.x, acc); acc = fmaf(z1, e4.y, acc); acc = fmaf(z2, e4.z, acc); acc = fmaf(z3, e4.w, acc);
            }
        }
        float best = __int_as_float(0x7f800000);
        int bidx = 0x7fffffff;
        if (k >= 0) {
            best = ref_distance(zz, ee[k], acc);
            bidx = k;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
            if (ob < best || (ob == best && oi < bidx)) {
                best = ob;
                bidx = oi;
            }
        }
        const int code = bidx == 0x7fffffff ? 0 : bidx;
        if (p.chunk_mode) {
            if (lane == 0)
                merge_running(p.run, row, best, p.code_base + code, p.chunk_mode);
            continue;
        }
        float r2 = 0.0f;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = lane + 32 * h;
            if (j < d) {
                const float zj = h ? zj1 : zj0;
                const float diff = __fsub_rn(__ldg(p.E + (size_t)code * d + j), zj);
                if (p.zq)
                    p.zq[row * d + j] = __fadd_rn(zj, diff);
                r2 = fmaf(diff, diff, r2);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        if (lane == 0) {
            if (p.idx)
                p.idx[row] = code;
            atomicAdd(p.counts + code, 1ULL);
            sq += (double)r2;
        }
    }
    __shared__ double red[8];
    if (lane == 0)
        red[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += red[w];
        partial_out[blockIdx.x] = p.accumulate ? partial_out[blockIdx.x] + t : t;
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
// D < 32 (a multiple of 4: TMA needs 16-byte row pitches) runs as D = 32 with zero columns: the tensor maps
// describe (N, D) tensors under 32-wide boxes, so loads zero-fill and stores clip the columns beyond D.
// 32 < D <= 64: two 32-component D-chunks per tile (two ring slots, products accumulated in the same TMEM buffer).
bool tc_shape_supported(int K, int D) { return D >= 4 && D <= 2 * tc::D && D % 4 == 0 && K >= 1 && K <= tc::KMAX; }

namespace {

bool make_map(CUtensorMap *map, const float *base, int64_t n_rows, int d)
{
    return tc::make_tensor_map_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, n_rows, d, tc::TILE_M, tc::D,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

}  // namespace

// ---------------------------------------------------------------------------------------
// finish kernel of the chunked path: the running best holds every vector's code; gather, z_q, loss,
// histogram and the int64 index are produced in one streaming pass (eight threads per vector).
// ---------------------------------------------------------------------------------------
constexpr int kFinishMaxK = 16384;     // shared-memory histogram (u32 per code)
// NF: float4 columns per thread (columns sub, sub + 8, ...: rows of up to 32 * NF components); U vectors per thread and step
template <int NF>
__global__ void __launch_bounds__(256) vq_tc_finish_kernel(const FwdParams p, double *__restrict__ partial_out)
{
    constexpr int U = NF > 2 ? 2 : 4;
    extern __shared__ unsigned fhist[];
    const int tid = threadIdx.x, sub = tid & 7, d = p.D;
    const int hist_k = p.chunk_mode == 3 ? 0 : p.K;
    for (int t = tid; t < hist_k; t += 256)
        fhist[t] = 0u;
    __syncthreads();
    const bool poisoned = p.hdr_in->poisoned_columns != 0;
    double sq = 0.0;
    // four vectors per thread and step, all of a step's loads issued before anything depends on them
    // (code -> E[code] is a dependent chain: one vector at a time is latency-bound, measured 1.48 ms -> see profiles)
    const int nf = (d + 31) / 32;                     // float4 columns per thread: sub, sub + 8, ... (wide rows)
    for (int64_t row0 = (int64_t)blockIdx.x * (32 * U) + (tid >> 3); row0 < p.z.n_rows; row0 += (int64_t)gridDim.x * (32 * U)) {
        int code[U];
        float4 zv[U][NF];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = row0 + 32 * u;
            code[u] = row < p.z.n_rows ? (int)(unsigned)(p.run[row] & 0xffffffffull) : -1;
#pragma unroll
            for (int h = 0; h < NF; ++h) {
                const int f = sub + 8 * h;
                zv[u][h] = (row < p.z.n_rows && h < nf && 4 * f < d)
                               ? __ldg(reinterpret_cast<const float4 *>(p.z.base + row * d) + f)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        float4 ev[U][NF];
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int h = 0; h < NF; ++h) {
                const int f = sub + 8 * h;
                ev[u][h] = (code[u] >= 0 && h < nf && 4 * f < d)
                               ? __ldg(reinterpret_cast<const float4 *>(p.E + (size_t)code[u] * d) + f)
                               : zv[u][h];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t row = row0 + 32 * u;
            if (code[u] < 0)
                continue;
#pragma unroll
            for (int h = 0; h < NF; ++h) {
                const int f = sub + 8 * h;
                if (h >= nf || 4 * f >= d)
                    continue;
                float4 e = ev[u][h];
                if (poisoned) {   // gather-by-GEMM semantics for a non-finite codebook (oracle column_poison)
                    float *evp = reinterpret_cast<float *>(&e);
                    for (int t = 0; t < 4; ++t) {
                        const int j = 4 * f + t, cc = p.colcnt[j];
                        if (!(cc == 0 || (cc == 1 && p.colwhich[j] == code[u] + 1)))
                            evp[t] = __int_as_float(0x7fc00000);
                    }
                }
                const float4 z4 = zv[u][h];
                float4 o;
                float dj, rs = 0.0f;
                dj = __fsub_rn(e.x, z4.x); rs = fmaf(dj, dj, rs); o.x = __fadd_rn(z4.x, dj);
                dj = __fsub_rn(e.y, z4.y); rs = fmaf(dj, dj, rs); o.y = __fadd_rn(z4.y, dj);
                dj = __fsub_rn(e.z, z4.z); rs = fmaf(dj, dj, rs); o.z = __fadd_rn(z4.z, dj);
                dj = __fsub_rn(e.w, z4.w); rs = fmaf(dj, dj, rs); o.w = __fadd_rn(z4.w, dj);
                if (p.zq)
                    __stcs(reinterpret_cast<float4 *>(p.zq + row * d) + f, o);
                sq += (double)rs;
            }
            if (sub == 0 && p.chunk_mode != 3) {      // (chunk_mode 3: ids and histogram are final already, vq_fwd_tcs.cu)
                p.idx[row] = code[u];     // same 8 bytes the running best lived in
                atomicAdd(fhist + code[u], 1u);
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < hist_k; t += 256) {
        const unsigned c = fhist[t];
        if (c)
            atomicAdd(p.counts + t, (unsigned long long)c);
    }
    __shared__ double red[8];
    sq = warp_sum(sq);
    if ((tid & 31) == 0)
        red[tid >> 5] = sq;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += red[w];
        partial_out[blockIdx.x] = p.accumulate ? partial_out[blockIdx.x] + t : t;
    }
}

// z_q and the loss for ids that are already final in p.idx (the tile-stationary kernel of vq_fwd_tcs.cu)
cudaError_t launch_tc_finish_ids(const FwdParams &p, int sm_count, int *n_ctas, cudaStream_t st)
{
    FwdParams pf = p;
    pf.run = reinterpret_cast<unsigned long long *>(p.idx);
    pf.chunk_mode = 3;
    const int64_t groups = (p.z.n_rows + 127) / 128;
    int grid = (int)(groups < (int64_t)sm_count * 6 ? groups : (int64_t)sm_count * 6);
    if (grid < 1)
        grid = 1;
    if (grid > kMaxPartials)
        grid = kMaxPartials;
    // (no histogram in this mode: one word of dynamic shared memory)
    if (p.D <= 32)
        vq_tc_finish_kernel<1><<<grid, 256, sizeof(unsigned), st>>>(pf, p.partials);
    else if (p.D <= 64)
        vq_tc_finish_kernel<2><<<grid, 256, sizeof(unsigned), st>>>(pf, p.partials);
    else
        vq_tc_finish_kernel<4><<<grid, 256, sizeof(unsigned), st>>>(pf, p.partials);
    *n_ctas = grid;
    return cudaGetLastError();
}

cudaError_t launch_fwd_tcs(const FwdParams &p, float *tc_scratch, int sm_count, int max_smem, int *n_ctas, int *n_launches,
                           cudaStream_t st, cudaEvent_t ev_begin, cudaEvent_t ev_end, bool image_ready);
// The single-product TF32 filter is the default for 16 < D <= 32 (measured at N = 2^24: (256, 32) 0.975 ms against 1.033,
// (128, 32) 0.825 / 0.848, (64, 32) 0.738 / 0.779; narrower vectors need fewer bf16 MMAs and stay on the three-product
// filter: (256, 16) 0.795 against 0.883 with TF32).  VQB_TF32=0 in the environment selects the bf16 filter everywhere
// (A/B runs), VQB_TF32=1 the TF32 filter for every D <= 32.
static int g_filter_override = -1;        // vqb_debug_set_filter: -1 automatic, 0 bf16 three-product, 1 TF32 single-product
void set_tc_filter(int mode) { g_filter_override = mode; }
bool tc_filter_is_tf32(int d);
static bool tf32_mode_requested(int d) { return tc_filter_is_tf32(d); }
// forced (vqb_debug_set_filter(1) / VQB_TF32=1), not merely chosen by the automatic rule: the tile-stationary kernel
// (vq_fwd_tcs.cu) takes its TF32 variant only then -- measured slower there (profiles/README.md, r02b)
bool tc_filter_forced_tf32()
{
    static const int env = [] { const char *e = getenv("VQB_TF32"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
    return g_filter_override >= 0 ? g_filter_override == 1 : env == 1;
}
bool tc_filter_is_tf32(int d)
{
    static const int env = [] { const char *e = getenv("VQB_TF32"); return !e ? -1 : (e[0] == '1' ? 1 : 0); }();
    const int v = g_filter_override >= 0 ? g_filter_override : env;
    return v < 0 ? d > 16 : v == 1;
}
// CTA pairs are an A/B switch while they are being measured: VQB_PAIR=1 in the environment
static bool pair_mode_requested()
{
    static const bool v = [] { const char *e = getenv("VQB_PAIR"); return e && e[0] == '1'; }();
    return v;
}
static bool wide_tcs_requested()
{
    static const bool v = [] { const char *e = getenv("VQB_WIDE_TCS"); return e && e[0] == '1'; }();
    return v;
}

static unsigned long long *g_trace_buf = nullptr;   // debug only, see vqb_debug_set_tc_trace
void set_tc_trace(unsigned long long *buf) { g_trace_buf = buf; }
unsigned long long *tc_trace_buf() { return g_trace_buf; }
size_t tc_trace_words() { return (size_t)kTraceCtas * kTraceTiles * kTraceEvents; }

cudaError_t launch_fwd_tc(const FwdParams &p, float *tc_scratch, int sm_count, int max_smem, int *n_ctas,
                          int *n_launches, cudaStream_t st, cudaEvent_t ev_begin, cudaEvent_t ev_end, bool image_ready)
{
    using namespace tc;
    // (A/B switch: VQB_WIDE_TCS=1 sends 32 < D <= 64 to the tile-stationary kernel instead of the two-slot schedule below)
    if (p.D > tc::D && p.idx && p.chunk_mode == 0 && wide_tcs_requested())
        return launch_fwd_tcs(p, tc_scratch, sm_count, max_smem, n_ctas, n_launches, st, ev_begin, ev_end, false);
    // (row coordinates of the TMA boxes and the queue entries of the fix-up kernel are 32-bit)
    if (!tc_shape_supported(p.K, p.D) || !p.z.rows_contiguous(p.D) || SMEM_ALLOC > max_smem || p.z.n_rows >= (1ll << 31))
        return cudaErrorNotSupported;
    unsigned char *img = reinterpret_cast<unsigned char *>(tc_scratch);
    CUtensorMap map_z, map_zq;
    if (!make_map(&map_z, p.z.base, p.z.n_rows, p.D))
        return cudaErrorNotSupported;
    if (p.zq) {
        if (!make_map(&map_zq, p.zq, p.z.n_rows, p.D))
            return cudaErrorNotSupported;
    } else {
        map_zq = map_z;
    }
    const int kp = ((p.K + 31) / 32) * 32;
    const bool wide = p.D > tc::D;
    // image_ready: the constant-operand image in tc_scratch was built by an earlier call for this very codebook
    // (the per-CTA queue counters need no reset: every CTA overwrites its own before anyone reads it)
    cudaError_t err = cudaSuccess;
    if (!image_ready) {
        if ((err = cudaMemsetAsync(img + IMG_CONST, 0, sizeof(Consts), st)) != cudaSuccess)
            return err;
        vq_tc_prep_kernel<<<(KMAX + 127) / 128, 128, 0, st>>>(p.E, p.ee, p.K, p.D, kp, img,
                                                              (!wide && p.chunk_mode == 0 && tf32_mode_requested(p.D)) ? 1 : 0);
        if ((err = cudaGetLastError()) != cudaSuccess)
            return err;
    }
    const int64_t tiles = (p.z.n_rows + TILE_M - 1) / TILE_M;
    int grid = (int)(tiles < sm_count ? tiles : sm_count);
    if (grid < 1)
        grid = 1;
    if (grid > WL_CTAS)
        grid = WL_CTAS;
    // CTA pairs (cta_group::2) for the plain pass over more than 128 codes: an even grid of 2-CTA clusters
    const bool tf32 = !wide && p.chunk_mode == 0 && tf32_mode_requested(p.D);
    const bool pair = !tf32 && !wide && p.chunk_mode == 0 && p.K > 128 && grid >= 2 && pair_mode_requested();
    if (pair)
        grid &= ~1;
    auto kern = wide ? vq_fwd_tc_kernel<false, true, false, false>
                : tf32 ? (g_trace_buf ? vq_fwd_tc_kernel<true, false, false, true> : vq_fwd_tc_kernel<false, false, false, true>)
                : g_trace_buf ? (pair ? vq_fwd_tc_kernel<true, false, true, false> : vq_fwd_tc_kernel<true, false, false, false>)
                              : pair ? vq_fwd_tc_kernel<false, false, true, false> : vq_fwd_tc_kernel<false, false, false, false>;
    const int smem_alloc = tf32 ? SMEM_ALLOC_TF32 : SMEM_ALLOC;
    if (smem_alloc > max_smem)
        return cudaErrorNotSupported;
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_alloc);
    if (err != cudaSuccess)
        return err;
    *n_ctas = grid * (1 + FIX_SPLIT);    // partials [0, grid): main kernel, then one per fix-up CTA
    if (ev_begin)
        cudaEventRecord(ev_begin, st);
    if (pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = SMEM_ALLOC;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const unsigned char *img_c = img;
        unsigned long long *trace_buf = g_trace_buf;
        err = cudaLaunchKernelEx(&cfg, kern, p, img_c, map_z, map_zq, kp, trace_buf);
    } else {
        kern<<<grid, THREADS, smem_alloc, st>>>(p, img, map_z, map_zq, kp, wide ? nullptr : g_trace_buf);
        err = cudaGetLastError();
    }
    if (err != cudaSuccess)
        return err;
    if (wide || p.chunk_mode != 0) {   // (the plain D <= 32 pass decides its queued vectors itself, see the kernel's tail)
        vq_tc_fixup_kernel<<<grid * FIX_SPLIT, 256, 0, st>>>(p, img, p.partials + grid);
        err = cudaGetLastError();
        *n_launches = image_ready ? 2 : 3;
    } else {
        *n_ctas = grid;
        *n_launches = image_ready ? 1 : 2;
    }
    if (ev_end)
        cudaEventRecord(ev_end, st);
    return err;
}

// ---------------------------------------------------------------------------------------
// K > 256: one pass of the kernels above per 256-code chunk (codebook operands re-prepared per pass), each
// merging the chunk's exact winner into the running best that lives in the caller's idx buffer, then the
// finish kernel.  Exactness carries over: every chunk winner is the oracle's argmin within its chunk, and
// the chunks are compared by oracle-order distances with the lower chunk winning ties.
// ---------------------------------------------------------------------------------------
// Shapes beyond the resident-codebook kernel (K > 256, or 64 < D <= 128) run on the tile-stationary kernel of
// vq_fwd_tcs.cu; the pass-per-chunk schedule below is kept for A/B runs (VQB_CHUNK_PASSES=1 in the environment).
bool tcs_shape_supported(int K, int D);
cudaError_t launch_fwd_tcs(const FwdParams &p, float *tc_scratch, int sm_count, int max_smem, int *n_ctas, int *n_launches,
                           cudaStream_t st, cudaEvent_t ev_begin, cudaEvent_t ev_end, bool image_ready);

static bool chunk_passes_requested()
{
    static const bool v = [] { const char *e = getenv("VQB_CHUNK_PASSES"); return e && e[0] == '1'; }();
    return v;
}

bool tc_chunked_supported(int K, int D)
{
    return !tc_shape_supported(K, D) && tcs_shape_supported(K, D);
}

cudaError_t launch_fwd_tc_chunked(const FwdParams &p, float *tc_scratch, int sm_count, int max_smem, int *n_ctas,
                                  int *n_launches, cudaStream_t st, cudaEvent_t ev_begin, cudaEvent_t ev_end)
{
    using namespace tc;
    if (!tc_chunked_supported(p.K, p.D) || !p.idx)
        return cudaErrorNotSupported;
    if (!(chunk_passes_requested() && tc_shape_supported(KMAX, p.D) && p.K <= kFinishMaxK))
        return launch_fwd_tcs(p, tc_scratch, sm_count, max_smem, n_ctas, n_launches, st, ev_begin, ev_end, false);
    if (ev_begin)
        cudaEventRecord(ev_begin, st);
    int launches = 0, pass_ctas = 1;
    const int n_chunks = (p.K + KMAX - 1) / KMAX;
    for (int c = 0; c < n_chunks; ++c) {
        FwdParams pc = p;
        pc.E = p.E + (size_t)c * KMAX * p.D;
        pc.ee = p.ee + c * KMAX;
        pc.K = p.K - c * KMAX < KMAX ? p.K - c * KMAX : KMAX;
        pc.zq = nullptr;
        pc.idx = nullptr;
        pc.need_sq = 0;
        pc.run = reinterpret_cast<unsigned long long *>(p.idx);
        pc.code_base = c * KMAX;
        pc.chunk_mode = c == 0 ? 1 : 2;
        int nl = 0;
        cudaError_t err = launch_fwd_tc(pc, tc_scratch, sm_count, max_smem, &pass_ctas, &nl, st, nullptr, nullptr, false);
        if (err != cudaSuccess)
            return err;
        launches += nl;
    }
    FwdParams pf = p;
    pf.run = reinterpret_cast<unsigned long long *>(p.idx);
    const int64_t groups = (p.z.n_rows + 127) / 128;
    int grid = (int)(groups < (int64_t)sm_count * 6 ? groups : (int64_t)sm_count * 6);
    if (grid < 1)
        grid = 1;
    if (grid > kMaxPartials)
        grid = kMaxPartials;
    auto fin = p.D <= 32 ? vq_tc_finish_kernel<1> : vq_tc_finish_kernel<2>;
    cudaError_t err = cudaFuncSetAttribute(fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(unsigned) * kFinishMaxK));
    if (err != cudaSuccess)
        return err;
    fin<<<grid, 256, sizeof(unsigned) * (size_t)p.K, st>>>(pf, p.partials);
    err = cudaGetLastError();
    if (ev_end)
        cudaEventRecord(ev_end, st);
    // the pass kernels left zeros in partials [0, pass_ctas); the finish kernel owns [0, grid)
    *n_ctas = pass_ctas > grid ? pass_ctas : grid;
    *n_launches = launches + 1;
    return err;
}

}  // namespace vqb
