// vq_bwd.cu -- fused straight-through backward.
//
// The reference gets its backward from autograd over model/vector_quantizer.py:103-111:
// a (K x N) @ (N x D) GEMM through the dense one-hot plus ~8 elementwise passes.  Closed
// form, M = N*D:
//     grad_z[i]  = g_zq[i] + g_loss * 2 * (z_i - E[idx_i]) / M
//     grad_E[c]  = g_loss * beta * 2 / M * sum_{i: idx_i = c} (E[c] - z_i)
// One streaming kernel reads g_zq, z and idx once, writes grad_z once (HBM-bound,
// 12*D + 8 bytes per vector) and accumulates the per-code residual sums in shared memory;
// a K*D-thread kernel scales them into grad_E.
#include <cuda.h>
#include <cstdio>

#include "vq_common.cuh"
#include "vq_ptx.cuh"

namespace vqb {

constexpr int kBwdThreads = 256;
constexpr int kBwdRowsPerTile = 64;

// SMEM_ACC: residual sums privatised per CTA in shared memory (K*D floats), flushed once;
// otherwise straight to global atomics (large codebooks).
template <bool SMEM_ACC>
__global__ void __launch_bounds__(kBwdThreads) vq_bwd_kernel(const float *__restrict__ g_zq,
                                                             const float *__restrict__ g_loss, const ZView z,
                                                             const int64_t *__restrict__ idx,
                                                             const float *__restrict__ E, int K, int D,
                                                             float *__restrict__ grad_z, float *acc_global)
{
    extern __shared__ float acc_s[];  // [K*D] when SMEM_ACC
    const int tid = threadIdx.x;
    if (SMEM_ACC) {
        for (int t = tid; t < K * D; t += kBwdThreads)
            acc_s[t] = 0.0f;
        __syncthreads();
    }
    const float gl = g_loss ? __ldg(g_loss) : 0.0f;
    // cz = g_loss * 2 / M, evaluated in double and rounded once
    const float cz = (float)((double)gl * 2.0 / ((double)z.n_rows * (double)D));
    const int64_t n_tiles = (z.n_rows + kBwdRowsPerTile - 1) / kBwdRowsPerTile;
    const int per_tile = kBwdRowsPerTile * D;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kBwdRowsPerTile;
        for (int t = tid; t < per_tile; t += kBwdThreads) {
            const int r = t / D, j = t - r * D;
            const int64_t row = row0 + r;
            if (row >= z.n_rows)
                break;
            const int64_t code = idx[row];
            const float zv = z.row(row)[j * z.s_d];
            const float diff = __fsub_rn(__ldg(E + code * D + j), zv);   // E[idx] - z
            if (grad_z) {
                const float g = g_zq ? g_zq[row * D + j] : 0.0f;
                grad_z[row * D + j] = fmaf(-cz, diff, g);
            }
            if (acc_global) {
                if (SMEM_ACC)
                    atomicAdd(acc_s + code * D + j, diff);
                else
                    atomicAdd(acc_global + code * D + j, diff);
            }
        }
    }
    if (SMEM_ACC && acc_global) {
        __syncthreads();
        for (int t = tid; t < K * D; t += kBwdThreads) {
            const float v = acc_s[t];
            if (v != 0.0f)
                atomicAdd(acc_global + t, v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// D = 32, K <= kBwd32MaxK: the streaming kernel.  Shared-memory float atomics are CAS loops on
// sm_100a (5.4 ms at N = 2^24 for the kernel above), so this one adds without atomics:
//   phase 1  eight threads per vector, 16-byte loads of z / g_zq / E[idx], grad_z stored, the
//            residual E[idx] - z parked in a 256-vector shared-memory tile next to its code;
//   phase 2  warp w owns the codes c with c % 8 == w: it finds its vectors in the tile with
//            ballots and adds their residual rows (lane = component) into the CTA's private
//            K x 32 accumulator with plain loads/stores -- one owner per code, no conflicts;
//   flush    one global atomicAdd per accumulator entry per CTA.
// Measured (N = 2^24, K = 256): 1.57 ms with both outputs, 1.25 ms for grad_z alone (phase 1), 1.35 ms for grad_E alone.
// A warp-specialised variant (8 streaming + 8 accumulating warps per CTA over double-buffered tiles, 2 CTAs/SM) was
// slower, 1.93 ms: phase 1 needs all 24 warps of an SM streaming to keep enough loads in flight.
// ---------------------------------------------------------------------------------------
#ifndef VQB_BWD_ROWS      // vectors per shared-memory tile; 128 and 512 were measured 15-18 % slower than 256
#define VQB_BWD_ROWS 256
#endif
constexpr int kBwd32Rows = VQB_BWD_ROWS;
constexpr int kBwd32MaxK = 1024;

// Phase 2 of both D = 32 kernels: the calling warp adds the residual rows (32 floats, lane = component) of the
// vectors in `mask` (lanes of a 32-vector chunk whose codes `c` this warp owns) into the accumulator rows of their
// codes, one vector after the other.  (A four-vectors-per-step variant that loads the rows together and chains equal
// codes in registers was measured: bit-identical, 2.4x the instructions, no faster -- the phase is bound by the
// warps' dependent-issue latency, not by the shared-memory round trips.)
__device__ __forceinline__ void accumulate_rows(unsigned mask, int c, const float *diff_chunk, float *acc_s, int lane)
{
    constexpr int D = 32;
    while (mask) {
        const int l = __ffs(mask) - 1;
        mask &= mask - 1;
        const int cc = __shfl_sync(0xffffffffu, c, l);
        float *a = acc_s + cc * D + lane;
        *a = *a + diff_chunk[l * D + lane];
    }
}

template <bool CONTIG>
__global__ void __launch_bounds__(256) vq_bwd32_kernel(const float *__restrict__ g_zq, const float *__restrict__ g_loss,
                                                       const ZView z, const int64_t *__restrict__ idx,
                                                       const float *__restrict__ E, int K, float *__restrict__ grad_z,
                                                       float *acc_global)
{
    constexpr int D = 32;
    extern __shared__ __align__(16) float bsm[];
    float *diff_s = bsm;                                   // [256][32]
    float *acc_s = bsm + kBwd32Rows * D;                   // [K][32]
    int *code_s = reinterpret_cast<int *>(acc_s + K * D);  // [256]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sub = tid & 7, rloc = tid >> 3;
    if (acc_global) {
        for (int t = tid; t < K * D; t += 256)
            acc_s[t] = 0.0f;
    }
    const float gl = g_loss ? __ldg(g_loss) : 0.0f;
    const float cz = (float)((double)gl * 2.0 / ((double)z.n_rows * (double)D));
    const int64_t n_tiles = (z.n_rows + kBwd32Rows - 1) / kBwd32Rows;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row0 = tile * kBwd32Rows;
        // loads of four vectors are issued together before anything depends on them (memory-level parallelism:
        // a pass that waits for its own idx -> E[idx] chain costs two round trips)
#pragma unroll 1
        for (int half = 0; half < kBwd32Rows / 128; ++half) {
            int64_t c64[4];
            float4 zv[4], g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t row = row0 + (half * 4 + u) * 32 + rloc;
                const bool valid = row < z.n_rows;
                c64[u] = valid ? __ldg(idx + row) : -1;
                zv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (valid) {
                    if (CONTIG) {
                        zv[u] = __ldg(reinterpret_cast<const float4 *>(z.base + row * D) + sub);
                    } else {
                        const float *zr = z.row(row) + (int64_t)(4 * sub) * z.s_d;
                        zv[u] = make_float4(zr[0], zr[z.s_d], zr[2 * z.s_d], zr[3 * z.s_d]);
                    }
                    if (g_zq)
                        g[u] = __ldg(reinterpret_cast<const float4 *>(g_zq + row * D) + sub);
                }
            }
            float4 e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool in = c64[u] >= 0 && c64[u] < K;
                e[u] = in ? __ldg(reinterpret_cast<const float4 *>(E + (size_t)c64[u] * D) + sub) : zv[u];   // diff = 0
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = (half * 4 + u) * 32 + rloc;
                const int64_t row = row0 + r;
                const bool in = c64[u] >= 0 && c64[u] < K;
                const float4 diff = make_float4(__fsub_rn(e[u].x, zv[u].x), __fsub_rn(e[u].y, zv[u].y),
                                                __fsub_rn(e[u].z, zv[u].z), __fsub_rn(e[u].w, zv[u].w));
                if (grad_z && row < z.n_rows) {
                    const float4 o = make_float4(fmaf(-cz, diff.x, g[u].x), fmaf(-cz, diff.y, g[u].y),
                                                 fmaf(-cz, diff.z, g[u].z), fmaf(-cz, diff.w, g[u].w));
                    __stcs(reinterpret_cast<float4 *>(grad_z + row * D) + sub, o);
                }
                reinterpret_cast<float4 *>(diff_s + r * D)[sub] = diff;
                if (sub == 0)
                    code_s[r] = in ? (int)c64[u] : -1;
            }
        }
        __syncthreads();
        if (acc_global) {
#pragma unroll 1
            for (int chunk = 0; chunk < kBwd32Rows / 32; ++chunk) {
                const int c = code_s[chunk * 32 + lane];
                unsigned mask = __ballot_sync(0xffffffffu, c >= 0 && (c & 7) == warp);
                accumulate_rows(mask, c, diff_s + chunk * 32 * D, acc_s, lane);
            }
        }
        __syncthreads();
    }
    if (acc_global) {
        for (int t = tid; t < K * D; t += 256) {
            const float v = acc_s[t];
            if (v != 0.0f)
                atomicAdd(acc_global + t, v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// D = 32, K <= 256, contiguous z / g_zq, all outputs wanted: the TMA-ring kernel (persistent, one CTA per SM).
// The kernel above keeps loads in flight only while a CTA is in its phase 1; here the data movement is decoupled
// from both phases the way the forward kernel does it:
//   warp 0      ring owner (one thread): TMA loads of a 128-vector tile of z and of g_zq plus a 1-D bulk copy of the
//               tile's ids into a 4-deep ring (2 x 16 KB + 1 KB per slot), refill of a slot as soon as phase 2 is
//               done with it
//   warps 1-8   phase 1: eight threads per vector; residual E[idx] - z written IN PLACE over the z tile,
//               grad_z = g_zq - cz * residual stored to global memory (16 bytes per thread, 512 contiguous bytes
//               per warp; the codebook is resident in shared memory)
//   warps 9-24  phase 2: warp w owns the codes c with c % 16 == w and adds the residual rows of its vectors
//               into the CTA's K x 32 accumulator (accumulate_rows), while phase 1 works on the next slot
//   flush       one global atomicAdd per accumulator entry per CTA.
// Measured at N = 2^24, K = 256 (tools/bwd_time.py, tools/bwd_trace.py; profiles/README.md): 1.32 ms = 77 % of the HBM
// roofline (the kernel above: 1.57 ms).  What the clock64 timeline of a -DBW_TRACE=1 build shows: phase 1 takes ~2800
// clocks per tile and phase 2 ~2200 per warp against a tile period of ~2700 -- both phases run one dependent
// instruction per ~12 clocks and warp, so it is their instruction COUNT per tile that bounds the kernel now, not the
// ring (3 stages 1.42 ms, 4 stages 1.32 ms, 5 stages 1.55 ms: no L1 left).  Measured and rejected on the way:
//   * grad_z through a TMA store out of the slot (refill waits for the store to have read it: 1.55 ms);
//   * ids fetched by the phase-1 threads one tile ahead with ordinary loads (the __syncwarp behind the barrier wait
//     then waits for those loads as well, 3.5 us under load: 1.51 ms) -- they travel through the ring now;
//   * phase 2 sums in registers behind a warp-uniform switch over the code (divergence-safe code: 1.98 ms);
//   * two / four vectors per phase-2 step with equal codes chained in registers (1.39 / 1.59 ms: more instructions);
//   * phase 1 handing every vector to its owner's list with shared-memory atomics instead of phase 2 scanning the
//     tile's codes (2.8 ms); a bounds-check-free phase 1 for full tiles (phase 1 2800 -> 2100 clocks, kernel 1.34 ms:
//     the faster phase 1 only takes issue slots from phase 2);
//   * codes and owners as bytes (two loads per tile instead of four + compares) with 16 / 20 / 23 phase-2 warps:
//     1.50 / 1.77 / 1.84 ms -- the byte stores of phase 1 serialise and more warps only add contention.
// ---------------------------------------------------------------------------------------
#ifdef BW_TRACE   // debug build (tools/ab_build.py trace:-DBW_TRACE=1): clock64 stamps of CTA 0, tiles 40..103
__device__ long long bw_trace_buf[6 * 64];
#define BW_STAMP(ev, it) do { if (blockIdx.x == 0 && (it) >= 40 && (it) < 104) bw_trace_buf[(ev) * 64 + (it) - 40] = clock64(); } while (0)
#else
#define BW_STAMP(ev, it) do {} while (0)
#endif
namespace bw {
// Experiment knobs (tools/ab_build.py name:-DBW_STAGES=3 ...; tools/ab_bwd.py): ring depth, and two ablations whose results
// are wrong by construction -- BW_SKIP_ACC (phase 2 arrives without accumulating), BW_SKIP_STORE (no grad_z stores).
#ifndef BW_STAGES
#define BW_STAGES 4
#endif
constexpr int D = 32, ROWS = 128, STAGES = BW_STAGES, KMAX = 256;
constexpr int TILE_BYTES = ROWS * D * 4;                       // 16 KB
constexpr int OFF_Z = 0;                                       // STAGES x 16 KB   z tiles -> residual tiles
constexpr int OFF_G = OFF_Z + STAGES * TILE_BYTES;             // STAGES x 16 KB   g_zq tiles -> grad_z tiles
constexpr int OFF_E = OFF_G + STAGES * TILE_BYTES;             // 32 KB            codebook rows
constexpr int OFF_ACC = OFF_E + KMAX * D * 4;                  // 32 KB            per-code residual sums
constexpr int OFF_CODE = OFF_ACC + KMAX * D * 4;               // STAGES x 512     code of every vector of the tile (-1: none)
constexpr int OFF_IDX = OFF_CODE + STAGES * ROWS * 4;          // STAGES x 1 KB    int64 ids of the tile (bulk copy)
constexpr int OFF_BARS = OFF_IDX + STAGES * ROWS * 8;
constexpr int SMEM_BYTES = OFF_BARS + 256;
constexpr int ACC_WARPS = 16;                                  // phase-2 warps: warp w owns the codes c % 16 == w
constexpr int THREADS = 32 + 256 + 32 * ACC_WARPS;
}  // namespace bw

__global__ void __launch_bounds__(bw::THREADS, 1)
vq_bwd32_tma_kernel(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_g,
                    float *__restrict__ grad_z, const float *__restrict__ g_loss, int64_t n_rows,
                    const int64_t *__restrict__ idx, const float *__restrict__ E, int K, float *acc_global)
{
    using namespace tc;
    using namespace bw;
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    enum { FULL = 0, COMPUTED = FULL + STAGES, ACC_DONE = COMPUTED + STAGES, N_BARS = ACC_DONE + STAGES };
    static_assert(8 * N_BARS <= 256, "barrier area");
    auto bar = [&](int i) { return sbase + OFF_BARS + 8 * i; };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *acc_s = reinterpret_cast<float *>(smem + OFF_ACC);
    float *e_s = reinterpret_cast<float *>(smem + OFF_E);

    const int64_t n_tiles = (n_rows + ROWS - 1) / ROWS;
    const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar(FULL + s), 1);
            mbar_init(bar(COMPUTED + s), 8);
            mbar_init(bar(ACC_DONE + s), ACC_WARPS);
        }
        fence_barrier_init();
    }
    for (int t = tid; t < KMAX * D; t += THREADS) {
        acc_s[t] = 0.0f;
        e_s[t] = t < K * D ? __ldg(E + t) : 0.0f;
    }
    __syncthreads();
    const float gl = g_loss ? __ldg(g_loss) : 0.0f;
    const float cz = (float)((double)gl * 2.0 / ((double)n_rows * (double)D));

    if (warp == 0) {
        // ================= ring owner =================
        if (lane == 0) {
            auto load_tile = [&](int it) {
                const int s = it % STAGES;
                const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
                BW_STAMP(0, it);
                // ids: 1-D bulk copy of the tile's valid rows, an even number of them (16-byte granules); the odd last
                // row of the tensor, if any, is fetched by its phase-1 thread
                const int64_t left = n_rows - tile * ROWS;
                const uint32_t id_bytes = (uint32_t)(left >= ROWS ? ROWS : (left & ~int64_t(1))) * 8u;
                mbar_expect_tx(bar(FULL + s), 2 * TILE_BYTES + id_bytes);
                tma_load_2d(sbase + OFF_Z + s * TILE_BYTES, &map_z, bar(FULL + s), 0, (int)(tile * ROWS));
                tma_load_2d(sbase + OFF_G + s * TILE_BYTES, &map_g, bar(FULL + s), 0, (int)(tile * ROWS));
                if (id_bytes)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(sbase + OFF_IDX + s * ROWS * 8), "l"(idx + tile * ROWS), "r"(id_bytes), "r"(bar(FULL + s))
                                 : "memory");
            };
            for (int it = 0; it < my_tiles && it < STAGES; ++it)
                load_tile(it);
            // (grad_z leaves through plain 16-byte stores of the phase-1 threads -- a warp writes 512 contiguous bytes.
            // A TMA store out of the slot was measured first: with the slot's refill waiting for the store to have
            // read it, one store per SM is in flight at a time and the kernel stays at 1.55 ms.)
            for (int it = 0; it + STAGES < my_tiles; ++it) {
                const int s = it % STAGES;
                mbar_wait<64>(bar(ACC_DONE + s), (uint32_t)((it / STAGES) & 1));   // residual tile accumulated: slot free
                BW_STAMP(5, it);
                load_tile(it + STAGES);
            }
        }
    } else if (warp <= 8) {
        // ================= phase 1: residual + grad_z, eight threads per vector =================
        const int t = tid - 32, sub = t & 7, rloc = t >> 3;       // rows rloc + 32 u
        // (The ids come through the ring as well.  Fetching them with ordinary loads one tile ahead was measured: the
        // __syncwarp behind the barrier wait then waits for those loads too -- 3.5 us under load, twice the tile period.)
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % STAGES;
            if (lane == 0)                                 // one lane polls, __syncwarp orders the rest behind it
                mbar_wait<32>(bar(FULL + s), (uint32_t)((it / STAGES) & 1));
            if (tid == 32) BW_STAMP(1, it);
            __syncwarp();
            float4 *zs = reinterpret_cast<float4 *>(smem + OFF_Z + s * TILE_BYTES);
            const float4 *gs = reinterpret_cast<const float4 *>(smem + OFF_G + s * TILE_BYTES);
            const int64_t row0 = (blockIdx.x + (int64_t)it * gridDim.x) * ROWS;
            const int64_t *idx_s = reinterpret_cast<const int64_t *>(smem + OFF_IDX + s * ROWS * 8);
            const int64_t odd_last = (n_rows & 1) ? n_rows - 1 : -1;
            int *code_s = reinterpret_cast<int *>(smem + OFF_CODE + s * ROWS * 4);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = 32 * u + rloc;
                const int64_t row = row0 + r;
                const int64_t c64 = row >= n_rows ? -1 : row == odd_last ? __ldg(idx + row) : idx_s[r];
                const int code = (c64 >= 0 && c64 < K) ? (int)c64 : -1;
                const float4 zv = zs[r * 8 + sub];
                const float4 g = gs[r * 8 + sub];
                const float4 e = code >= 0 ? reinterpret_cast<const float4 *>(e_s)[code * 8 + sub] : zv;   // diff = 0
                const float4 diff = make_float4(__fsub_rn(e.x, zv.x), __fsub_rn(e.y, zv.y), __fsub_rn(e.z, zv.z),
                                                __fsub_rn(e.w, zv.w));
                zs[r * 8 + sub] = diff;
#ifndef BW_SKIP_STORE
                if (row < n_rows)
#else
                if (row < n_rows && cz == 123.0f)
#endif
                    __stcs(reinterpret_cast<float4 *>(grad_z + row * D) + sub,
                           make_float4(fmaf(-cz, diff.x, g.x), fmaf(-cz, diff.y, g.y), fmaf(-cz, diff.z, g.z),
                                       fmaf(-cz, diff.w, g.w)));
                if (sub == 0)
                    code_s[r] = code;
            }
            __syncwarp();
            if (tid == 32) BW_STAMP(2, it);
            if (lane == 0)
                mbar_arrive(bar(COMPUTED + s));
        }
    } else {
        // ================= phase 2: per-code residual sums, owner warp per code =================
        // (Keeping the warp's 16 sums in registers behind a warp-uniform switch over the code was measured: the
        // compiler emits divergence-safe code for the switch, 1.98 ms against 1.32 ms for the shared-memory rows.)
        const int w = warp - 9;
        for (int it = 0; it < my_tiles; ++it) {
            const int s = it % STAGES;
            if (lane == 0)
                mbar_wait<32>(bar(COMPUTED + s), (uint32_t)((it / STAGES) & 1));
            if (w == 0 && lane == 0) BW_STAMP(3, it);
            __syncwarp();
            const float *diff_s = reinterpret_cast<const float *>(smem + OFF_Z + s * TILE_BYTES);
            const int *code_s = reinterpret_cast<const int *>(smem + OFF_CODE + s * ROWS * 4);
#pragma unroll 1
            for (int chunk = 0; chunk < ROWS / 32; ++chunk) {
                const int c = code_s[chunk * 32 + lane];
                const unsigned mask = __ballot_sync(0xffffffffu, c >= 0 && (c & (ACC_WARPS - 1)) == w);
#ifndef BW_SKIP_ACC
                accumulate_rows(mask, c, diff_s + chunk * 32 * D, acc_s, lane);
#endif
            }
            __syncwarp();
            if (w == 0 && lane == 0) BW_STAMP(4, it);
            if (lane == 0)
                mbar_arrive(bar(ACC_DONE + s));
        }
    }
    __syncthreads();
    for (int t = tid; t < K * D; t += THREADS) {
        const float v = acc_s[t];
        if (v != 0.0f)
            atomicAdd(acc_global + t, v);
    }
}

namespace {

// (N, 32) fp32 row-major tensor under 128-row boxes, no swizzle: eight threads per row read / write 16-byte chunks
bool bw_make_map(CUtensorMap *map, const float *base, int64_t n_rows)
{
    return tc::make_tensor_map_2d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, n_rows, bw::D, bw::ROWS, bw::D,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B);
}

}  // namespace

__global__ void vq_bwd_scale_kernel(float *grad_E, int n, const float *__restrict__ g_loss, float beta,
                                    int64_t n_rows, int D)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    const float gl = g_loss ? __ldg(g_loss) : 0.0f;
    const float ce = (float)((double)gl * (double)beta * 2.0 / ((double)n_rows * (double)D));
    grad_E[t] = ce * grad_E[t];
}

cudaError_t launch_bwd(const float *g_zq, const float *g_loss, const ZView &z, const int64_t *idx,
                       const float *E, int K, int D, float beta, float *grad_z, float *grad_E,
                       int sm_count, int max_smem, cudaStream_t st)
{
    cudaError_t err;
    if (grad_E) {
        err = cudaMemsetAsync(grad_E, 0, sizeof(float) * (size_t)K * D, st);
        if (err != cudaSuccess)
            return err;
    }
    const size_t smem32 = sizeof(float) * ((size_t)kBwd32Rows * 32 + (size_t)K * 32) + sizeof(int) * kBwd32Rows;
    CUtensorMap map_z, map_g;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(z.base) | reinterpret_cast<uintptr_t>(g_zq) |
                             reinterpret_cast<uintptr_t>(grad_z) | reinterpret_cast<uintptr_t>(idx)) & 15) == 0;
    if (z.n_rows >= bw::ROWS && z.n_rows < (int64_t(1) << 31) && D == 32 && K <= bw::KMAX && g_zq && grad_z && grad_E &&
        aligned16 && z.rows_contiguous(32) && bw::SMEM_BYTES <= max_smem && bw_make_map(&map_z, z.base, z.n_rows) &&
        bw_make_map(&map_g, g_zq, z.n_rows)) {
        const int64_t tiles = (z.n_rows + bw::ROWS - 1) / bw::ROWS;
        const int grid = (int)(tiles < sm_count ? tiles : sm_count);
        err = cudaFuncSetAttribute(vq_bwd32_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bw::SMEM_BYTES);
        if (err != cudaSuccess)
            return err;
        vq_bwd32_tma_kernel<<<grid, bw::THREADS, bw::SMEM_BYTES, st>>>(map_z, map_g, grad_z, g_loss, z.n_rows, idx, E, K,
                                                                        grad_E);
        err = cudaGetLastError();
        if (err != cudaSuccess)
            return err;
    } else if (z.n_rows > 0 && D == 32 && K <= kBwd32MaxK && smem32 <= (size_t)max_smem) {
        const int64_t tiles = (z.n_rows + kBwd32Rows - 1) / kBwd32Rows;
        const int per_sm = (int)((size_t)max_smem / (smem32 + 1024));
        const int64_t cap = (int64_t)sm_count * (per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm);
        const int grid = (int)(tiles < cap ? tiles : cap);
        const bool contig = z.rows_contiguous(32) && (reinterpret_cast<uintptr_t>(z.base) & 15) == 0;
        auto kern = contig ? vq_bwd32_kernel<true> : vq_bwd32_kernel<false>;
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem32);
        if (err != cudaSuccess)
            return err;
        kern<<<grid, 256, smem32, st>>>(g_zq, g_loss, z, idx, E, K, grad_z, grad_E);
        err = cudaGetLastError();
        if (err != cudaSuccess)
            return err;
    } else if (z.n_rows > 0) {
        const int64_t tiles = (z.n_rows + kBwdRowsPerTile - 1) / kBwdRowsPerTile;
        const size_t smem = sizeof(float) * (size_t)K * D;
        const bool use_smem = smem <= (size_t)max_smem / 2 - 1024;
        const int per_sm = use_smem ? (smem > 48 * 1024 ? 2 : 4) : 8;
        int grid = (int)(tiles < (int64_t)sm_count * per_sm ? tiles : (int64_t)sm_count * per_sm);
        if (use_smem) {
            err = cudaFuncSetAttribute(vq_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (err != cudaSuccess)
                return err;
            vq_bwd_kernel<true><<<grid, kBwdThreads, smem, st>>>(g_zq, g_loss, z, idx, E, K, D, grad_z, grad_E);
        } else {
            vq_bwd_kernel<false><<<grid, kBwdThreads, 0, st>>>(g_zq, g_loss, z, idx, E, K, D, grad_z, grad_E);
        }
        err = cudaGetLastError();
        if (err != cudaSuccess)
            return err;
    }
    if (grad_E) {
        const int n = K * D;
        vq_bwd_scale_kernel<<<(n + 255) / 256, 256, 0, st>>>(grad_E, n, g_loss, beta, z.n_rows, D);
        return cudaGetLastError();
    }
    return cudaSuccess;
}

}  // namespace vqb

#ifdef BW_TRACE
extern "C" int vqb_debug_bw_trace(long long *out)
{
    return (int)cudaMemcpyFromSymbol(out, vqb::bw_trace_buf, sizeof(long long) * 6 * 64);
}
#endif
