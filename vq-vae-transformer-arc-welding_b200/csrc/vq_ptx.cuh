// vq_ptx.cuh -- sm_100a PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, tcgen05.mma / TMEM,
// UMMA shared-memory and instruction descriptors).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace vqb {
namespace tc {

constexpr int TILE_M = 128;               // rows of a tcgen05 tile (= TMEM lanes)

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or
// ~`hint_ns` elapse, so a long wait costs a handful of instructions instead of a polling storm.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity, uint32_t hint_ns)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(hint_ns)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must trap instead of hanging the GPU.
template <int SLEEP_NS = 32>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    // Polling costs issue slots and power (the kernel runs at the board's power cap): sleep between
    // polls, and let one warp poll on behalf of its group (group_wait below).
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity, 1000u)) {
        if (SLEEP_NS > 0)
            __nanosleep(SLEEP_NS);
        if (++spins > (1u << 22))       // seconds: a pipeline bug traps instead of hanging the GPU
            __trap();
    }
}
// latency-critical waits (accumulator ready, operands ready): plain try_wait loop, the hardware
// suspends the thread inside try_wait and wakes it on completion
__device__ __forceinline__ void mbar_wait_tight(uint32_t bar, uint32_t parity)
{
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok)
            break;
        if (++spins > (1u << 26))
            __trap();
    }
}
// One warp (`leader`) polls the mbarrier, the other warps of the 128-thread group block in hardware
// on a named barrier until it has seen the phase complete.
template <int SLEEP_NS>
__device__ __forceinline__ void group_wait(bool leader, uint32_t bar, uint32_t parity, int bar_id)
{
    if (leader)
        mbar_wait<SLEEP_NS>(bar, parity);
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
}
// One lane of a converged warp.  tcgen05.mma / tcgen05.commit are issued under this predicate from WARP-UNIFORM code: the
// compiler then keeps the descriptors in uniform registers and emits one UTCHMMA per MMA.  Under `if (lane == 0)` it
// cannot know that a single thread is active and wraps every MMA in an election loop (ELECT / R2UR / BRA.U.ANY, ~35
// instructions per MMA): the single issuing thread then needs ~600 clocks per four MMAs and starves the tensor pipe
// (profiles/README.md, round 2: the encoder chain's issuer was busy on every sample).
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2)
{   // signed coordinates: parts of the box outside the tensor (e.g. c1 = -1) arrive as zeros
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32: operands are fp32 words in shared memory of which the tensor core uses the upper 19 bits (sign, exponent, 10
// mantissa bits); K = 8 per instruction (32 bytes of a K-major row), same tensor-pipe time as a K = 16 bf16 MMA
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::tf32, A = B = TF32 (format 2), D = F32, both K-major, M = 128, N = n
__host__ __device__ inline uint32_t idesc_tf32(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
__device__ __forceinline__ float round_tf32(float x)
{   // round to nearest (ties away) to 10 mantissa bits; the result is an fp32 whose low 13 bits are zero
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the SMs of one TPC run ONE tcgen05.mma of M = 256: each supplies
// its own 128 rows of A and HALF of the B operand from its own shared memory (same offsets in both), and receives its own
// 128 accumulator rows in its own TMEM.  Only the leader (cluster rank 0) issues; completion is multicast to both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster (works for the own CTA too).
// Default semantics (.release at CTA scope), as CUTLASS's ClusterBarrier::arrive does: what the waiter consumes are
// shared-memory operands of the ARRIVING CTA's own SM (ordered by fence.proxy.async) and tensor-memory reads ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync.  (A .release.cluster arrive was measured first: ~1000 clocks per
// arrival, it drains the thread's memory operations at cluster scope.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar, uint32_t cta)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar)
{   // arrives on the barrier at this offset in BOTH CTAs of the pair once all MMAs issued so far have completed
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// kind::f16, A = B = BF16, D = F32, both K-major, M = m (256 for a CTA pair), N = n
__host__ __device__ inline uint32_t idesc_bf16_mn(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&a)[16], uint32_t (&b)[16])
{   // tcgen05.wait::ld with both freshly loaded register groups threaded through (no use may move above the wait)
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
                 :
                 : "memory");
    asm volatile(""
                 : "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]), "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]),
                   "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]), "+r"(b[14]), "+r"(b[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld_fence16(uint32_t (&a)[16])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]),
                   "+r"(a[8]), "+r"(a[9]), "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float min3(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// MUFU.SQRT: 1 instruction, ~2^-22 relative error; callers scale the result up a little because it
// only feeds upper bounds
__device__ __forceinline__ float sqrt_approx(float x)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// K-major operand descriptors (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr)
{   // 8-row x 128-byte swizzle atoms, 1024 bytes apart along M/N
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_sw32(uint32_t saddr)
{   // 8-row x 32-byte swizzle atoms, 256 bytes apart along M/N
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(256 >> 4) << 32) | (1ull << 46) | (6ull << 61);
}
// kind::f16, A = B = BF16, D = F32, both K-major, M = 128, N = n
__host__ __device__ inline uint32_t idesc_bf16(int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

// tcgen05.wait::ld, with the freshly loaded registers threaded through the statement so that
// no use of them can be scheduled before the wait
__device__ __forceinline__ void tmem_wait_ld_fence(uint32_t (&v)[32])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
    asm volatile(""
                 : "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}


// ---- host side: 2-D tiled tensor maps (the driver entry point is fetched once through the runtime, no -lcuda) ----
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                           const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TensorMapEncodeTiledFn tensor_map_encoder()
{
    static TensorMapEncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<TensorMapEncodeTiledFn>(ptr);
    }
    return fn;
}

// (rows, cols) row-major tensor of `elem_bytes`-wide elements under boxes of box_rows x box_cols; out-of-range parts of a
// box are zero-filled on load and clipped on store
inline bool make_tensor_map_2d(CUtensorMap *map, CUtensorMapDataType dtype, int elem_bytes, const void *base, int64_t rows,
                               int64_t cols, int box_rows, int box_cols, CUtensorMapSwizzle swizzle,
                               CUtensorMapL2promotion l2)
{
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc)
        return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * (cuuint64_t)elem_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, dtype, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// (n2, n1, cols) row-major tensor of `elem_bytes`-wide elements under boxes of box2 x box1 x box_cols
inline bool make_tensor_map_3d(CUtensorMap *map, CUtensorMapDataType dtype, int elem_bytes, const void *base, int64_t n2,
                               int64_t n1, int64_t cols, int box2, int box1, int box_cols, CUtensorMapSwizzle swizzle,
                               CUtensorMapL2promotion l2)
{
    TensorMapEncodeTiledFn enc = tensor_map_encoder();
    if (!enc)
        return false;
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)n1, (cuuint64_t)n2};
    const cuuint64_t strides[2] = {(cuuint64_t)cols * (cuuint64_t)elem_bytes, (cuuint64_t)cols * (cuuint64_t)n1 * (cuuint64_t)elem_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box1, (cuuint32_t)box2};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, dtype, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, l2,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace vqb
