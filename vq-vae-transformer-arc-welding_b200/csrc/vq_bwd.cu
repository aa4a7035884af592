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
#include "vq_common.cuh"

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
                while (mask) {
                    const int l = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int cc = __shfl_sync(0xffffffffu, c, l);
                    float *a = acc_s + cc * D + lane;
                    *a = *a + diff_s[(chunk * 32 + l) * D + lane];
                }
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
    if (z.n_rows > 0 && D == 32 && K <= kBwd32MaxK && smem32 <= (size_t)max_smem) {
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
