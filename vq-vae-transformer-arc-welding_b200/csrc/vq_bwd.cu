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
    if (z.n_rows > 0) {
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
