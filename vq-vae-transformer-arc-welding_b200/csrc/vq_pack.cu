// vq_pack.cu -- packing copy for the reference encoder's output layout.
//
// SepCNNBlock.forward (model/vq_vae_patch_embedd.py:83-91) returns z_e as a permuted VIEW: logical (B, T, D) over a
// physical (B, D, T) tensor, strides (D*T, 1, T).  The reference's own reshape (model/vector_quantizer.py:88) turns that
// into contiguous rows with a generic strided copy; the tcgen05 / TMA kernels here need the same contiguous rows.  This
// kernel is that copy as a per-cycle D x T -> T x D transpose through shared memory: every global access is a run of
// consecutive floats (the stock strided copy moves the same 8*D bytes per vector at about a third of the rate).
#include "vq_common.cuh"

namespace vqb {

// src: element (b, t, j) at src[b * s_outer + j * T + t]  (block of D x T floats per b, T contiguous)
// dst: (B * T, D) contiguous rows
// VEC: 16-byte global accesses (T and D multiples of 4, 16-byte aligned blocks), else 4-byte ones
template <bool VEC>
__global__ void __launch_bounds__(256) vq_pack_rows_kernel(const float *__restrict__ src, float *__restrict__ dst,
                                                           int64_t n_outer, int T, int D, int64_t s_outer)
{
    extern __shared__ float tile[];                 // per warp: D x (T + 1) floats (padded rows: conflict-free transpose)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = T + 1, block = D * T;
    float *t = tile + warp * D * pitch;
    for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < n_outer; b += (int64_t)gridDim.x * 8) {
        const float *s = src + b * s_outer;
        float *o = dst + b * block;
        if (VEC) {
            for (int i = 4 * lane; i < block; i += 128) {    // i = j * T + tt, four consecutive tt of one j
                const float4 v = __ldg(reinterpret_cast<const float4 *>(s + i));
                const int j = i / T, tt = i - j * T;
                float *p = t + j * pitch + tt;
                p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
            }
            __syncwarp();
            for (int i = 4 * lane; i < block; i += 128) {    // i = tt * D + j, four consecutive j of one tt
                const int tt = i / D, j = i - tt * D;
                const float *p = t + j * pitch + tt;
                __stcs(reinterpret_cast<float4 *>(o + i), make_float4(p[0], p[pitch], p[2 * pitch], p[3 * pitch]));
            }
        } else {
            for (int i = lane; i < block; i += 32) {         // i = j * T + tt, consecutive in memory
                const int j = i / T, tt = i - j * T;
                t[j * pitch + tt] = __ldg(s + i);
            }
            __syncwarp();
            for (int i = lane; i < block; i += 32) {         // i = tt * D + j, consecutive in memory
                const int tt = i / D, j = i - tt * D;
                __stcs(o + i, t[j * pitch + tt]);
            }
        }
        __syncwarp();
    }
}

bool pack_rows_supported(int64_t n_inner, int d, int64_t s_outer, int64_t s_inner, int64_t s_d)
{
    return s_inner == 1 && s_d == n_inner && n_inner >= 1 && n_inner <= 64 && d >= 1 && d <= 128 &&
           s_outer >= (int64_t)d * n_inner && (int64_t)d * (n_inner + 1) * 8 * (int64_t)sizeof(float) <= 160 * 1024;
}

cudaError_t launch_pack_rows(const float *src, float *dst, int64_t n_outer, int64_t n_inner, int d, int64_t s_outer,
                             int64_t s_inner, int64_t s_d, int sm_count, cudaStream_t st)
{
    if (!pack_rows_supported(n_inner, d, s_outer, s_inner, s_d))
        return cudaErrorNotSupported;
    if (n_outer == 0)
        return cudaSuccess;
    const int smem = (int)(sizeof(float) * d * (n_inner + 1) * 8);
    const bool vec = n_inner % 4 == 0 && d % 4 == 0 && s_outer % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
    auto kern = vec ? vq_pack_rows_kernel<true> : vq_pack_rows_kernel<false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess)
        return err;
    const int64_t blocks = (n_outer + 7) / 8;
    const int grid = (int)(blocks < (int64_t)sm_count * 8 ? blocks : (int64_t)sm_count * 8);
    kern<<<grid, 256, smem, st>>>(src, dst, n_outer, (int)n_inner, d, s_outer);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Autoregressive pairs for the transformer (dataloader/base_dataloader.py:74-110, MyLatentAutoregressiveDataset):
//   x[w] = [start, ids[w][0 .. n-1]],   y[w] = [ids[w][0 .. n-1], end]     (both n + 1 long)
// One pass, ids read once: thread i of a window handles position i of both rows.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vq_ar_pairs_kernel(const int64_t *__restrict__ ids, int64_t n_windows, int n,
                                                          int64_t start_token, int64_t end_token, int64_t *__restrict__ x,
                                                          int64_t *__restrict__ y)
{
    const int64_t total = n_windows * (int64_t)(n + 1);
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t w = i / (n + 1);
        const int pos = (int)(i - w * (n + 1));
        const int64_t *row = ids + w * n;
        x[i] = pos == 0 ? start_token : row[pos - 1];
        y[i] = pos == n ? end_token : row[pos];
    }
}

cudaError_t launch_ar_pairs(const int64_t *ids, int64_t n_windows, int n, int64_t start_token, int64_t end_token, int64_t *x,
                            int64_t *y, int sm_count, cudaStream_t st)
{
    if (n_windows == 0)
        return cudaSuccess;
    const int64_t blocks = (n_windows * (int64_t)(n + 1) + 255) / 256;
    const int grid = (int)(blocks < (int64_t)sm_count * 8 ? blocks : (int64_t)sm_count * 8);
    vq_ar_pairs_kernel<<<grid, 256, 0, st>>>(ids, n_windows, n, start_token, end_token, x, y);
    return cudaGetLastError();
}

}  // namespace vqb
