// vq_fwd_fma.cu -- exact CUDA-core forward path + the small kernels around it.
//
// Replaces model/vector_quantizer.py:88-119 of the reference for ANY (K, D) and any
// strided input.  This path evaluates the oracle-order expression for every
// (vector, code) pair, so it is FMA-bound (2*K*D flop per vector on the fp32 pipe);
// the tcgen05 path in vq_fwd_tc.cu is the fast one and uses the device functions of
// this file's expression only for its exact refinement.
//
// Kernels
//   vq_prep_kernel       ||E_k||^2 in oracle order + non-finite column census (K threads)
//   vq_fwd_fma_kernel    one vector per thread (R vectors per thread), codebook broadcast
//                        from shared memory, running (min, argmin) in registers
//   vq_fwd_generic_kernel  any D (not a multiple of 4, or > 128): operands through L1
//   vq_finalize_kernel   loss = m + beta*m, perplexity from the integer histogram
//   vq_gather_kernel / vq_one_hot_kernel
#include "vq_common.cuh"

namespace vqb {

// ---------------------------------------------------------------------------------------
// prep: one thread per code
// ---------------------------------------------------------------------------------------
__global__ void vq_prep_kernel(const float *__restrict__ E, int K, int D, int kpad, float *__restrict__ ee,
                               int *colcnt, int *colwhich, WsHeader *hdr)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= kpad)
        return;
    if (k >= K) {
        ee[k] = __int_as_float(0x7f800000);  // +inf: a pad code never wins
        return;
    }
    const float *e = E + (size_t)k * D;
    float acc = 0.0f;
    for (int j = 0; j < D; ++j) {
        const float v = __ldg(e + j);
        acc = fmaf(v, v, acc);
        if (!isfinite(v)) {  // model/vector_quantizer.py:103 gathers with a GEMM: see oracle column_poison()
            if (atomicAdd(colcnt + j, 1) == 0)
                atomicAdd(&hdr->poisoned_columns, 1);
            atomicMax(colwhich + j, k + 1);
        }
    }
    ee[k] = acc;
}

cudaError_t launch_prep(const float *E, int K, int D, int kpad, float *ee, int *colcnt, int *colwhich,
                        WsHeader *hdr, cudaStream_t st)
{
    vq_prep_kernel<<<(kpad + 127) / 128, 128, 0, st>>>(E, K, D, kpad, ee, colcnt, colwhich, hdr);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// helpers shared by the forward kernels
// ---------------------------------------------------------------------------------------
// Value the reference's gather-by-GEMM yields for column j of a row that picked `code`.
__device__ __forceinline__ float gathered(const float *__restrict__ E, int D, int code, int j,
                                          const int *__restrict__ colcnt, const int *__restrict__ colwhich,
                                          bool poisoned)
{
    const float e = __ldg(E + (size_t)code * D + j);
    if (!poisoned)
        return e;
    const int c = colcnt[j];
    if (c == 0 || (c == 1 && colwhich[j] == code + 1))
        return e;
    return __int_as_float(0x7fc00000);
}

// torch.argmin over a row that contains a NaN distance: index of the first NaN.
__device__ __noinline__ int first_nan_code(const float *zrow, int64_t s_d, float zz, const float *__restrict__ E,
                                           const float *__restrict__ ee, int K, int D)
{
    for (int k = 0; k < K; ++k) {
        const float *e = E + (size_t)k * D;
        float acc = 0.0f;
        for (int j = 0; j < D; ++j)
            acc = fmaf(zrow[j * s_d], __ldg(e + j), acc);
        const float dist = ref_distance(zz, ee[k], acc);
        if (dist != dist)
            return k;
    }
    return 0;
}

__device__ __forceinline__ void block_store_partial(double v, double *partials, int accumulate)
{
    __shared__ double red[32];
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double t = lane < nw ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0)
            partials[blockIdx.x] = accumulate ? partials[blockIdx.x] + t : t;
    }
}

// ---------------------------------------------------------------------------------------
// main FMA kernel: compile-time D (multiple of 4), R vectors per thread
// ---------------------------------------------------------------------------------------
template <int DT, int R>
__global__ void __launch_bounds__(256, (DT * R <= 64) ? 2 : 1) vq_fwd_fma_kernel(const FwdParams p)
{
    extern __shared__ __align__(16) float smem[];
    constexpr int ES = DT + 4;  // padded row stride: keeps 16-byte alignment
    float *cb = smem;                          // [kt][ES]
    float *ees = smem + (size_t)p.kt * ES;     // [kt]

    const int tid = threadIdx.x;
    const int K = p.K;
    const bool single = p.kt >= K;
    const bool poisoned = p.hdr_in->poisoned_columns != 0;
    const bool vec_in = p.z.s_d == 1 && ((reinterpret_cast<uintptr_t>(p.z.base) & 15) == 0) &&
                        (p.z.s_outer % 4 == 0) && (p.z.s_inner % 4 == 0);

    auto stage = [&](int k0) {
        const int nk = min(p.kt, ((K - k0) + 3) & ~3);
        for (int e = tid; e < nk * (DT / 4); e += 256) {
            const int c = e / (DT / 4), q = e - c * (DT / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + c < K)
                v = __ldg(reinterpret_cast<const float4 *>(p.E + (size_t)(k0 + c) * DT) + q);
            *reinterpret_cast<float4 *>(cb + (size_t)c * ES + 4 * q) = v;
        }
        for (int c = tid; c < nk; c += 256)
            ees[c] = p.ee[k0 + c];  // pads hold +inf
    };

    if (single) {
        stage(0);
        __syncthreads();
    }

    double sq = 0.0;
    const int64_t rows_per_tile = 256 * R;
    const int64_t n_tiles = (p.z.n_rows + rows_per_tile - 1) / rows_per_tile;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        float zr[R][DT];
        float zz[R], best[R];
        int bidx[R];
        int64_t row[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            row[r] = tile * rows_per_tile + (int64_t)r * 256 + tid;
            const bool ok = row[r] < p.z.n_rows;
            const float *src = p.z.row(ok ? row[r] : 0);
            if (vec_in) {
#pragma unroll
                for (int q = 0; q < DT / 4; ++q) {
                    const float4 v = ok ? __ldg(reinterpret_cast<const float4 *>(src) + q)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
                    zr[r][4 * q + 0] = v.x; zr[r][4 * q + 1] = v.y;
                    zr[r][4 * q + 2] = v.z; zr[r][4 * q + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < DT; ++j)
                    zr[r][j] = ok ? __ldg(src + j * p.z.s_d) : 0.f;
            }
            float acc = 0.0f;
#pragma unroll
            for (int j = 0; j < DT; ++j)
                acc = fmaf(zr[r][j], zr[r][j], acc);
            zz[r] = acc;
            best[r] = __int_as_float(0x7f800000);
            bidx[r] = 0;
        }

        for (int k0 = 0; k0 < K; k0 += p.kt) {
            if (!single) {
                __syncthreads();
                stage(k0);
                __syncthreads();
            }
            const int nk = min(p.kt, ((K - k0) + 3) & ~3);
            for (int c = 0; c < nk; c += 4) {
                float acc[R][4];
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        acc[r][u] = 0.0f;
#pragma unroll
                for (int q = 0; q < DT / 4; ++q) {
                    float4 e[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        e[u] = *reinterpret_cast<const float4 *>(cb + (size_t)(c + u) * ES + 4 * q);
#pragma unroll
                    for (int r = 0; r < R; ++r)
#pragma unroll
                        for (int u = 0; u < 4; ++u) {  // ascending chain per (vector, code)
                            acc[r][u] = fmaf(zr[r][4 * q + 0], e[u].x, acc[r][u]);
                            acc[r][u] = fmaf(zr[r][4 * q + 1], e[u].y, acc[r][u]);
                            acc[r][u] = fmaf(zr[r][4 * q + 2], e[u].z, acc[r][u]);
                            acc[r][u] = fmaf(zr[r][4 * q + 3], e[u].w, acc[r][u]);
                        }
                }
                const float4 ee4 = *reinterpret_cast<const float4 *>(ees + c);
                const float eev[4] = {ee4.x, ee4.y, ee4.z, ee4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const float dist = ref_distance(zz[r], eev[u], acc[r][u]);
                        bidx[r] = dist < best[r] ? k0 + c + u : bidx[r];  // strict <: lowest index on ties
                        best[r] = min_nan(best[r], dist);
                    }
            }
        }

#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = row[r] < p.z.n_rows;
            int code = bidx[r];
            if (ok && best[r] != best[r]) {  // some distance is NaN: torch.argmin returns the first one
                code = first_nan_code(p.z.row(row[r]), p.z.s_d, zz[r], p.E, p.ee, K, DT);
                if (p.stats)
                    atomicAdd(p.stats + 2, 1ULL);
            }
            warp_histogram_add(p.counts, ok ? code : -1);
            if (!ok)
                continue;
            if (p.idx)
                p.idx[row[r]] = code;
            float rsq = 0.0f;
            float *dst = p.zq ? p.zq + row[r] * DT : nullptr;
#pragma unroll
            for (int q = 0; q < DT / 4; ++q) {
                float o[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int j = 4 * q + t;
                    const float e = gathered(p.E, DT, code, j, p.colcnt, p.colwhich, poisoned);
                    const float diff = __fsub_rn(e, zr[r][j]);     // :107-108 residual
                    rsq = __fadd_rn(rsq, __fmul_rn(diff, diff));
                    o[t] = __fadd_rn(zr[r][j], diff);              // :111 straight-through value
                }
                if (dst)
                    *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
            }
            sq += (double)rsq;
        }
    }
    block_store_partial(sq, p.partials, p.accumulate);
}

// ---------------------------------------------------------------------------------------
// fully general kernel: any D, any strides; operands read through L1/L2
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vq_fwd_generic_kernel(const FwdParams p)
{
    const int K = p.K, D = p.D;
    const bool poisoned = p.hdr_in->poisoned_columns != 0;
    double sq = 0.0;
    const int64_t n_tiles = (p.z.n_rows + 255) / 256;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t row = tile * 256 + threadIdx.x;
        const bool ok = row < p.z.n_rows;
        const float *zrow = p.z.row(ok ? row : 0);
        const int64_t sd = p.z.s_d;
        float zz = 0.0f;
        for (int j = 0; j < D; ++j) {
            const float v = zrow[j * sd];
            zz = fmaf(v, v, zz);
        }
        float best = __int_as_float(0x7f800000);
        int code = 0;
        for (int k = 0; k < K; ++k) {
            const float *e = p.E + (size_t)k * D;
            float acc = 0.0f;
            for (int j = 0; j < D; ++j)
                acc = fmaf(zrow[j * sd], __ldg(e + j), acc);
            const float dist = ref_distance(zz, p.ee[k], acc);
            code = dist < best ? k : code;
            best = min_nan(best, dist);
        }
        if (ok && best != best) {
            code = first_nan_code(zrow, sd, zz, p.E, p.ee, K, D);
            if (p.stats)
                atomicAdd(p.stats + 2, 1ULL);
        }
        warp_histogram_add(p.counts, ok ? code : -1);
        if (!ok)
            continue;
        if (p.idx)
            p.idx[row] = code;
        float rsq = 0.0f;
        for (int j = 0; j < D; ++j) {
            const float zj = zrow[j * sd];
            const float e = gathered(p.E, D, code, j, p.colcnt, p.colwhich, poisoned);
            const float diff = __fsub_rn(e, zj);
            rsq = __fadd_rn(rsq, __fmul_rn(diff, diff));
            if (p.zq)
                p.zq[row * D + j] = __fadd_rn(zj, diff);
        }
        sq += (double)rsq;
    }
    block_store_partial(sq, p.partials, p.accumulate);
}

template <int DT, int R>
static cudaError_t launch_fma_t(FwdParams p, int sm_count, int max_smem, int *n_ctas, cudaStream_t st)
{
    constexpr int ES = DT + 4;
    const size_t per_code = sizeof(float) * (ES + 1);
    const int blocks_per_sm = (DT * R <= 64) ? 2 : 1;
    // Shared-memory budget per CTA so that `blocks_per_sm` CTAs stay resident.
    const size_t budget = (size_t)max_smem / blocks_per_sm - 2048;
    int kt = (int)((p.K + 3) & ~3);
    if ((size_t)kt * per_code > budget)
        kt = (int)(budget / per_code) & ~3;
    p.kt = kt;
    const size_t smem = (size_t)kt * per_code;
    cudaError_t err = cudaFuncSetAttribute(vq_fwd_fma_kernel<DT, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
    if (err != cudaSuccess)
        return err;
    const int64_t tiles = (p.z.n_rows + 256 * R - 1) / (256 * R);
    int grid = (int)(tiles < (int64_t)sm_count * blocks_per_sm ? tiles : (int64_t)sm_count * blocks_per_sm);
    if (grid > kMaxPartials)
        grid = kMaxPartials;
    if (grid < 1)
        grid = 1;
    *n_ctas = grid;
    vq_fwd_fma_kernel<DT, R><<<grid, 256, smem, st>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_fwd_fma(const FwdParams &p, int sm_count, int max_smem, int *n_ctas, cudaStream_t st)
{
    switch (p.D) {
    case 4:   return launch_fma_t<4, 2>(p, sm_count, max_smem, n_ctas, st);
    case 8:   return launch_fma_t<8, 2>(p, sm_count, max_smem, n_ctas, st);
    case 16:  return launch_fma_t<16, 2>(p, sm_count, max_smem, n_ctas, st);
    case 32:  return launch_fma_t<32, 2>(p, sm_count, max_smem, n_ctas, st);
    case 64:  return launch_fma_t<64, 1>(p, sm_count, max_smem, n_ctas, st);
    case 128: return launch_fma_t<128, 1>(p, sm_count, max_smem, n_ctas, st);
    default: break;
    }
    const int64_t tiles = (p.z.n_rows + 255) / 256;
    int grid = (int)(tiles < (int64_t)sm_count * 4 ? tiles : (int64_t)sm_count * 4);
    if (grid > kMaxPartials)
        grid = kMaxPartials;
    if (grid < 1)
        grid = 1;
    *n_ctas = grid;
    vq_fwd_generic_kernel<<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// finalize: loss and perplexity (model/vector_quantizer.py:107-108,114-115)
// sq_mode: 0 = total is the sum of `partials`; 1 = add partials into *sq_total_io and use
// the running total (host path, several chunks); 2 = like 1 but do not emit scalars yet.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vq_finalize_kernel(const unsigned long long *__restrict__ counts, int K,
                                                           const double *__restrict__ partials, int n_partials,
                                                           double *sq_total_io, int sq_mode, int64_t n_rows, int D,
                                                           float beta, float *loss, float *perplexity)
{
    __shared__ double red[8];
    __shared__ double total_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double s = 0.0;
    for (int i = tid; i < n_partials; i += 256)  // fixed order: deterministic
        s += partials[i];
    s = warp_sum(s);
    if (lane == 0)
        red[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w)
            t += red[w];
        if (sq_mode != 0) {
            t += *sq_total_io;
            *sq_total_io = t;
        }
        total_s = t;
    }
    __syncthreads();
    if (sq_mode == 2)
        return;
    if (tid == 0 && loss) {
        const float m = n_rows > 0 ? (float)(total_s / ((double)n_rows * (double)D)) : __int_as_float(0x7fc00000);
        *loss = __fadd_rn(m, __fmul_rn(beta, m));  // both means are the same number (:107-108)
    }
    if (perplexity) {
        double h = 0.0;
        for (int k = tid; k < K; k += 256) {
            const float pk = (float)counts[k] / (float)n_rows;   // :114 mean of the one-hot column
            h += (double)(pk * logf(pk + 1e-10f));               // :115
        }
        h = warp_sum(h);
        __syncthreads();
        if (lane == 0)
            red[warp] = h;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w)
                t += red[w];
            *perplexity = expf((float)(-t));
        }
    }
}

cudaError_t launch_finalize(const unsigned long long *counts, int K, const double *partials, int n_partials,
                            double *sq_total_io, int sq_mode, int64_t n_rows, int D, float beta,
                            float *loss, float *perplexity, cudaStream_t st)
{
    vq_finalize_kernel<<<1, 256, 0, st>>>(counts, K, partials, n_partials, sq_total_io, sq_mode, n_rows, D, beta,
                                          loss, perplexity);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// gather (model/vector_quantizer.py:121-131) and one-hot (:98-100)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vq_gather_kernel(const int64_t *__restrict__ idx, int64_t n,
                                                         const float *__restrict__ E, int K, int D,
                                                         float *__restrict__ out, int *bad_index)
{
    const int64_t total = n * D;
    for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
        const int64_t i = t / D;
        const int j = (int)(t - i * D);
        const int64_t c = idx[i];
        if (c < 0 || c >= K) {
            out[t] = __int_as_float(0x7fc00000);
            if (bad_index && j == 0)
                *bad_index = 1;
        } else {
            out[t] = __ldg(E + c * D + j);
        }
    }
}

cudaError_t launch_gather(const int64_t *idx, int64_t n, const float *E, int K, int D, float *out,
                          int *bad_index, cudaStream_t st)
{
    if (n == 0)
        return cudaSuccess;
    const int64_t blocks = (n * D + 255) / 256;
    vq_gather_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(idx, n, E, K, D, out, bad_index);
    return cudaGetLastError();
}

// each row of K floats is written once (zeros and the single 1); 16-byte stores when K % 4 == 0
template <bool VEC>
__global__ void __launch_bounds__(256) vq_one_hot_kernel(const int64_t *__restrict__ idx, int64_t n, int K,
                                                          float *__restrict__ onehot)
{
    if (VEC) {
        const int kq = K >> 2;
        const int64_t total = n * kq;
        for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
            const int64_t i = t / kq;
            const int c = (int)(t - i * kq) << 2;
            const int hit = (int)(idx[i] - c);      // 0..3 when the 1 falls into this quad
            float4 v;
            v.x = hit == 0 ? 1.0f : 0.0f; v.y = hit == 1 ? 1.0f : 0.0f;
            v.z = hit == 2 ? 1.0f : 0.0f; v.w = hit == 3 ? 1.0f : 0.0f;
            reinterpret_cast<float4 *>(onehot)[t] = v;
        }
    } else {
        const int64_t total = n * K;
        for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
            const int64_t i = t / K;
            const int c = (int)(t - i * K);
            onehot[t] = (idx[i] == c) ? 1.0f : 0.0f;
        }
    }
}

cudaError_t launch_one_hot(const int64_t *idx, int64_t n, int K, float *onehot, cudaStream_t st)
{
    if (n == 0)
        return cudaSuccess;
    const bool vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(onehot) & 15) == 0);
    const int64_t work = vec ? n * (K / 4) : n * K;
    const int64_t blocks = (work + 255) / 256;
    const int grid = (int)(blocks < 148 * 32 ? blocks : 148 * 32);
    if (vec)
        vq_one_hot_kernel<true><<<grid, 256, 0, st>>>(idx, n, K, onehot);
    else
        vq_one_hot_kernel<false><<<grid, 256, 0, st>>>(idx, n, K, onehot);
    return cudaGetLastError();
}

}  // namespace vqb
