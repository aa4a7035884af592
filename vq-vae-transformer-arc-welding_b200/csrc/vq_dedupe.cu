// vq_dedupe.cu -- bit-pattern fingerprints of cycles and their grouping, for the de-duplicating data-set builder.
//
// The reference's windows overlap (dataloader/asimow_dataloader.py:185-206: a stride of one cycle, 20 cycles per window), and
// its bulk loops (dataloader/latentspace_dataloader.py:171-263) encode every cycle of every window.  The builder here encodes
// every DISTINCT cycle of a batch once; "distinct" is decided on the bit pattern of the samples:
//
//   vq_row_keys_kernel      two 64-bit multiplicative hashes per row, key_s = sum_j uint32(word_j) * mult_s[j] (mod 2^64):
//                           the same function the host-side torch code evaluates (so keys agree wherever they are computed);
//                           one warp per row, 16-byte loads -- HBM-bound: the rows are read once (4 * words bytes per row)
//   vq_dedupe_insert_kernel open-addressing table over key 0 (atomicCAS claims a slot, atomicMin keeps the SMALLEST row
//                           index among the rows that share it): deterministic whatever the thread order
//   vq_dedupe_resolve_kernel first[i] = that smallest row index if its key 1 equals row i's as well, else i itself
//
// The caller still compares every row word for word with the row `first` names (a collision of both hashes costs a second
// look, never a wrong id) -- the kernels only propose the grouping.
#include "vq_common.cuh"

namespace vqb {

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;        // what cudaMemset(0xFF) leaves in the key column

__device__ __forceinline__ unsigned long long table_key(unsigned long long k) { return k == kEmpty ? kEmpty - 1 : k; }
__device__ __forceinline__ unsigned table_hash(unsigned long long k)
{   // the keys are sums of products with odd multipliers: well mixed in their high bits
    return (unsigned)((k * 0x9E3779B97F4A7C15ull) >> 32);
}

template <bool VEC>
__global__ void __launch_bounds__(256) vq_row_keys_kernel(const int *__restrict__ rows, int64_t n, int words,
                                                          const unsigned long long *__restrict__ mult,
                                                          unsigned long long *__restrict__ keys)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * 8;
    for (int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += warps) {
        const int *row = rows + r * words;
        unsigned long long a = 0, b = 0;
        if (VEC) {
            // all of a lane's 16-byte loads of a 128-chunk stretch are issued before the first product: with one load in
            // flight per lane the kernel ran at 1.4 TB/s (64 warps x 512 B per SM against ~1 us of loaded-HBM latency)
            const int4 *row4 = reinterpret_cast<const int4 *>(row);
            const int nv = words / 4;
            for (int base = 0; base < nv; base += 128) {
                int4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = base + 32 * u + lane;
                    v[u] = j < nv ? __ldg(row4 + j) : make_int4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = base + 32 * u + lane;
                    if (j < nv) {
                        const ulonglong2 *m0 = reinterpret_cast<const ulonglong2 *>(mult + 4 * j);
                        const ulonglong2 *m1 = reinterpret_cast<const ulonglong2 *>(mult + words + 4 * j);
                        const ulonglong2 p0 = __ldg(m0), p1 = __ldg(m0 + 1), q0 = __ldg(m1), q1 = __ldg(m1 + 1);
                        // words ZERO-extended to 64 bits, products and sums wrap: a 32 x 64 -> 64-bit multiply-add is one
                        // IMAD.WIDE.U32 + one IMAD (sign extension costs a 64 x 64 product, ~5 instructions: measured 2.7 TB/s)
                        a += (unsigned long long)(unsigned)v[u].x * p0.x + (unsigned long long)(unsigned)v[u].y * p0.y +
                             (unsigned long long)(unsigned)v[u].z * p1.x + (unsigned long long)(unsigned)v[u].w * p1.y;
                        b += (unsigned long long)(unsigned)v[u].x * q0.x + (unsigned long long)(unsigned)v[u].y * q0.y +
                             (unsigned long long)(unsigned)v[u].z * q1.x + (unsigned long long)(unsigned)v[u].w * q1.y;
                    }
                }
            }
        } else {
            for (int j = lane; j < words; j += 32) {
                const unsigned long long v = (unsigned long long)(unsigned)__ldg(row + j);
                a += v * __ldg(mult + j);
                b += v * __ldg(mult + words + j);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
        }
        if (lane == 0) {
            keys[2 * r] = a;
            keys[2 * r + 1] = b;
        }
    }
}

__global__ void __launch_bounds__(256) vq_dedupe_insert_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                               unsigned long long *__restrict__ tab_k0, int *__restrict__ tab_row,
                                                               unsigned mask, int *__restrict__ pos)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const unsigned long long k0 = table_key(keys[2 * i]);
        unsigned h = table_hash(k0) & mask;
        for (;;) {                        // the table has at least twice as many slots as there are rows: the walk ends
            const unsigned long long old = atomicCAS(tab_k0 + h, kEmpty, k0);
            if (old == kEmpty || old == k0) {
                atomicMin(tab_row + h, (int)i);
                pos[i] = (int)h;
                break;
            }
            h = (h + 1) & mask;
        }
    }
}

__global__ void __launch_bounds__(256) vq_dedupe_resolve_kernel(const unsigned long long *__restrict__ keys, int64_t n,
                                                                const int *__restrict__ tab_row, const int *__restrict__ pos,
                                                                int64_t *__restrict__ first)
{
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const int r = tab_row[pos[i]];                 // the smallest row index with this key 0 (final: kernel boundary)
        first[i] = keys[2 * (int64_t)r + 1] == keys[2 * i + 1] ? (int64_t)r : i;
    }
}

cudaError_t launch_row_keys(const void *rows, int64_t n, int words, const unsigned long long *mult, unsigned long long *keys,
                            int sm_count, cudaStream_t st)
{
    if (n == 0)
        return cudaSuccess;
    const int64_t blocks = (n + 7) / 8;
    const int grid = (int)(blocks < (int64_t)sm_count * 8 ? blocks : (int64_t)sm_count * 8);
    const bool vec = words % 4 == 0 && ((reinterpret_cast<uintptr_t>(rows) | reinterpret_cast<uintptr_t>(mult)) & 15) == 0;
    if (vec)
        vq_row_keys_kernel<true><<<grid, 256, 0, st>>>((const int *)rows, n, words, mult, keys);
    else
        vq_row_keys_kernel<false><<<grid, 256, 0, st>>>((const int *)rows, n, words, mult, keys);
    return cudaGetLastError();
}

// scratch layout: [tab_k0 cap x 8][tab_row cap x 4][pos n x 4], cap = dedupe_capacity(n)
int64_t dedupe_capacity(int64_t n)
{
    int64_t cap = 1024;
    while (cap < 2 * n)
        cap <<= 1;
    return cap;
}
size_t dedupe_scratch_bytes(int64_t n) { return (size_t)dedupe_capacity(n) * 12 + (size_t)(n > 0 ? n : 0) * 4; }

cudaError_t launch_dedupe_first(const unsigned long long *keys, int64_t n, void *scratch, int64_t *first, int sm_count,
                                cudaStream_t st)
{
    if (n == 0)
        return cudaSuccess;
    const int64_t cap = dedupe_capacity(n);
    unsigned long long *tab_k0 = (unsigned long long *)scratch;
    int *tab_row = (int *)(tab_k0 + cap);
    int *pos = tab_row + cap;
    cudaError_t err = cudaMemsetAsync(tab_k0, 0xFF, (size_t)cap * 8, st);                 // kEmpty
    if (err != cudaSuccess)
        return err;
    if ((err = cudaMemsetAsync(tab_row, 0x7F, (size_t)cap * 4, st)) != cudaSuccess)      // 0x7F7F7F7F: above every row index
        return err;
    const int64_t blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (int64_t)sm_count * 8 ? blocks : (int64_t)sm_count * 8);
    vq_dedupe_insert_kernel<<<grid, 256, 0, st>>>(keys, n, tab_k0, tab_row, (unsigned)(cap - 1), pos);
    if ((err = cudaGetLastError()) != cudaSuccess)
        return err;
    vq_dedupe_resolve_kernel<<<grid, 256, 0, st>>>(keys, n, tab_row, pos, first);
    return cudaGetLastError();
}

}  // namespace vqb
