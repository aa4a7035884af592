// vq_hostpipe.cu -- host-buffer entry points (vqb_host_*, vqb_encode_host).
//
// This is the call shape of the reference's offline tokeniser
// (dataloader/latentspace_dataloader.py:225-238): host arrays in, ids (and optionally the
// quantised vectors) back on the host.  The reference does one blocking H2D copy, one
// launch storm and one blocking .cpu() per 8192 vectors; here successive chunks are
// pipelined over `depth` streams so that H2D, the kernels and D2H of neighbouring chunks
// overlap, and the histogram / loss sums stay on the device until the end.
#include <new>
#include <vector>

#include "vq_common.cuh"

using namespace vqb;

struct vqb_host_ctx {
    int device = 0;
    int64_t chunk_rows = 0;
    int d = 0, k = 0, depth = 0;
    vqb_device_info info{};
    WsLayout L{};
    char *ws = nullptr;                 // shared: header, ee, column census, counts, sq
    float *d_codebook = nullptr;
    double *d_partials = nullptr;       // [depth][kMaxPartials]
    float *d_scalars = nullptr;         // loss, perplexity
    unsigned long long *d_stats = nullptr;
    bool codebook_set = false;
    bool base_ready = false;            // code norms / census in `ws` belong to the current codebook
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    float last_ms = -1.f;
    struct Slot {
        cudaStream_t stream = nullptr;
        cudaEvent_t done = nullptr;
        float *d_z = nullptr, *d_zq = nullptr;
        int64_t *d_idx = nullptr;
        float *tc_scratch = nullptr;
        void *d_ids_small = nullptr;    // u8 / u16 copy of the ids (VQB_IDS_U8 / VQB_IDS_U16)
        bool image_ready = false;       // tc_scratch holds the tcgen05 operand image of the current codebook
    };
    std::vector<Slot> slots;
};

#define VQB_TRY(expr)                         \
    do {                                      \
        cudaError_t e__ = (expr);             \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

extern "C" {

int vqb_query(int device, vqb_device_info *out);

int vqb_host_destroy(vqb_host_ctx *ctx)
{
    if (!ctx)
        return VQB_OK;
    cudaSetDevice(ctx->device);
    for (auto &s : ctx->slots) {
        if (s.stream) cudaStreamSynchronize(s.stream);
        if (s.d_z) cudaFree(s.d_z);
        if (s.d_zq) cudaFree(s.d_zq);
        if (s.d_idx) cudaFree(s.d_idx);
        if (s.tc_scratch) cudaFree(s.tc_scratch);
        if (s.d_ids_small) cudaFree(s.d_ids_small);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->d_codebook) cudaFree(ctx->d_codebook);
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_scalars) cudaFree(ctx->d_scalars);
    if (ctx->d_stats) cudaFree(ctx->d_stats);
    delete ctx;
    return VQB_OK;
}

int vqb_host_create(int device, int64_t chunk_rows, int d, int k, int depth, vqb_host_ctx **out)
{
    if (!out || chunk_rows <= 0 || d <= 0 || k <= 0 || depth < 1 || depth > 16)
        return VQB_E_ARG;
    vqb_device_info info;
    int rc = vqb_query(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    VQB_TRY(cudaSetDevice(device));
    vqb_host_ctx *ctx = new (std::nothrow) vqb_host_ctx();
    if (!ctx)
        return VQB_E_HOSTCTX;
    ctx->device = device; ctx->chunk_rows = chunk_rows; ctx->d = d; ctx->k = k; ctx->depth = depth;
    ctx->info = info;
    ctx->L = ws_layout(k, d);
    cudaError_t err = cudaSuccess;
    auto fail = [&](cudaError_t e) { vqb_host_destroy(ctx); return (int)e; };
    if ((err = cudaMalloc(&ctx->ws, ctx->L.total)) != cudaSuccess) return fail(err);
    if ((err = cudaMalloc(&ctx->d_codebook, sizeof(float) * (size_t)k * d)) != cudaSuccess) return fail(err);
    if ((err = cudaMalloc(&ctx->d_partials, sizeof(double) * (size_t)depth * kMaxPartials)) != cudaSuccess) return fail(err);
    if ((err = cudaMalloc(&ctx->d_scalars, sizeof(float) * 4)) != cudaSuccess) return fail(err);
    if ((err = cudaMalloc(&ctx->d_stats, sizeof(unsigned long long) * 4)) != cudaSuccess) return fail(err);
    if ((err = cudaEventCreate(&ctx->ev_begin)) != cudaSuccess) return fail(err);
    if ((err = cudaEventCreate(&ctx->ev_end)) != cudaSuccess) return fail(err);
    ctx->slots.resize(depth);
    for (auto &s : ctx->slots) {
        if ((err = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(err);
        if ((err = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return fail(err);
        if ((err = cudaMalloc(&s.d_z, sizeof(float) * (size_t)chunk_rows * d)) != cudaSuccess) return fail(err);
        if ((err = cudaMalloc(&s.d_zq, sizeof(float) * (size_t)chunk_rows * d)) != cudaSuccess) return fail(err);
        if ((err = cudaMalloc(&s.d_idx, sizeof(int64_t) * (size_t)chunk_rows)) != cudaSuccess) return fail(err);
        if ((err = cudaMalloc(&s.tc_scratch, sizeof(float) * tc_scratch_floats(k, d))) != cudaSuccess) return fail(err);
        if ((err = cudaMalloc(&s.d_ids_small, sizeof(uint16_t) * (size_t)chunk_rows)) != cudaSuccess) return fail(err);
    }
    *out = ctx;
    return VQB_OK;
}

int vqb_host_set_codebook(vqb_host_ctx *ctx, const float *codebook_host)
{
    if (!ctx || !codebook_host)
        return VQB_E_ARG;
    VQB_TRY(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    VQB_TRY(cudaMemcpyAsync(ctx->d_codebook, codebook_host, sizeof(float) * (size_t)ctx->k * ctx->d,
                            cudaMemcpyHostToDevice, st));
    VQB_TRY(cudaStreamSynchronize(st));
    ctx->codebook_set = true;
    ctx->base_ready = false;
    for (auto &s : ctx->slots)
        s.image_ready = false;
    return VQB_OK;
}

}  // extern "C"

// int64 ids -> u8 / u16 (K <= 256 / 65536): the id array the tokeniser keeps is 8x / 4x smaller on the way back
template <typename T>
__global__ void vq_narrow_ids_kernel(const int64_t *__restrict__ idx, T *__restrict__ out, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (T)idx[i];
}

static int encode_host_impl(vqb_host_ctx *ctx, const float *z_host, int64_t n, float beta, float *zq_host,
                            void *idx_host, float *loss_host, float *perplexity_host,
                            unsigned long long *counts_host, unsigned flags, int *launches_out)
{
    const int d = ctx->d, k = ctx->k;
    const WsLayout &L = ctx->L;
    char *ws = ctx->ws;
    WsHeader *hdr = (WsHeader *)ws;
    float *ee = (float *)(ws + L.off_ee);
    int *colcnt = (int *)(ws + L.off_colcnt);
    int *colwhich = (int *)(ws + L.off_colwhich);
    unsigned long long *cnt = (unsigned long long *)(ws + L.off_counts);
    int launches = 0;

    // the path and its validity are decided before anything is enqueued
    const unsigned id_kind = flags & VQB_IDS_MASK;
    const int id_bytes = id_kind == VQB_IDS_U8 ? 1 : id_kind == VQB_IDS_U16 ? 2 : 8;
    if ((id_kind == VQB_IDS_U8 && k > 256) || (id_kind == VQB_IDS_U16 && k > 65536) || id_kind == VQB_IDS_MASK)
        return VQB_E_ARG;
    const bool want_idx = idx_host != nullptr;
    const bool tc_chunked = tc_chunked_supported(k, d) && want_idx;   // running best lives in d_idx
    const bool tc_ok = (tc_shape_supported(k, d) || tc_chunked) && ctx->chunk_rows < (1ll << 31);
    const unsigned path_req = flags & VQB_PATH_MASK;
    if (path_req == VQB_PATH_TC && !tc_ok)
        return VQB_E_UNSUPPORTED;
    if (path_req == VQB_PATH_MASK)
        return VQB_E_ARG;

    cudaStream_t s0 = ctx->slots[0].stream;
    VQB_TRY(cudaEventRecord(ctx->ev_begin, s0));
    if (!ctx->base_ready) {
        VQB_TRY(cudaMemsetAsync(ws, 0, L.off_counts, s0));                       // header, census
        VQB_TRY(launch_prep(ctx->d_codebook, k, d, L.kpad, ee, colcnt, colwhich, hdr, s0));
        ++launches;
    }
    VQB_TRY(cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * k, s0));
    VQB_TRY(cudaMemsetAsync(ctx->d_partials, 0, sizeof(double) * (size_t)ctx->depth * kMaxPartials, s0));
    VQB_TRY(cudaMemsetAsync(ctx->d_stats, 0, sizeof(unsigned long long) * 4, s0));
    VQB_TRY(cudaEventRecord(ctx->slots[0].done, s0));
    for (int s = 1; s < ctx->depth; ++s)
        VQB_TRY(cudaStreamWaitEvent(ctx->slots[s].stream, ctx->slots[0].done, 0));

    const int64_t n_chunks = (n + ctx->chunk_rows - 1) / ctx->chunk_rows;
    for (int64_t c = 0; c < n_chunks; ++c) {
        auto &sl = ctx->slots[c % ctx->depth];
        const int64_t row0 = c * ctx->chunk_rows;
        const int64_t rows = (n - row0 < ctx->chunk_rows) ? (n - row0) : ctx->chunk_rows;
        VQB_TRY(cudaMemcpyAsync(sl.d_z, z_host + row0 * d, sizeof(float) * (size_t)rows * d,
                                cudaMemcpyHostToDevice, sl.stream));
        FwdParams p{};
        p.z.base = sl.d_z; p.z.n_rows = rows; p.z.n_inner = 1; p.z.s_outer = d; p.z.s_inner = d; p.z.s_d = 1;
        p.E = ctx->d_codebook; p.K = k; p.D = d; p.ee = ee; p.colcnt = colcnt; p.colwhich = colwhich; p.hdr_in = hdr;
        p.zq = zq_host ? sl.d_zq : nullptr;
        p.idx = want_idx ? sl.d_idx : nullptr;
        p.counts = cnt;
        p.partials = ctx->d_partials + (size_t)(c % ctx->depth) * kMaxPartials;
        p.accumulate = 1;
        p.stats = ctx->d_stats;
        p.need_sq = loss_host != nullptr;
        unsigned path = path_req;
        if (path == VQB_PATH_AUTO)
            path = (tc_ok && rows >= 128) ? VQB_PATH_TC : VQB_PATH_FMA;
        int n_ctas = 0;
        if (path == VQB_PATH_TC) {
            int nl = 0;
            if (tc_chunked) {
                VQB_TRY(launch_fwd_tc_chunked(p, sl.tc_scratch, ctx->info.sm_count, ctx->info.max_smem_per_block, &n_ctas,
                                              &nl, sl.stream, nullptr, nullptr));
                sl.image_ready = false;
            } else {
                VQB_TRY(launch_fwd_tc(p, sl.tc_scratch, ctx->info.sm_count, ctx->info.max_smem_per_block, &n_ctas, &nl,
                                      sl.stream, nullptr, nullptr, sl.image_ready));
                sl.image_ready = true;       // this slot's image now belongs to the current codebook
            }
            launches += nl;
        } else {
            VQB_TRY(launch_fwd_fma(p, ctx->info.sm_count, ctx->info.max_smem_per_block, &n_ctas, sl.stream));
            ++launches;
        }
        if (zq_host)
            VQB_TRY(cudaMemcpyAsync(zq_host + row0 * d, sl.d_zq, sizeof(float) * (size_t)rows * d,
                                    cudaMemcpyDeviceToHost, sl.stream));
        if (want_idx) {
            const void *src = sl.d_idx;
            if (id_bytes != 8) {
                const int grid = (int)((rows + 255) / 256 < 4096 ? (rows + 255) / 256 : 4096);
                if (id_bytes == 1)
                    vq_narrow_ids_kernel<uint8_t><<<grid, 256, 0, sl.stream>>>(sl.d_idx, (uint8_t *)sl.d_ids_small, rows);
                else
                    vq_narrow_ids_kernel<uint16_t><<<grid, 256, 0, sl.stream>>>(sl.d_idx, (uint16_t *)sl.d_ids_small, rows);
                VQB_TRY(cudaGetLastError());
                ++launches;
                src = sl.d_ids_small;
            }
            VQB_TRY(cudaMemcpyAsync((char *)idx_host + (size_t)row0 * id_bytes, src, (size_t)rows * id_bytes,
                                    cudaMemcpyDeviceToHost, sl.stream));
        }
    }
    for (int s = 1; s < ctx->depth; ++s) {
        VQB_TRY(cudaEventRecord(ctx->slots[s].done, ctx->slots[s].stream));
        VQB_TRY(cudaStreamWaitEvent(s0, ctx->slots[s].done, 0));
    }
    VQB_TRY(launch_finalize(cnt, k, ctx->d_partials, ctx->depth * kMaxPartials, (double *)(ws + L.off_sq), 0, n, d,
                            beta, ctx->d_scalars, ctx->d_scalars + 1, s0));
    ++launches;
    float scal[2];
    VQB_TRY(cudaMemcpyAsync(scal, ctx->d_scalars, sizeof(float) * 2, cudaMemcpyDeviceToHost, s0));
    if (counts_host)
        VQB_TRY(cudaMemcpyAsync(counts_host, cnt, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, s0));
    VQB_TRY(cudaEventRecord(ctx->ev_end, s0));
    VQB_TRY(cudaStreamSynchronize(s0));
    VQB_TRY(cudaEventElapsedTime(&ctx->last_ms, ctx->ev_begin, ctx->ev_end));
    ctx->base_ready = true;
    count_launches(launches);
    if (loss_host) *loss_host = scal[0];
    if (perplexity_host) *perplexity_host = scal[1];
    if (launches_out) *launches_out = launches;
    return VQB_OK;
}

extern "C" {

int vqb_encode_host(vqb_host_ctx *ctx, const float *z_host, int64_t n, float beta,
                    float *zq_host, void *idx_host, float *loss_host, float *perplexity_host,
                    unsigned long long *counts_host, unsigned flags, int *launches_out)
{
    if (!ctx || n < 0 || (n > 0 && !z_host))
        return VQB_E_ARG;
    if (!ctx->codebook_set)
        return VQB_E_HOSTCTX;
    VQB_TRY(cudaSetDevice(ctx->device));
    const int rc = encode_host_impl(ctx, z_host, n, beta, zq_host, idx_host, loss_host, perplexity_host, counts_host,
                                    flags, launches_out);
    if (rc != VQB_OK) {
        // an error return must not leave copies in flight that still target the caller's buffers, and what the
        // workspace / the slots hold can no longer be vouched for
        for (auto &s : ctx->slots) {
            cudaStreamSynchronize(s.stream);
            s.image_ready = false;
        }
        ctx->base_ready = false;
    }
    return rc;
}

int vqb_host_last_ms(vqb_host_ctx *ctx, float *ms)
{
    if (!ctx || !ms)
        return VQB_E_ARG;
    *ms = ctx->last_ms;
    return VQB_OK;
}

}  // extern "C"
