// vq_api.cu -- the extern "C" boundary declared in include/vqb200.h.
//
// Every device entry point: validates arguments, carves the caller-owned workspace,
// enqueues kernels on the caller's stream and returns without synchronising.
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>

#include "vq_common.cuh"

using namespace vqb;

namespace {

struct DevInfo {
    bool known = false;
    cudaError_t err = cudaSuccess;
    vqb_device_info info{};
};
constexpr int kMaxDevices = 64;
DevInfo g_dev[kMaxDevices];
std::mutex g_mu;

int device_info(int device, vqb_device_info *out)
{
    if (device < 0 || device >= kMaxDevices)
        return VQB_E_ARG;
    std::lock_guard<std::mutex> lock(g_mu);
    DevInfo &d = g_dev[device];
    if (!d.known) {
        cudaDeviceProp prop;
        d.err = cudaGetDeviceProperties(&prop, device);
        if (d.err == cudaSuccess) {
            d.info.device = device;
            d.info.cc_major = prop.major;
            d.info.cc_minor = prop.minor;
            d.info.sm_count = prop.multiProcessorCount;
            d.info.max_smem_per_block = (int)prop.sharedMemPerBlockOptin;
            d.info.l2_bytes = (size_t)prop.l2CacheSize;
            d.info.total_mem = prop.totalGlobalMem;
            d.known = true;
        }
    }
    if (d.err != cudaSuccess)
        return (int)d.err;
    *out = d.info;
    return VQB_OK;
}

std::atomic<long long> g_launches{0};
std::atomic<int> g_profile{0};
std::mutex g_prof_mu;
// brackets are kept per device: vqb_profile_collect returns those of the calling thread's current device only,
// so processes that drive several GPUs do not mix their launches
struct ProfBracket {
    int dev;
    cudaEvent_t first, second;
};
std::vector<ProfBracket> g_prof_events;

inline bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

}  // namespace

namespace vqb {
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace vqb

extern "C" {

long long vqb_launch_counter(void) { return g_launches.load(std::memory_order_relaxed); }

/* Debug: route the tcgen05 kernel through its tracing build; `buf` is a device buffer of
 * vqb_debug_tc_trace_words() uint64 (NULL switches tracing off). */
int vqb_debug_set_tc_trace(unsigned long long *buf)
{
    set_tc_trace(buf);
    return VQB_OK;
}
size_t vqb_debug_tc_trace_words(void) { return tc_trace_words(); }

int vqb_debug_set_filter(int mode)
{
    if (mode < -1 || mode > 1)
        return VQB_E_ARG;
    set_tc_filter(mode);
    return VQB_OK;
}

int vqb_profile_enable(int on)
{
    g_profile.store(on ? 1 : 0);
    return VQB_OK;
}

int vqb_profile_collect(double *ms_sum, int *launches)
{
    int dev = 0;
    cudaError_t derr = cudaGetDevice(&dev);
    if (derr != cudaSuccess)
        return (int)derr;
    std::lock_guard<std::mutex> lock(g_prof_mu);
    double total = 0.0;
    int n = 0;
    cudaError_t first_err = cudaSuccess;
    std::vector<ProfBracket> other;
    for (auto &pr : g_prof_events) {
        if (pr.dev != dev) {
            other.push_back(pr);
            continue;
        }
        float ms = 0.f;
        cudaError_t err = cudaEventSynchronize(pr.second);
        if (err == cudaSuccess)
            err = cudaEventElapsedTime(&ms, pr.first, pr.second);
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
        if (err != cudaSuccess) {
            if (first_err == cudaSuccess)
                first_err = err;
            continue;
        }
        total += ms;
        ++n;
    }
    g_prof_events.swap(other);
    if (first_err != cudaSuccess)
        return (int)first_err;
    if (ms_sum) *ms_sum = total;
    if (launches) *launches = n;
    return VQB_OK;
}

int vqb_version(void) { return VQB_VERSION; }

const char *vqb_error_string(int code)
{
    switch (code) {
    case VQB_OK: return "ok";
    case VQB_E_ARG: return "vqb: invalid argument (null pointer, non-positive size, bad stride or alignment)";
    case VQB_E_WORKSPACE: return "vqb: workspace too small or misaligned (see vqb_workspace_bytes)";
    case VQB_E_UNSUPPORTED: return "vqb: shape not supported by the requested kernel path";
    case VQB_E_DEVICE: return "vqb: device is not an sm_100 (B200) GPU";
    case VQB_E_DRIVER: return "vqb: CUDA driver entry point unavailable";
    case VQB_E_HOSTCTX: return "vqb: host context misuse";
    default: break;
    }
    if (code > 0)
        return cudaGetErrorString((cudaError_t)code);
    return "vqb: unknown error";
}

int vqb_query(int device, vqb_device_info *out)
{
    if (!out)
        return VQB_E_ARG;
    return device_info(device, out);
}

size_t vqb_workspace_bytes(int k, int d)
{
    if (k <= 0 || d <= 0)
        return 0;
    return ws_layout(k, d).total;
}

int vqb_select_path(int device, int64_t n, int k, int d, int64_t stride_row, int64_t stride_d)
{
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (n <= 0 || k <= 0 || d <= 0)
        return VQB_E_ARG;
    const bool contiguous = stride_d == 1 && stride_row == d;
    if (info.cc_major == 10 && contiguous && (tc_shape_supported(k, d) || tc_chunked_supported(k, d)) && n >= 128 &&
        n < (1ll << 31))
        return (int)VQB_PATH_TC;
    return (int)VQB_PATH_FMA;
}

int vqb_forward(int device, const float *z, int64_t n_outer, int64_t n_inner, int d,
                int64_t stride_outer, int64_t stride_inner, int64_t stride_d,
                const float *codebook, int k, float beta,
                float *zq, int64_t *idx, float *loss, float *perplexity,
                unsigned long long *counts, unsigned long long *stats,
                void *workspace, size_t workspace_bytes, unsigned flags, void *stream)
{
    if (n_outer < 0 || n_inner < 0 || d <= 0 || k <= 0 || !codebook || !workspace)
        return VQB_E_ARG;
    const int64_t n = n_outer * n_inner;
    if (n > 0 && !z)
        return VQB_E_ARG;
    if (zq && !aligned(zq, 16))
        return VQB_E_ARG;
    const WsLayout L = ws_layout(k, d);
    if (workspace_bytes < L.total || !aligned(workspace, 256))
        return VQB_E_WORKSPACE;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    WsHeader *hdr = (WsHeader *)ws;
    unsigned long long *cnt = counts ? counts : (unsigned long long *)(ws + L.off_counts);

    float *ee = (float *)(ws + L.off_ee);
    int *colcnt = (int *)(ws + L.off_colcnt);
    int *colwhich = (int *)(ws + L.off_colwhich);
    const bool keep = (flags & VQB_KEEP_CODEBOOK) != 0;
    if (!keep) {
        // zero: header, (ee), colcnt, colwhich -- one contiguous prefix of the workspace -- then the codebook prep
        if ((err = cudaMemsetAsync(ws, 0, L.off_counts, st)) != cudaSuccess) return (int)err;
        if ((err = launch_prep(codebook, k, d, L.kpad, ee, colcnt, colwhich, hdr, st)) != cudaSuccess) return (int)err;
        count_launches(1);
    }
    if ((err = cudaMemsetAsync(cnt, 0, sizeof(unsigned long long) * k, st)) != cudaSuccess) return (int)err;
    if (stats && (err = cudaMemsetAsync(stats, 0, sizeof(unsigned long long) * 4, st)) != cudaSuccess) return (int)err;

    int n_ctas = 1;
    double *partials = (double *)(ws + L.off_partials);
    if (n > 0) {
        FwdParams p{};
        p.z.base = z; p.z.n_rows = n; p.z.n_inner = n_inner > 0 ? n_inner : 1;
        p.z.s_outer = stride_outer; p.z.s_inner = stride_inner; p.z.s_d = stride_d;
        p.E = codebook; p.K = k; p.D = d; p.ee = ee; p.colcnt = colcnt; p.colwhich = colwhich; p.hdr_in = hdr;
        p.zq = zq; p.idx = idx; p.counts = cnt; p.partials = partials; p.accumulate = 0;
        p.stats = stats;
        p.need_sq = loss != nullptr;

        unsigned path = flags & VQB_PATH_MASK;
        // K > 256 runs as one tcgen05 pass per 256-code chunk; the running best lives in the idx buffer
        const bool tc_chunked = tc_chunked_supported(k, d) && idx != nullptr;
        // (the tcgen05 kernels address rows with 32-bit TMA coordinates: 2^31 rows and more take the FMA path)
        const bool tc_ok = p.z.rows_contiguous(d) && (tc_shape_supported(k, d) || tc_chunked) && aligned(z, 16) &&
                           n < (1ll << 31);
        if (path == VQB_PATH_TC && !tc_ok)
            return VQB_E_UNSUPPORTED;
        const bool was_auto = path == VQB_PATH_AUTO;
        if (was_auto)
            path = (tc_ok && n >= 128) ? VQB_PATH_TC : VQB_PATH_FMA;
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        if (g_profile.load()) {
            if ((err = cudaEventCreate(&ev0)) != cudaSuccess) return (int)err;
            if ((err = cudaEventCreate(&ev1)) != cudaSuccess) {
                cudaEventDestroy(ev0);
                return (int)err;
            }
        }
        int launches = 1;
        if (path == VQB_PATH_TC) {
            launches = 0;
            // the tcgen05 launcher brackets only its main kernel (after its own codebook prep)
            if (tc_shape_supported(k, d))
                err = launch_fwd_tc(p, (float *)(ws + L.off_tc), info.sm_count, info.max_smem_per_block, &n_ctas,
                                    &launches, st, ev0, ev1, keep && (flags & VQB_KEEP_TC_IMAGE) != 0);
            else
                err = launch_fwd_tc_chunked(p, (float *)(ws + L.off_tc), info.sm_count, info.max_smem_per_block, &n_ctas,
                                            &launches, st, ev0, ev1);
            // the launcher could not set the tcgen05 kernel up (tensor-map encode, driver entry point, shared-memory limit):
            // nothing of it has run; a call that left the choice to the library takes the exact CUDA-core kernel instead,
            // one that asked for the tcgen05 path is told so
            if (err == cudaErrorNotSupported) {
                (void)cudaGetLastError();
                if (!was_auto) {
                    if (ev0) {
                        cudaEventDestroy(ev0);
                        cudaEventDestroy(ev1);
                    }
                    return VQB_E_UNSUPPORTED;
                }
                path = VQB_PATH_FMA;
                launches = 1;
            }
        }
        if (path != VQB_PATH_TC) {
            if (ev0) cudaEventRecord(ev0, st);
            err = launch_fwd_fma(p, info.sm_count, info.max_smem_per_block, &n_ctas, st);
            if (ev1) cudaEventRecord(ev1, st);
        }
        if (err != cudaSuccess) {
            if (ev0) {
                cudaEventDestroy(ev0);
                cudaEventDestroy(ev1);
            }
            return (int)err;
        }
        count_launches(launches);
        if (ev0) {
            std::lock_guard<std::mutex> lock(g_prof_mu);
            g_prof_events.push_back(ProfBracket{device, ev0, ev1});
        }
    } else {
        if ((err = cudaMemsetAsync(partials, 0, sizeof(double), st)) != cudaSuccess) return (int)err;
    }
    if (loss || perplexity) {
        err = launch_finalize(cnt, k, partials, n_ctas, (double *)(ws + L.off_sq), 0, n, d, beta, loss, perplexity, st);
        if (err != cudaSuccess) return (int)err;
        count_launches(1);
    }
    return VQB_OK;
}

int vqb_token_linear(int device, const void *a_bf16, const void *w_bf16, const float *bias, float *h, void *out_bf16,
                     int64_t n_tokens, int k, int n, unsigned mode, void *stream)
{
    if (!a_bf16 || !w_bf16 || !bias || n_tokens < 0 || mode > 2u || (mode != 0u && !h) || (mode == 0u && !out_bf16))
        return VQB_E_ARG;
    if (!tok_linear_supported(k, n) || !aligned(a_bf16, 16) || !aligned(w_bf16, 16) || !aligned(bias, 16) ||
        (h && !aligned(h, 16)) || (out_bf16 && !aligned(out_bf16, 16)))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_linear(a_bf16, w_bf16, bias, h, out_bf16, n_tokens, k, n, (int)mode, info.sm_count,
                            info.max_smem_per_block, (cudaStream_t)stream);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_conv(int device, const void *a_bf16, const void *w_bf16, const float *bias, float *h, void *out_bf16,
                   int64_t n_tokens, int k_in, int n, unsigned mode, int taps, int tokens_per_cycle, int out_gelu, void *stream)
{
    if (!a_bf16 || !w_bf16 || !bias || n_tokens < 0 || mode > 2u || (mode != 0u && !h) || (mode == 0u && !out_bf16) ||
        (taps != 1 && taps != 3) || tokens_per_cycle < 1)
        return VQB_E_ARG;
    if (!tok_linear_supported(k_in, n) || !aligned(a_bf16, 16) || !aligned(w_bf16, 16) || !aligned(bias, 16) ||
        (h && !aligned(h, 16)) || (out_bf16 && !aligned(out_bf16, 16)))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_linear(a_bf16, w_bf16, bias, h, out_bf16, n_tokens, k_in, n, (int)mode, info.sm_count,
                            info.max_smem_per_block, (cudaStream_t)stream, taps, tokens_per_cycle, out_gelu ? 1 : 0);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_linear_split(int device, const void *a_pair, const void *w_pair, const float *bias, float *h, void *out_pair,
                           int64_t n_tokens, int k, int n, unsigned mode, int out_gelu, void *stream)
{
    if (!a_pair || !w_pair || !bias || n_tokens < 0 || mode > 2u || (mode != 0u && !h) || (mode == 0u && !out_pair))
        return VQB_E_ARG;
    if (!tok_linear_supported(k, n) || !aligned(a_pair, 16) || !aligned(w_pair, 16) || !aligned(bias, 16) ||
        (h && !aligned(h, 16)) || (out_pair && !aligned(out_pair, 16)))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_linear(a_pair, w_pair, bias, h, out_pair, n_tokens, k, n, (int)mode, info.sm_count,
                            info.max_smem_per_block, (cudaStream_t)stream, 1, 1, out_gelu ? 1 : 0, 1);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_pair(int device, const float *h, void *out_pair, int64_t n_tokens, int n, int apply_gelu, void *stream)
{
    if (!h || !out_pair || n_tokens < 0 || n <= 0)
        return VQB_E_ARG;
    if (n % 4 != 0 || !aligned(h, 16) || !aligned(out_pair, 8))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_pair(h, out_pair, n_tokens, n, apply_gelu ? 1 : 0, info.sm_count, (cudaStream_t)stream);
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_out_proj(int device, const void *a_bf16, const float *w, float bias, float *out, int64_t n_rows, int hidden, int p,
                       void *stream)
{
    if (!a_bf16 || !w || !out || n_rows < 0 || hidden <= 0 || p <= 0)
        return VQB_E_ARG;
    if (!tok_out_proj_supported(hidden, p) || !aligned(a_bf16, 16) || !aligned(w, 4) || !aligned(out, 4))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_out_proj(a_bf16, w, bias, out, n_rows, hidden, p, info.sm_count, (cudaStream_t)stream);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_rows > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_conv_split(int device, const void *a_pair, const void *w_pair, const float *bias, float *h, void *out_pair,
                         int64_t n_tokens, int k_in, int n, unsigned mode, int taps, int tokens_per_cycle, int out_gelu, void *stream)
{
    if (!a_pair || !w_pair || !bias || n_tokens < 0 || mode > 2u || (mode != 0u && !h) || (mode == 0u && !out_pair) ||
        (taps != 1 && taps != 3) || tokens_per_cycle < 1)
        return VQB_E_ARG;
    if (!tok_linear_supported(k_in, n) || !aligned(a_pair, 16) || !aligned(w_pair, 16) || !aligned(bias, 16) ||
        (h && !aligned(h, 16)) || (out_pair && !aligned(out_pair, 16)))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_linear(a_pair, w_pair, bias, h, out_pair, n_tokens, k_in, n, (int)mode, info.sm_count,
                            info.max_smem_per_block, (cudaStream_t)stream, taps, tokens_per_cycle, out_gelu ? 1 : 0, 1);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_out_proj_pair(int device, const void *a_pair, const float *w, float bias, float *out, int64_t n_tokens, int group,
                            int hidden, int p, void *stream)
{
    if (!a_pair || !w || !out || n_tokens < 0 || hidden <= 0 || p <= 0 || group < 1)
        return VQB_E_ARG;
    if (!tok_out_proj_supported(hidden, p) || !aligned(a_pair, 16) || !aligned(w, 4) || !aligned(out, 4))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    const int64_t half = (int64_t)group * hidden;                 // elements of the hi half of a token's row
    const uint16_t *a = (const uint16_t *)a_pair;                 // bf16 elements
    err = launch_tok_out_proj(a, w, bias, out, n_tokens * group, hidden, p, info.sm_count, (cudaStream_t)stream, group, 2 * half, 0);
    if (err == cudaSuccess)
        err = launch_tok_out_proj(a + half, w, 0.0f, out, n_tokens * group, hidden, p, info.sm_count, (cudaStream_t)stream, group,
                                  2 * half, 1);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(2);
    return VQB_OK;
}

size_t vqb_encoder_chain_scratch_bytes(int device, int hidden)
{
    vqb_device_info info;
    if (device_info(device, &info) != VQB_OK || hidden <= 0)
        return 0;
    return enc_chain_scratch_bytes(hidden, info.sm_count);
}

int vqb_encoder_chain(int device, const void *a0_bf16, float *h, const void *w_bf16, const float *bias, int64_t n_tokens,
                      int hidden, int n_layers, void *scratch, size_t scratch_bytes, const float *proj_bias, float *z_e,
                      int proj_dim, const float *pre_bias, void *stream)
{
    if (!a0_bf16 || !w_bf16 || !bias || !scratch || n_tokens < 0 || hidden <= 0 || n_layers <= 0 || proj_dim < 0)
        return VQB_E_ARG;
    if (proj_dim > 0 && (!proj_bias || !z_e))
        return VQB_E_ARG;
    if (!(pre_bias && proj_dim > 0) && !h)          // h is the input without PRE and the output without PROJ
        return VQB_E_ARG;
    if (!enc_chain_supported(hidden, n_layers) || !aligned(a0_bf16, 16) || !aligned(w_bf16, 16) || !aligned(bias, 16) ||
        (h && !aligned(h, 16)) || !aligned(scratch, 16) || (proj_dim > 0 && (!aligned(proj_bias, 4) || !aligned(z_e, 16))) ||
        (pre_bias && !aligned(pre_bias, 4)))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    if (scratch_bytes < enc_chain_scratch_bytes(hidden, info.sm_count))
        return VQB_E_WORKSPACE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_enc_chain(a0_bf16, h, w_bf16, bias, n_tokens, hidden, n_layers, (float *)scratch, scratch_bytes,
                           proj_bias, z_e, proj_dim, pre_bias, info.sm_count, info.max_smem_per_block, (cudaStream_t)stream);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_patch_split(int device, const float *x, int64_t n_cycles, int seq_len, int channels, int patch, void *out_bf16,
                    void *stream)
{
    if (!x || !out_bf16 || n_cycles < 0 || seq_len <= 0 || channels <= 0 || patch <= 0)
        return VQB_E_ARG;
    if (!patch_split_supported(seq_len, channels, patch))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_patch_split(x, out_bf16, n_cycles, seq_len, channels, patch, info.sm_count, (cudaStream_t)stream);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_cycles > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_token_bias_gelu(int device, float *h, const float *bias, void *out_bf16, int64_t n_tokens, int n, void *stream)
{
    if (!h || !bias || !out_bf16 || n_tokens < 0 || n <= 0)
        return VQB_E_ARG;
    if (n % 4 != 0 || !aligned(h, 16) || !aligned(bias, 16) || !aligned(out_bf16, 8))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_tok_bias_gelu(h, bias, out_bf16, n_tokens, n, info.sm_count, (cudaStream_t)stream);
    if (err != cudaSuccess)
        return (int)err;
    if (n_tokens > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_patch_embed(int device, const float *x, int64_t n_cycles, int seq_len, int channels, int patch, const float *w,
                    const float *bias, float *h, void *out_bf16, int hidden, void *stream)
{
    if (!x || !w || !bias || !h || n_cycles < 0 || seq_len <= 0 || channels <= 0 || patch <= 0 || hidden <= 0)
        return VQB_E_ARG;
    if (!patch_embed_supported(seq_len, channels, patch, hidden) || !aligned(h, 16) || (out_bf16 && !aligned(out_bf16, 8)))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_patch_embed(x, w, bias, h, out_bf16, n_cycles, seq_len, channels, patch, hidden, info.sm_count,
                             info.max_smem_per_block, (cudaStream_t)stream);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_cycles > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_pack_rows(int device, const float *z, int64_t n_outer, int64_t n_inner, int d, int64_t stride_outer,
                  int64_t stride_inner, int64_t stride_d, float *out, void *stream)
{
    if (!z || !out || n_outer < 0 || n_inner <= 0 || d <= 0)
        return VQB_E_ARG;
    if (!pack_rows_supported(n_inner, d, stride_outer, stride_inner, stride_d))
        return VQB_E_UNSUPPORTED;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_pack_rows(z, out, n_outer, n_inner, d, stride_outer, stride_inner, stride_d, info.sm_count,
                           (cudaStream_t)stream);
    if (err == cudaErrorNotSupported)
        return VQB_E_UNSUPPORTED;
    if (err != cudaSuccess)
        return (int)err;
    if (n_outer > 0)
        count_launches(1);
    return VQB_OK;
}

int vqb_ar_pairs(int device, const int64_t *ids, int64_t n_windows, int n_tokens, int64_t start_token, int64_t end_token,
                 int64_t *x, int64_t *y, void *stream)
{
    if (n_windows < 0 || n_tokens <= 0 || (n_windows > 0 && (!ids || !x || !y)))
        return VQB_E_ARG;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_ar_pairs(ids, n_windows, n_tokens, start_token, end_token, x, y, info.sm_count, (cudaStream_t)stream);
    if (err != cudaSuccess)
        return (int)err;
    count_launches(n_windows > 0 ? 1 : 0);
    return VQB_OK;
}

int vqb_row_keys(int device, const void *rows, int64_t n_rows, int words, const unsigned long long *mult,
                 unsigned long long *keys, void *stream)
{
    if (n_rows < 0 || words <= 0 || !mult || (n_rows > 0 && (!rows || !keys)) || !aligned(rows, 4) || !aligned(mult, 8) ||
        !aligned(keys, 8))
        return VQB_E_ARG;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_row_keys(rows, n_rows, words, mult, keys, info.sm_count, (cudaStream_t)stream);
    if (err != cudaSuccess)
        return (int)err;
    count_launches(n_rows > 0 ? 1 : 0);
    return VQB_OK;
}

size_t vqb_dedupe_scratch_bytes(int64_t n_rows) { return dedupe_scratch_bytes(n_rows); }

int vqb_dedupe_first(int device, const unsigned long long *keys, int64_t n_rows, void *scratch, size_t scratch_bytes,
                     int64_t *first, void *stream)
{
    if (n_rows < 0 || (n_rows > 0 && (!keys || !scratch || !first)) || !aligned(keys, 8) || !aligned(scratch, 8) ||
        !aligned(first, 8))
        return VQB_E_ARG;
    if (n_rows >= 0x7F000000ll)                  // row indices live in 32-bit table entries
        return VQB_E_UNSUPPORTED;
    if (scratch_bytes < dedupe_scratch_bytes(n_rows))
        return VQB_E_WORKSPACE;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    err = launch_dedupe_first(keys, n_rows, scratch, first, info.sm_count, (cudaStream_t)stream);
    if (err != cudaSuccess)
        return (int)err;
    count_launches(n_rows > 0 ? 2 : 0);
    return VQB_OK;
}

int vqb_backward(int device, const float *g_zq, const float *g_loss,
                 const float *z, int64_t n_outer, int64_t n_inner, int d,
                 int64_t stride_outer, int64_t stride_inner, int64_t stride_d,
                 const int64_t *idx, const float *codebook, int k, float beta,
                 float *grad_z, float *grad_codebook,
                 void *workspace, size_t workspace_bytes, void *stream)
{
    (void)workspace; (void)workspace_bytes;
    if (n_outer < 0 || n_inner < 0 || d <= 0 || k <= 0 || !codebook)
        return VQB_E_ARG;
    const int64_t n = n_outer * n_inner;
    if (n > 0 && (!z || !idx))
        return VQB_E_ARG;
    vqb_device_info info;
    int rc = device_info(device, &info);
    if (rc != VQB_OK)
        return rc;
    if (info.cc_major != 10)
        return VQB_E_DEVICE;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    ZView zv{z, n, n_inner > 0 ? n_inner : 1, stride_outer, stride_inner, stride_d};
    err = launch_bwd(g_zq, g_loss, zv, idx, codebook, k, d, beta, grad_z, grad_codebook, info.sm_count,
                     info.max_smem_per_block, (cudaStream_t)stream);
    if (err == cudaSuccess)
        count_launches((n > 0 ? 1 : 0) + (grad_codebook ? 1 : 0));
    return (int)err;
}

int vqb_gather(int device, const int64_t *idx, int64_t n, const float *codebook, int k, int d,
               float *out, int *bad_index, void *stream)
{
    if (n < 0 || k <= 0 || d <= 0 || !codebook || (n > 0 && (!idx || !out)))
        return VQB_E_ARG;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    if (bad_index && (err = cudaMemsetAsync(bad_index, 0, sizeof(int), (cudaStream_t)stream)) != cudaSuccess)
        return (int)err;
    count_launches(n > 0 ? 1 : 0);
    return (int)launch_gather(idx, n, codebook, k, d, out, bad_index, (cudaStream_t)stream);
}

int vqb_one_hot(int device, const int64_t *idx, int64_t n, int k, float *onehot, void *stream)
{
    if (n < 0 || k <= 0 || (n > 0 && (!idx || !onehot)))
        return VQB_E_ARG;
    cudaError_t err = cudaSetDevice(device);
    if (err != cudaSuccess)
        return (int)err;
    count_launches(n > 0 ? 1 : 0);
    return (int)launch_one_hot(idx, n, k, onehot, (cudaStream_t)stream);
}

}  // extern "C"
