// gq_api.cu -- the extern "C" surface of libgq_b200.so (see include/gq_b200.h).
// Argument validation mirrors the TORCH_CHECKs of the reference wrapper
// (gq_cuda_extension/gq_cuda/csrc/cuda/gq_cuda.cu:91-101) at the pointer/size level; dtype and
// device checks live in the Python shim where tensors still carry that information.
#include "gq_common.cuh"
#include "gq_internal.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace gq {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s launch failed: %s", what, cudaGetErrorString(e));
        return GQ_ERR_CUDA;
    }
    return GQ_OK;
}

static int g_sm_count = -1, g_cc_major = -1, g_dev = -1;
static void query_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { g_sm_count = 148; g_cc_major = 0; return; }
    if (dev == g_dev && g_sm_count > 0) return;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { g_sm_count = 148; g_cc_major = 0; return; }
    g_dev = dev;
    g_sm_count = prop.multiProcessorCount;
    g_cc_major = prop.major;
}
int device_sm_count() { query_device(); return g_sm_count; }
int device_cc_major() { query_device(); return g_cc_major; }

void set_tc_window_max(float w);

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

}  // namespace gq

using namespace gq;

namespace gq { bool g_chained_launch = true; }

extern "C" {

int gq_abi_version(void) { return GQ_B200_ABI_VERSION; }
void gq_set_chained_launch(int on) { gq::g_chained_launch = on != 0; }
const char *gq_last_error(void) { return g_err; }
uint64_t gq_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

/* tuning hook (not part of the reference-facing surface): upper limit of the per-row error
 * window above which a row is routed to the exact scan. */
void gq_set_window_max(float w) { set_tc_window_max(w); }

void gq_profile_enable(int on) { prof_enable(on); }
int gq_profile_read(double *total_ms, int *count) { return prof_read(total_ms, count); }

/* diagnostics used by tests/ and scripts/ only */
int gq_debug_tc_scores(const gq_codebook_t *cb, const float *mu, const float *std, const float *log_std, int64_t M,
                       void *workspace, int64_t workspace_bytes, float *dump, float *wrow_out, gq_stream_t stream)
{
    if (!cb || !mu || !std || !dump || !wrow_out || M <= 0) { set_error("gq_debug_tc_scores: bad argument"); return GQ_ERR_INVALID; }
    return debug_tc_scores(cb, mu, std, log_std, M, workspace, workspace_bytes, dump, wrow_out, (cudaStream_t)stream);
}
void gq_debug_plan(int64_t N, int64_t M, int sms, int acc16, int64_t *out) { debug_plan(N, M, sms, acc16, out); }

int gq_scores_f32(const float *mu, const float *std, const float *codebook, float *out, int64_t dim,
                  int64_t M, int64_t N, double beta, gq_stream_t stream)
{
    if (M < 0 || N < 0 || dim <= 0) { set_error("gq_scores_f32: bad sizes dim=%lld M=%lld N=%lld", (long long)dim, (long long)M, (long long)N); return GQ_ERR_INVALID; }
    if (M == 0 || N == 0) return GQ_OK;
    if (!mu || !std || !codebook || !out) { set_error("gq_scores_f32: null pointer"); return GQ_ERR_INVALID; }
    if (dim > 1024) { set_error("gq_scores_f32: dim %lld too large", (long long)dim); return GQ_ERR_INVALID; }
    return launch_scores_legacy(mu, std, codebook, out, dim, M, N, beta, (cudaStream_t)stream);
}

int gq_codebook_create(const float *codebook, const float *nlp, int64_t N, int g, float beta,
                       gq_stream_t stream, gq_codebook_t **out)
{
    if (!out) { set_error("gq_codebook_create: out is null"); return GQ_ERR_INVALID; }
    *out = nullptr;
    if (!codebook || N <= 0 || N >= ((int64_t)1 << 31)) { set_error("gq_codebook_create: bad codebook (N=%lld)", (long long)N); return GQ_ERR_INVALID; }
    if (!(g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32)) {
        set_error("gq_codebook_create: group %d unsupported (supported: 1,2,4,8,16,32)", g);
        return GQ_ERR_INVALID;
    }
    gq_codebook *cb = new gq_codebook();
    cb->N = N;
    cb->g = g;
    cb->kpad = gq_image_kpad(g);
    cb->n_pad = align_up(N, 128);
    cb->n_pad2 = align_up(N, 256);
    cb->beta = beta;
    cudaGetDevice(&cb->device);
    int64_t off_cb = 0;
    int64_t off_nlpb = align_up(off_cb + N * g * 4, 256);
    int64_t off_bimg = align_up(off_nlpb + N * g * 4, 256);
    cb->kpad_robust = gq_has_robust_image(g) ? gq_image_kpad(g, true) : 0;
    cb->mode = GQ_MODE_AUTO;
    cb->usage_hist = nullptr;
    int64_t off_bimg2 = align_up(off_bimg + cb->n_pad2 * cb->kpad * 2, 256);
    int64_t off_meta = align_up(off_bimg2 + cb->n_pad2 * cb->kpad_robust * 2, 256);
    int64_t off_wide = align_up(off_meta + 4 * (3 + 2 * kMaxG), 256);
    int64_t off_sortv = off_wide + 256;
    int64_t off_sorti = align_up(off_sortv + N * g * 4, 256);
    int64_t total = align_up(off_sorti + N * g * 4, 256);
    cudaError_t e = cudaMalloc(&cb->base, (size_t)total);
    if (e != cudaSuccess) { set_error("gq_codebook_create: cudaMalloc(%lld): %s", (long long)total, cudaGetErrorString(e)); delete cb; return GQ_ERR_CUDA; }
    char *b = reinterpret_cast<char *>(cb->base);
    cb->cb = reinterpret_cast<float *>(b + off_cb);
    cb->nlpb = reinterpret_cast<float *>(b + off_nlpb);
    cb->bimg = reinterpret_cast<__half *>(b + off_bimg);
    cb->meta = reinterpret_cast<float *>(b + off_meta);
    cb->bimg_robust = cb->kpad_robust ? reinterpret_cast<__half *>(b + off_bimg2) : nullptr;
    cb->wide_dev = reinterpret_cast<uint32_t *>(b + off_wide);
    cb->sort_val = reinterpret_cast<float *>(b + off_sortv);
    cb->sort_idx = reinterpret_cast<uint32_t *>(b + off_sorti);
    // feedback word for GQ_MODE_AUTO: pinned + mapped, written by the last kernel of a call, read by the
    // host when it sets up the next one (never synchronised on)
    cb->fb_host = nullptr;
    cb->fb_dev = nullptr;
    {
        void *h = nullptr, *d = nullptr;
        if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
            memset(h, 0, 64);
            cb->fb_host = static_cast<volatile uint32_t *>(h);
            cb->fb_dev = static_cast<uint32_t *>(d);
        } else {
            if (h) cudaFreeHost(h);
            (void)cudaGetLastError();   // no feedback: AUTO behaves like FAST
        }
    }
    cudaMemsetAsync(cb->wide_dev, 0, 256, (cudaStream_t)stream);
    e = cudaMemcpyAsync(cb->cb, codebook, (size_t)(N * g * 4), cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("gq_codebook_create: copy: %s", cudaGetErrorString(e)); cudaFree(cb->base); delete cb; return GQ_ERR_CUDA; }
    int rc = launch_codebook_prep(cb, nlp, (cudaStream_t)stream);
    if (rc) { cudaFree(cb->base); delete cb; return rc; }
    *out = cb;
    return GQ_OK;
}

void gq_codebook_destroy(gq_codebook_t *cb)
{
    if (!cb) return;
    cudaFree(cb->base);
    if (cb->fb_host) cudaFreeHost(const_cast<uint32_t *>(cb->fb_host));
    delete cb;
}

int gq_codebook_set_usage_histogram(gq_codebook_t *cb, uint64_t *hist)
{
    if (!cb) { set_error("gq_codebook_set_usage_histogram: codebook handle is null"); return GQ_ERR_INVALID; }
    cb->usage_hist = reinterpret_cast<unsigned long long *>(hist);
    return GQ_OK;
}

int gq_codebook_set_mode(gq_codebook_t *cb, int mode)
{
    if (!cb) { set_error("gq_codebook_set_mode: codebook handle is null"); return GQ_ERR_INVALID; }
    if (mode != GQ_MODE_AUTO && mode != GQ_MODE_FAST && mode != GQ_MODE_ROBUST) { set_error("gq_codebook_set_mode: bad mode %d", mode); return GQ_ERR_INVALID; }
    cb->mode = mode;
    return GQ_OK;
}

int64_t gq_codebook_size(const gq_codebook_t *cb) { return cb ? cb->N : 0; }
int gq_codebook_group(const gq_codebook_t *cb) { return cb ? cb->g : 0; }
const float *gq_codebook_data(const gq_codebook_t *cb) { return cb ? cb->cb : nullptr; }

int64_t gq_quantize_workspace_bytes(const gq_codebook_t *cb, int64_t M)
{
    if (!cb || M < 0) return -1;
    return tc_workspace_bytes(cb, M);
}

static int check_rows(const char *fn, const gq_codebook_t *cb, const float *mu, const float *std, int64_t M,
                      const int64_t *idx)
{
    if (!cb) { set_error("%s: codebook handle is null", fn); return GQ_ERR_INVALID; }
    if (M < 0) { set_error("%s: M=%lld", fn, (long long)M); return GQ_ERR_INVALID; }
    if (M > 0 && (!mu || !std || !idx)) { set_error("%s: null pointer", fn); return GQ_ERR_INVALID; }
    return GQ_OK;
}

int gq_quantize(const gq_codebook_t *cb, const float *mu, const float *std, const float *log_std, int64_t M,
                int64_t *idx, float *zhat, void *workspace, int64_t workspace_bytes, uint32_t *stats,
                gq_stream_t stream)
{
    int rc = check_rows("gq_quantize", cb, mu, std, M, idx);
    if (rc) return rc;
    return launch_quantize_tc(cb, mu, std, log_std, M, idx, zhat, workspace, workspace_bytes, stats,
                              (cudaStream_t)stream);
}

int64_t gq_quantize_latent_workspace_bytes(const gq_codebook_t *cb, int64_t b, int64_t l, int C)
{
    if (!cb || b < 0 || l < 0 || C <= 0 || C % cb->g) return -1;
    return latent_workspace_bytes(cb, b * l * (C / cb->g));
}

int gq_quantize_latent(const gq_codebook_t *cb, const float *z, const float *eps, int64_t b, int64_t l, int C,
                       int layout, float lv_lo, float lv_hi, int64_t *idx, float *zhat, float *zhat_noquant,
                       void *workspace, int64_t workspace_bytes, uint32_t *stats, gq_stream_t stream)
{
    if (!cb) { set_error("gq_quantize_latent: codebook handle is null"); return GQ_ERR_INVALID; }
    if (b < 0 || l < 0 || C <= 0) { set_error("gq_quantize_latent: bad sizes"); return GQ_ERR_INVALID; }
    if (layout != GQ_LAYOUT_BLC && layout != GQ_LAYOUT_BCHW) { set_error("gq_quantize_latent: bad layout %d", layout); return GQ_ERR_INVALID; }
    if (b * l == 0) return GQ_OK;
    if (!z || !idx || !zhat) { set_error("gq_quantize_latent: null pointer"); return GQ_ERR_INVALID; }
    if ((eps == nullptr) != (zhat_noquant == nullptr)) { set_error("gq_quantize_latent: eps and zhat_noquant go together"); return GQ_ERR_INVALID; }
    return launch_quantize_latent(cb, z, eps, b, l, C, layout, lv_lo, lv_hi, idx, zhat, zhat_noquant, workspace,
                                  workspace_bytes, stats, (cudaStream_t)stream);
}

int gq_quantize_exact(const gq_codebook_t *cb, const float *mu, const float *std, const float *log_std,
                      int64_t M, int64_t *idx, float *zhat, gq_stream_t stream)
{
    int rc = check_rows("gq_quantize_exact", cb, mu, std, M, idx);
    if (rc) return rc;
    if (M == 0) return GQ_OK;
    return launch_quantize_exact(cb, mu, std, log_std, M, nullptr, nullptr, idx, zhat, 1 << 30,
                                 (cudaStream_t)stream);
}

int gq_dequant(const gq_codebook_t *cb, const int64_t *idx, float *out, int64_t b, int64_t l, int ncb,
               int layout, int32_t *err_flag, gq_stream_t stream)
{
    if (!cb) { set_error("gq_dequant: codebook handle is null"); return GQ_ERR_INVALID; }
    if (b < 0 || l < 0 || ncb <= 0) { set_error("gq_dequant: bad sizes"); return GQ_ERR_INVALID; }
    if (layout != GQ_LAYOUT_BLC && layout != GQ_LAYOUT_BCHW) { set_error("gq_dequant: bad layout %d", layout); return GQ_ERR_INVALID; }
    if (b * l == 0) return GQ_OK;
    if (!idx || !out) { set_error("gq_dequant: null pointer"); return GQ_ERR_INVALID; }
    return launch_dequant(cb, idx, 8, out, b, l, ncb, layout, err_flag, (cudaStream_t)stream);
}

int gq_index_width(int64_t N) { return index_width(N); }

int gq_dequant_packed(const gq_codebook_t *cb, const void *packed, float *out, int64_t b, int64_t l, int ncb,
                      int layout, int32_t *err_flag, gq_stream_t stream)
{
    if (!cb) { set_error("gq_dequant_packed: codebook handle is null"); return GQ_ERR_INVALID; }
    if (b < 0 || l < 0 || ncb <= 0) { set_error("gq_dequant_packed: bad sizes"); return GQ_ERR_INVALID; }
    if (layout != GQ_LAYOUT_BLC && layout != GQ_LAYOUT_BCHW) { set_error("gq_dequant_packed: bad layout %d", layout); return GQ_ERR_INVALID; }
    if (b * l == 0) return GQ_OK;
    if (!packed || !out) { set_error("gq_dequant_packed: null pointer"); return GQ_ERR_INVALID; }
    return launch_dequant(cb, packed, index_width(cb->N), out, b, l, ncb, layout, err_flag, (cudaStream_t)stream);
}

int gq_pack_indices(const int64_t *idx, int64_t n, int64_t N, void *packed, int32_t *err_flag, gq_stream_t stream)
{
    if (n < 0 || N <= 0 || N > 0x7fffffffll) { set_error("gq_pack_indices: bad sizes n=%lld N=%lld", (long long)n, (long long)N); return GQ_ERR_INVALID; }
    if (n == 0) return GQ_OK;
    if (!idx || !packed) { set_error("gq_pack_indices: null pointer"); return GQ_ERR_INVALID; }
    return launch_pack_indices(idx, n, N, packed, err_flag, (cudaStream_t)stream);
}

int gq_unpack_indices(const void *packed, int64_t n, int64_t N, int64_t *idx, gq_stream_t stream)
{
    if (n < 0 || N <= 0 || N > 0x7fffffffll) { set_error("gq_unpack_indices: bad sizes n=%lld N=%lld", (long long)n, (long long)N); return GQ_ERR_INVALID; }
    if (n == 0) return GQ_OK;
    if (!idx || !packed) { set_error("gq_unpack_indices: null pointer"); return GQ_ERR_INVALID; }
    return launch_unpack_indices(packed, n, N, idx, (cudaStream_t)stream);
}

int gq_peer_gather_indices(const int64_t *idx, int64_t n, int64_t N, int width, void *const *peer_bufs, void *mc_ptr,
                           int world, int64_t elem_offset, int32_t *err_flag, gq_stream_t stream)
{
    if (n < 0 || N <= 0 || N > 0x7fffffffll || elem_offset < 0) { set_error("gq_peer_gather_indices: bad sizes n=%lld N=%lld offset=%lld", (long long)n, (long long)N, (long long)elem_offset); return GQ_ERR_INVALID; }
    if (width != 8 && width != index_width(N)) { set_error("gq_peer_gather_indices: width %d is neither 8 nor gq_index_width(N) = %d", width, index_width(N)); return GQ_ERR_INVALID; }
    if (world < 1 || world > 64) { set_error("gq_peer_gather_indices: bad world size %d", world); return GQ_ERR_INVALID; }
    if (n == 0) return GQ_OK;
    if (!idx || !peer_bufs) { set_error("gq_peer_gather_indices: null pointer"); return GQ_ERR_INVALID; }
    return launch_peer_gather(idx, n, N, width, peer_bufs, mc_ptr, world, elem_offset, err_flag, (cudaStream_t)stream);
}

static int check_peer_sync(const char *fn, int world, int rank, int64_t flag_offset, uint32_t seq)
{
    if (world < 1 || world > 64 || rank < 0 || rank >= world) { set_error("%s: bad rank/world %d/%d", fn, rank, world); return GQ_ERR_INVALID; }
    if (flag_offset < 0 || (flag_offset & 15)) { set_error("%s: flag_offset must be a non-negative multiple of 16", fn); return GQ_ERR_INVALID; }
    if (seq == 0u) { set_error("%s: seq must be >= 1 and grow by one per call", fn); return GQ_ERR_INVALID; }
    return GQ_OK;
}

int gq_peer_gather_indices_sync(const int64_t *idx, int64_t n, int64_t N, int width, void *const *peer_bufs, void *mc_ptr,
                                int world, int rank, int64_t elem_offset, int64_t flag_offset, uint32_t seq,
                                int32_t *err_flag, gq_stream_t stream)
{
    if (n < 0 || N <= 0 || N > 0x7fffffffll || elem_offset < 0) { set_error("gq_peer_gather_indices_sync: bad sizes n=%lld N=%lld offset=%lld", (long long)n, (long long)N, (long long)elem_offset); return GQ_ERR_INVALID; }
    if (width != 8 && width != index_width(N)) { set_error("gq_peer_gather_indices_sync: width %d is neither 8 nor gq_index_width(N) = %d", width, index_width(N)); return GQ_ERR_INVALID; }
    int rc = check_peer_sync("gq_peer_gather_indices_sync", world, rank, flag_offset, seq);
    if (rc) return rc;
    if (!peer_bufs || (n > 0 && !idx)) { set_error("gq_peer_gather_indices_sync: null pointer"); return GQ_ERR_INVALID; }
    return launch_peer_gather(idx, n, N, width, peer_bufs, mc_ptr, world, elem_offset, err_flag, (cudaStream_t)stream, rank, flag_offset, seq);
}

int gq_peer_kl_fold(const float *stats, void *const *peer_bufs, int world, int rank, int64_t data_offset, int64_t flag_offset,
                    uint32_t seq, double lam, double global_batch, double total_groups, float *out4, int32_t *err_flag,
                    gq_stream_t stream)
{
    int rc = check_peer_sync("gq_peer_kl_fold", world, rank, flag_offset, seq);
    if (rc) return rc;
    if (data_offset < 0 || (data_offset & 15)) { set_error("gq_peer_kl_fold: data_offset must be a non-negative multiple of 16"); return GQ_ERR_INVALID; }
    if (!stats || !peer_bufs || !out4) { set_error("gq_peer_kl_fold: null pointer"); return GQ_ERR_INVALID; }
    if (!(global_batch > 0.0) || !(total_groups > 0.0)) { set_error("gq_peer_kl_fold: global_batch and total_groups must be positive"); return GQ_ERR_INVALID; }
    return launch_peer_kl_fold(stats, peer_bufs, world, rank, data_offset, flag_offset, seq, lam, global_batch,
                               total_groups, out4, err_flag, (cudaStream_t)stream);
}

int gq_index_histogram(const int64_t *idx, int64_t n, int64_t N, uint64_t *hist, int32_t *err_flag, gq_stream_t stream)
{
    if (n < 0 || N <= 0) { set_error("gq_index_histogram: bad sizes n=%lld N=%lld", (long long)n, (long long)N); return GQ_ERR_INVALID; }
    if (!hist) { set_error("gq_index_histogram: hist is null"); return GQ_ERR_INVALID; }
    if (n == 0) return GQ_OK;
    if (!idx) { set_error("gq_index_histogram: idx is null"); return GQ_ERR_INVALID; }
    return launch_index_histogram(idx, n, N, reinterpret_cast<unsigned long long *>(hist), err_flag, (cudaStream_t)stream);
}

int gq_usage_entropy(const uint64_t *hist, int64_t N, float *out, gq_stream_t stream)
{
    if (N <= 0) { set_error("gq_usage_entropy: bad size N=%lld", (long long)N); return GQ_ERR_INVALID; }
    if (!hist || !out) { set_error("gq_usage_entropy: null pointer"); return GQ_ERR_INVALID; }
    return launch_usage_entropy(reinterpret_cast<const unsigned long long *>(hist), N, out, (cudaStream_t)stream);
}

int64_t gq_kl_workspace_bytes(int64_t b, int64_t l, int C) { return kl_workspace_bytes(b, l, C); }

static int check_kl(const char *fn, const float *z, int64_t b, int64_t l, int C, int g, int layout)
{
    if (!z) { set_error("%s: z is null", fn); return GQ_ERR_INVALID; }
    if (b <= 0 || l <= 0 || C <= 0 || g <= 0 || C % g) { set_error("%s: bad sizes b=%lld l=%lld C=%d g=%d", fn, (long long)b, (long long)l, C, g); return GQ_ERR_INVALID; }
    if (layout != GQ_LAYOUT_BLC && layout != GQ_LAYOUT_BCHW) { set_error("%s: bad layout %d", fn, layout); return GQ_ERR_INVALID; }
    return GQ_OK;
}

int gq_kl_fwd(const float *z, const float *eps, float *zhat, float *w_out, float *out_stats, int64_t b,
              int64_t l, int C, int g, int layout, float lv_lo, float lv_hi, float target_bits, float tol,
              float lam, float lam_min, float lam_max, void *workspace, int64_t workspace_bytes,
              gq_stream_t stream)
{
    int rc = check_kl("gq_kl_fwd", z, b, l, C, g, layout);
    if (rc) return rc;
    if (!out_stats || !workspace) { set_error("gq_kl_fwd: null pointer"); return GQ_ERR_INVALID; }
    return launch_kl_fwd(z, eps, zhat, w_out, out_stats, b, l, C, g, layout, lv_lo, lv_hi, target_bits, tol, lam,
                         lam_min, lam_max, workspace, workspace_bytes, (cudaStream_t)stream);
}

int gq_kl_bwd(const float *z, const float *eps, const float *w, const float *grad_zhat, const float *grad_kl,
              float *grad_z, int64_t b, int64_t l, int C, int g, int layout, float lv_lo, float lv_hi, float lam,
              gq_stream_t stream)
{
    int rc = check_kl("gq_kl_bwd", z, b, l, C, g, layout);
    if (rc) return rc;
    if (!grad_z) { set_error("gq_kl_bwd: grad_z is null"); return GQ_ERR_INVALID; }
    return launch_kl_bwd(z, eps, w, grad_zhat, grad_kl, grad_z, b, l, C, g, layout, lv_lo, lv_hi, lam,
                         (cudaStream_t)stream);
}

int gq_kl_fold(const float *all_stats, int world, double lam, double global_batch, double total_groups, float *out4,
               gq_stream_t stream)
{
    if (!all_stats || !out4 || world <= 0 || !(global_batch > 0.0) || !(total_groups > 0.0)) { set_error("gq_kl_fold: bad argument"); return GQ_ERR_INVALID; }
    return launch_kl_fold(all_stats, world, lam, global_batch, total_groups, out4, (cudaStream_t)stream);
}

}  // extern "C"
