/*
 * gq_b200.h -- C ABI of libgq_b200.so: the B200 (sm_100a) GaussianQuant hot path.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes
 * and a CUDA stream (as void*); there are no torch types in any signature.  All calls
 * are asynchronous on `stream`, never synchronise the device, never allocate device
 * memory (callers pass workspaces sized by the *_workspace_bytes helpers) and return
 * GQ_OK or a negative gq_status; gq_last_error() gives the message for the calling
 * thread.  The caller owns every buffer.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repo
 * tongdaxu/VQ-VAE-from-Gaussian-VAE):
 *
 *   gq_scores_f32        gq_cuda_extension/gq_cuda/csrc/cuda/gq_cuda.cu:77-118  (op
 *                        `extension_cpp::gq`, schema gq_cuda/csrc/gq_cuda.cpp:30,
 *                        python gq_cuda/ops.py:7-8)
 *   gq_codebook_create   pit/quantization/gaussian.py:50-52 (prior_samples / normal_log_prob
 *                        buffers; the codebook bits themselves still come from the host's
 *                        prior_samples(), gaussian.py:15-19)
 *   gq_quantize          pit/quantization/gaussian.py:124-150 (both backends of the eval
 *                        branch: score, argmax, index_select) -- fused, no [M,N] matrix
 *   gq_quantize_exact    same semantics, CUDA-core exact-order scan (anchor / fallback)
 *   gq_dequant           pit/quantization/gaussian.py:162-178
 *   gq_kl_fwd/_bwd       pit/quantization/gaussian.py:76-98 (+ autograd of the same)
 *   gq_index_histogram   eval.py:127,152-154 (all_hist += torch.histogram(indices, range(0, N)))
 *   gq_usage_entropy     eval.py:137-141 (cal_ent: codebook usage and entropy of the histogram)
 *   gq_peer_gather_indices   eval.py:167-201 (per-rank results -> all_gather), as one kernel over NVLink peer memory
 *   gq_pack_indices / gq_unpack_indices / gq_dequant_packed
 *                        no reference counterpart: the reference ships int64 indices
 *                        (gaussian.py:138,159); opt-in 16/32-bit stream, dequant as gaussian.py:162-178
 *
 * Row convention (gaussian.py:122-123): rows M = b*l*ncb, row r = (token)*ncb + j holds
 * channels {k*ncb + j : k < g} of its token.
 */
#ifndef GQ_B200_H
#define GQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GQ_B200_ABI_VERSION 4

typedef void *gq_stream_t; /* cudaStream_t */

typedef enum gq_status {
    GQ_OK = 0,
    GQ_ERR_INVALID = -1,     /* bad argument (shape, null pointer, unsupported group) */
    GQ_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed */
    GQ_ERR_UNSUPPORTED = -3, /* valid request this build cannot serve (e.g. not sm_100) */
    GQ_ERR_WORKSPACE = -4    /* workspace too small */
} gq_status;

/* Search precision of a codebook with group 8 or 16 (other groups ignore it).  FAST: fp16 operand
 * image; rows whose error window is too wide (very small variances) take the exact CUDA-core scan.
 * ROBUST: split-precision image (three fp16 products per term, K = 6g), ~1.4x the search time, windows
 * ~250x tighter.  AUTO (default): FAST until a completed call on this codebook had more than 1/16 of
 * its rows over the window cap, ROBUST while that stays true (the count is computed in both modes and
 * read from pinned host memory without synchronising, so the switch lags by a call or two). */
enum gq_mode { GQ_MODE_AUTO = 0, GQ_MODE_FAST = 1, GQ_MODE_ROBUST = 2 };

/* tensor layouts of the regularizer's latent / index tensors (gaussian.py:33,65-74) */
typedef enum gq_layout {
    GQ_LAYOUT_BLC = 0, /* [b, l, c]      channel innermost */
    GQ_LAYOUT_BCHW = 1 /* [b, c, h, w]   token (h*w) innermost */
} gq_layout;

int gq_abi_version(void);
const char *gq_last_error(void);
/* number of kernels this library has launched in the calling process (bench accounting) */
uint64_t gq_launch_count(void);

/* ---- legacy operator: out[M,N] = sum_k( -((x-mu)/std)^2 + x*x*beta ) ------------------
 * gq_cuda.cu:12-40.  mu,std [M,dim]; codebook [N,dim]; out [M,N]; all fp32 contiguous. */
int gq_scores_f32(const float *mu, const float *std, const float *codebook, float *out,
                  int64_t dim, int64_t M, int64_t N, double beta, gq_stream_t stream);

/* ---- prepared codebook ---------------------------------------------------------------
 * Device-resident, read-only after creation: the fp32 codebook [N,g], beta*normal_log_prob
 * [N,g] and the fp16 [x^2 | x] operand image the tensor-core search streams.  `nlp` may be
 * NULL (then Normal(0,1).log_prob is evaluated as gaussian.py:51-52 does).  The handle owns
 * one device allocation made at creation time (not on the hot path). */
typedef struct gq_codebook gq_codebook_t;
int gq_codebook_create(const float *codebook, const float *nlp, int64_t N, int g, float beta,
                       gq_stream_t stream, gq_codebook_t **out);
void gq_codebook_destroy(gq_codebook_t *cb);
int gq_codebook_set_mode(gq_codebook_t *cb, int mode);
/* hist: device uint64[N] (caller-owned, zeroed by the caller) or NULL.  While set, every gq_quantize_latent
 * on this codebook also accumulates the indices it emits into hist, fused into its last pass -- the
 * reference's `all_hist += torch.histogram(indices, ...)` (eval.py:127,152-154) without another launch. */
int gq_codebook_set_usage_histogram(gq_codebook_t *cb, uint64_t *hist);
int64_t gq_codebook_size(const gq_codebook_t *cb);
int gq_codebook_group(const gq_codebook_t *cb);
const float *gq_codebook_data(const gq_codebook_t *cb); /* device fp32 [N,g] */

/* ---- eval-branch quantizer ------------------------------------------------------------
 * idx[r] = first argmax_n S[r,n], S in the reference torch op order (gaussian.py:141-147);
 * zhat[r,:] = codebook[idx[r],:] (may be NULL).  mu,std,log_std [M,g] fp32 rows
 * (log_std = log(std) as the host framework computes it; NULL => logf in a prologue).
 * gq_quantize: tcgen05 fp16 search with a rigorous per-row error window, exact fp32
 * re-scoring of the surviving candidates, exact-scan fallback for ill-conditioned rows.
 * stats (may be NULL): device uint32[8] = {rows, candidates rescored, fallback rows,
 * overflow rows, 0...}, accumulated (caller zeroes). */
int64_t gq_quantize_workspace_bytes(const gq_codebook_t *cb, int64_t M);
int gq_quantize(const gq_codebook_t *cb, const float *mu, const float *std,
                const float *log_std, int64_t M, int64_t *idx, float *zhat, void *workspace,
                int64_t workspace_bytes, uint32_t *stats, gq_stream_t stream);
int gq_quantize_exact(const gq_codebook_t *cb, const float *mu, const float *std,
                      const float *log_std, int64_t M, int64_t *idx, float *zhat,
                      gq_stream_t stream);

/* ---- fused eval branch from the raw latent ----------------------------------------------
 * The whole eval branch of GaussianQuantRegularizer.forward (gaussian.py:62-80,121-159) in one
 * call: z fp32 [b,l,2C] (BLC) or [b,2C,h,w] (BCHW, l = h*w) -> idx int64 [b,l,ncb] / [b,ncb,h,w],
 * zhat fp32 [b,l,C] / [b,C,h,w], and (optionally) zhat_noquant = mu + eps*std in zhat's layout.
 * eps: fp32 noise in zhat's layout, drawn by the host framework (torch.randn_like keeps RNG parity:
 * in bchw the reference's randn_like(mu) fills memory in [b,C,h,w] order because mu is a view);
 * eps == NULL <=> zhat_noquant == NULL.  std = exp(0.5*clamp(logvar, lv_lo, lv_hi)) and log(std)
 * are evaluated with the libdevice functions torch's own CUDA kernels use. */
int64_t gq_quantize_latent_workspace_bytes(const gq_codebook_t *cb, int64_t b, int64_t l, int C);
int gq_quantize_latent(const gq_codebook_t *cb, const float *z, const float *eps, int64_t b,
                       int64_t l, int C, int layout, float lv_lo, float lv_hi, int64_t *idx,
                       float *zhat, float *zhat_noquant, void *workspace, int64_t workspace_bytes,
                       uint32_t *stats, gq_stream_t stream);

/* ---- dequant --------------------------------------------------------------------------
 * idx: int64, [b,l,ncb] (BLC) or [b,ncb,h,w] (BCHW); out: fp32 [b,l,ncb*g] or
 * [b,ncb*g,h,w]; out channel k*ncb + j of a token = codebook[idx[token,j]][k].
 * err_flag (may be NULL): device int32 set to 1 if any index is outside [0,N). */
int gq_dequant(const gq_codebook_t *cb, const int64_t *idx, float *out, int64_t b, int64_t l,
               int ncb, int layout, int32_t *err_flag, gq_stream_t stream);

/* ---- packed index stream (opt-in; int64 stays the default wire format) -----------------------
 * Width in bytes: gq_index_width(N) = 2 when N <= 65536, else 4 (N < 2^31).  packed: n elements of
 * that width, same element order as idx.  err_flag (may be NULL): set to 1 if an index is outside
 * [0,N) (it is then stored truncated).  gq_dequant_packed == gq_dequant on the unpacked indices. */
int gq_index_width(int64_t N);
int gq_pack_indices(const int64_t *idx, int64_t n, int64_t N, void *packed, int32_t *err_flag,
                    gq_stream_t stream);
int gq_unpack_indices(const void *packed, int64_t n, int64_t N, int64_t *idx, gq_stream_t stream);
int gq_dequant_packed(const gq_codebook_t *cb, const void *packed, float *out, int64_t b, int64_t l,
                      int ncb, int layout, int32_t *err_flag, gq_stream_t stream);

/* ---- index exchange over NVLink peer memory ------------------------------------------------
 * For row-sharded callers that want the whole batch's indices on every rank (the reference gathers
 * per-rank results with all_gather, eval.py:167-201).  ONE kernel converts this rank's n int64
 * indices to `width` bytes each (8 = unchanged, or gq_index_width(N)) and stores them at element
 * offset elem_offset into EVERY rank's gather buffer: peer_bufs is a DEVICE array of `world` device
 * pointers to the ranks' buffers as mapped into this process (e.g. torch symmetric memory's
 * buffer_ptrs_dev); when mc_ptr is not NULL it is an NVSwitch multicast address of the same buffers
 * and full 16-byte groups go out as one multimem.st each (elements outside a 16-byte group still
 * use peer_bufs).  The call does not synchronise ranks: the caller puts its cross-rank barrier
 * after it (and must not overwrite a buffer a peer may still read).  err_flag as gq_pack_indices. */
int gq_peer_gather_indices(const int64_t *idx, int64_t n, int64_t N, int width, void *const *peer_bufs,
                           void *mc_ptr, int world, int64_t elem_offset, int32_t *err_flag,
                           gq_stream_t stream);
/* The same with the cross-rank barrier INSIDE the kernel (one launch per exchange, no collective
 * library): after its stores the grid's last block writes `seq` into flag slot [rank] of every
 * rank's flag block and waits until all `world` slots of its own flag block hold >= seq; kernels
 * that follow on the stream may read the gathered tensor.  The flag block lives at byte offset
 * flag_offset (multiple of 16) of every rank's buffer: 64 uint32 slots + one arrival counter
 * (272 bytes), zeroed once before the first call on all ranks.  seq starts at 1 and grows by one
 * per call, on every rank alike.  n may be 0 (barrier only).  A rank that never makes the matching
 * call is detected after ~4 s of polling: err_flag (may be NULL) is set to 2.
 * gq_peer_kl_fold: global KL statistics in ONE kernel -- this rank's float[8] gq_kl_fwd statistics
 * go to 32-byte slot [rank] at byte offset data_offset of every rank's buffer (peer stores), the
 * same hand-shake follows, and the world-size rows are folded exactly as gq_kl_fold does. */
int gq_peer_gather_indices_sync(const int64_t *idx, int64_t n, int64_t N, int width, void *const *peer_bufs,
                                void *mc_ptr, int world, int rank, int64_t elem_offset, int64_t flag_offset,
                                uint32_t seq, int32_t *err_flag, gq_stream_t stream);
int gq_peer_kl_fold(const float *stats, void *const *peer_bufs, int world, int rank, int64_t data_offset,
                    int64_t flag_offset, uint32_t seq, double lam, double global_batch, double total_groups,
                    float *out4, int32_t *err_flag, gq_stream_t stream);

/* ---- codebook usage ------------------------------------------------------------------------
 * gq_index_histogram: hist[idx[i]] += 1 for i < n (hist: device uint64[N], ACCUMULATED -- zero it
 * once, call per batch, like the reference's all_hist); out-of-range indices are skipped and
 * flagged.  gq_usage_entropy: out = device float[2] = {usage, entropy_bits} with
 * usage = 1 - (#empty bins)/N, p = hist/sum(hist), entropy = -sum p*log2(p + 1e-5), fp32. */
int gq_index_histogram(const int64_t *idx, int64_t n, int64_t N, uint64_t *hist, int32_t *err_flag,
                       gq_stream_t stream);
int gq_usage_entropy(const uint64_t *hist, int64_t N, float *out, gq_stream_t stream);

/* ---- train-branch KL penalty ------------------------------------------------------------
 * z: fp32 [b,l,2C] or [b,2C,h,w]; eps: fp32 noise shaped like mu (torch.randn_like, kept on
 * the host framework's RNG); zhat out = mu + eps*exp(0.5*clamp(logvar)) in z's layout.
 * out_stats: device float[8] = {kl_loss, bits_mean, bits_min, bits_max, sum_kl2, sum_w_kl2,
 * 0, 0}; kl_loss = lam * sum(w*kl2) / b.  w_out: fp32 [b*l*ncb] band weight per group
 * (lam_max / 1 / lam_min), kept for backward.  workspace: gq_kl_workspace_bytes. */
int64_t gq_kl_workspace_bytes(int64_t b, int64_t l, int C);
int gq_kl_fwd(const float *z, const float *eps, float *zhat, float *w_out, float *out_stats,
              int64_t b, int64_t l, int C, int g, int layout, float lv_lo, float lv_hi,
              float target_bits, float tol, float lam, float lam_min, float lam_max,
              void *workspace, int64_t workspace_bytes, gq_stream_t stream);
/* grad_z = d(kl_loss)/dz * grad_kl + d(zhat)/dz . grad_zhat ; grad_kl: device scalar
 * (may be NULL => 0), grad_zhat may be NULL => 0. */
int gq_kl_bwd(const float *z, const float *eps, const float *w, const float *grad_zhat,
              const float *grad_kl, float *grad_z, int64_t b, int64_t l, int C, int g,
              int layout, float lv_lo, float lv_hi, float lam, gq_stream_t stream);

/* Fold the gq_kl_fwd statistics of `world` row shards (all_stats: world x 8 floats, the out_stats of every rank
 * as an all-gather delivers them) into the unsharded batch's [kl_loss, bits_mean, bits_min, bits_max] -> out4:
 * kl_loss = lam * sum(w*kl2) / global_batch, bits_mean = sum(kl2) / total_groups (gaussian.py:89-98 over the
 * global batch), sums in double in rank order (every rank computes the same bits).  One tiny kernel: with the
 * all-gather it is the whole KL exchange of a sharded training step. */
int gq_kl_fold(const float *all_stats, int world, double lam, double global_batch, double total_groups,
               float *out4, gq_stream_t stream);

/* ---- measurement and diagnostics (no reference counterpart) ------------------------------
 * gq_profile_enable(n): every n-th later gq_quantize (n = 1: every one; 0: off) records a CUDA-event pair around its search
 * kernel on the caller's stream; gq_profile_read sums and clears them (synchronises on the
 * recorded events).  gq_set_window_max: rows whose rigorous error window exceeds `w` (32 `w` with the split-precision image; score
 * units) go to the exact scan (default 3.0).  gq_debug_tc_scores dumps the raw tensor-core
 * score matrix [M_pad, n_pad] and the scaled per-row windows [M_pad]; gq_debug_plan returns
 * the stream-K decomposition {units, tiles, ctas, steps_per_cta, nseg, M_pad, unit_rows, tile_cols} of the fp32-accumulator
 * kernel (acc16 = 0: 256-row units, 128-codeword tiles) or the fp16-accumulator kernel (1: 128-row units, 256-codeword tiles). */
void gq_profile_enable(int on);
int gq_profile_read(double *total_ms, int *count);
void gq_set_window_max(float w);
/* The kernels of one gq_quantize* call are chained with programmatic dependent launch (each starts
 * with griddepcontrol.wait); gq_set_chained_launch(0) falls back to plain stream order (default 1). */
void gq_set_chained_launch(int on);
int gq_debug_tc_scores(const gq_codebook_t *cb, const float *mu, const float *std,
                       const float *log_std, int64_t M, void *workspace, int64_t workspace_bytes,
                       float *dump, float *wrow_out, gq_stream_t stream);
void gq_debug_plan(int64_t N, int64_t M, int sms, int acc16, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* GQ_B200_H */
