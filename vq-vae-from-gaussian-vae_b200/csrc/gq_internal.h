// gq_internal.h -- launch-level prototypes shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct gq_codebook;

namespace gq {

int launch_scores_legacy(const float *mu, const float *std, const float *cb, float *out,
                         int64_t dim, int64_t M, int64_t N, double beta, cudaStream_t stream);

int launch_quantize_exact(const gq_codebook *cb, const float *mu, const float *std,
                          const float *lstd, int64_t M, const uint32_t *row_list,
                          const uint32_t *row_count, int64_t *idx, float *zhat, int max_blocks,
                          cudaStream_t stream);

// The module's output tensors (gaussian.py:153-158) for callers whose rows are not already in final form: indices
// [b,l,ncb] / [b,ncb,h,w] and zhat [b,l,C] / [b,C,h,w] with channel k*ncb + j = codebook[idx][k] of row
// r = (bi*l + t)*ncb + j.  idx_final == nullptr: nothing to do.
struct FinalLayout {
    int64_t *idx_final;
    float *zhat_final;
    int64_t b, l;
    int ncb, layout;
    unsigned long long *usage_hist;   // optional codebook-usage histogram (eval.py:152-154), fused
};

// `fin` (optional): the kernel also turns the row-form indices idx[] into the final tensors -- every row whose idx is
// >= 0 in a pass of its own at the start, every listed row (idx = -1, left so) when its scan completes.
int launch_exact_sliced(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                        const uint32_t *row_list, const uint32_t *row_count, unsigned long long *keys,
                        uint32_t *done, int64_t *idx, float *zhat, cudaStream_t stream, const FinalLayout *fin = nullptr);

int launch_codebook_prep(gq_codebook *cb, const float *nlp_in, cudaStream_t stream);

int launch_dequant(const gq_codebook *cb, const void *idx, int idx_bytes, float *out, int64_t b, int64_t l,
                   int ncb, int layout, int32_t *err_flag, cudaStream_t stream);

// index stream utilities (gq_tokens.cu)
int launch_index_histogram(const int64_t *idx, int64_t n, int64_t N, unsigned long long *hist, int32_t *err_flag,
                           cudaStream_t stream);
int launch_usage_entropy(const unsigned long long *hist, int64_t N, float *out, cudaStream_t stream);
int index_width(int64_t N);
int launch_pack_indices(const int64_t *idx, int64_t n, int64_t N, void *packed, int32_t *err_flag, cudaStream_t stream);
int launch_unpack_indices(const void *packed, int64_t n, int64_t N, int64_t *idx, cudaStream_t stream);
int launch_peer_gather(const int64_t *idx, int64_t n, int64_t N, int width, void *const *peer_bufs, void *mc_ptr, int world,
                       int64_t elem_offset, int32_t *err_flag, cudaStream_t stream, int rank = 0, int64_t flag_off = 0, uint32_t seq = 0);
int launch_peer_kl_fold(const float *stats, void *const *peer_bufs, int world, int rank, int64_t data_off, int64_t flag_off,
                        uint32_t seq, double lam, double global_batch, double total_groups, float *out4, int32_t *err_flag,
                        cudaStream_t stream);

// tensor-core search (gq_tc.cu)
int64_t tc_workspace_bytes(const gq_codebook *cb, int64_t M);
int launch_quantize_tc(const gq_codebook *cb, const float *mu, const float *std, const float *lstd,
                       int64_t M, int64_t *idx, float *zhat, void *ws, int64_t ws_bytes,
                       uint32_t *stats, cudaStream_t stream);

int64_t latent_workspace_bytes(const gq_codebook *cb, int64_t M);
int launch_quantize_latent(const gq_codebook *cb, const float *z, const float *eps, int64_t b, int64_t l, int C,
                           int layout, float lv_lo, float lv_hi, int64_t *idx_out, float *zhat_out, float *zhat_nq,
                           void *ws, int64_t ws_bytes, uint32_t *stats, cudaStream_t stream);

int debug_tc_scores(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                    void *ws, int64_t ws_bytes, float *dump, float *wrow_out, cudaStream_t stream);
void debug_plan(int64_t N, int64_t M, int sms, int acc16, int64_t *out);

void prof_enable(int on);
int prof_read(double *total_ms, int *count);

// KL penalty (gq_kl.cu)
int64_t kl_workspace_bytes(int64_t b, int64_t l, int C);
int launch_kl_fwd(const float *z, const float *eps, float *zhat, float *w_out, float *out_stats,
                  int64_t b, int64_t l, int C, int g, int layout, float lv_lo, float lv_hi,
                  float target_bits, float tol, float lam, float lam_min, float lam_max, void *ws,
                  int64_t ws_bytes, cudaStream_t stream);
int launch_kl_bwd(const float *z, const float *eps, const float *w, const float *grad_zhat,
                  const float *grad_kl, float *grad_z, int64_t b, int64_t l, int C, int g, int layout,
                  float lv_lo, float lv_hi, float lam, cudaStream_t stream);

int launch_kl_fold(const float *all_stats, int world, double lam, double global_batch, double total_groups, float *out4,
                   cudaStream_t stream);

int device_sm_count();
int device_cc_major();

}  // namespace gq
