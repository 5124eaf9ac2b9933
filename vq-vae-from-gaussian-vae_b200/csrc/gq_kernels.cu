// gq_kernels.cu -- CUDA-core kernels of the GaussianQuant path (sm_100a):
//   * legacy score matrix            (gq_cuda.cu:12-40 semantics)
//   * exact-order quantizer scan     (gaussian.py:134-150 semantics; anchor + fallback)
//   * codebook preparation           (gaussian.py:50-52 + fp16 operand image)
//   * dequant gather                 (gaussian.py:162-178)
// The tensor-core search lives in gq_tc.cu, the KL penalty in gq_kl.cu.
#include "gq_common.cuh"
#include "gq_internal.h"

#include <cub/device/device_radix_sort.cuh>

namespace gq {

// ------------------------------------------------------------------------------------------
// Legacy operator.  One thread per codeword n, looping over a tile of rows; out[r, n] is
// written coalesced along n.  Arithmetic reproduces the reference kernel as nvcc compiles it
// (-fmad=true): iv = (x-mu)/std (IEEE divide), lw = fma(-iv, iv, lw),
// lw = float(fma(double(x*x), beta, double(lw))).  For beta == 1 the double step equals one
// fp32 add exactly (innocuous double rounding, 53 >= 2*24+2), so it is done in fp32.
// ------------------------------------------------------------------------------------------
constexpr int kLegacyRows = 32;

template <bool kBetaOne>
__global__ void __launch_bounds__(256)
scores_legacy_kernel(const float *__restrict__ mu, const float *__restrict__ std,
                     const float *__restrict__ cb, float *__restrict__ out, int dim, int64_t M,
                     int64_t N, double beta)
{
    extern __shared__ float smem[];  // [kLegacyRows][dim] mu, then std
    float *s_mu = smem;
    float *s_sd = smem + kLegacyRows * dim;
    const int64_t r0 = (int64_t)blockIdx.y * kLegacyRows;
    const int rows = (int)min((int64_t)kLegacyRows, M - r0);
    for (int i = threadIdx.x; i < rows * dim; i += blockDim.x) {
        s_mu[i] = mu[r0 * dim + i];
        s_sd[i] = std[r0 * dim + i];
    }
    __syncthreads();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float *x = cb + n * dim;
    for (int r = 0; r < rows; ++r) {
        float lw = 0.0f;
        for (int i = 0; i < dim; ++i) {
            float co = __ldg(x + i);
            float iv = __fdiv_rn(__fsub_rn(co, s_mu[r * dim + i]), s_sd[r * dim + i]);
            float cc = __fmul_rn(co, co);
            lw = __fmaf_rn(-iv, iv, lw);
            if (kBetaOne) lw = __fadd_rn(lw, cc);
            else lw = (float)__fma_rn((double)cc, beta, (double)lw);
        }
        out[(r0 + r) * N + n] = lw;
    }
}

int launch_scores_legacy(const float *mu, const float *std, const float *cb, float *out,
                         int64_t dim, int64_t M, int64_t N, double beta, cudaStream_t stream)
{
    if (M == 0 || N == 0) return GQ_OK;
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)((M + kLegacyRows - 1) / kLegacyRows));
    if (grid.y > 65535u) {
        // split rows so gridDim.y stays legal
        int64_t chunk = (int64_t)65535 * kLegacyRows;
        for (int64_t r = 0; r < M; r += chunk) {
            int64_t m = min(chunk, M - r);
            int rc = launch_scores_legacy(mu + r * dim, std + r * dim, cb, out + r * N, dim, m, N,
                                          beta, stream);
            if (rc) return rc;
        }
        return GQ_OK;
    }
    size_t smem = sizeof(float) * 2 * kLegacyRows * (size_t)dim;
    if (beta == 1.0)
        scores_legacy_kernel<true><<<grid, 256, smem, stream>>>(mu, std, cb, out, (int)dim, M, N, beta);
    else
        scores_legacy_kernel<false><<<grid, 256, smem, stream>>>(mu, std, cb, out, (int)dim, M, N, beta);
    count_launch();
    return check_launch("scores_legacy_kernel");
}

// ------------------------------------------------------------------------------------------
// Exact-order scan.  A block owns kExactRows rows; every thread walks codewords n = tid,
// tid+256, ... with the codeword in registers and scores it against the block's rows in the
// reference op order.  Strict '>' per thread keeps the lowest n; the block reduction breaks
// ties toward the lower index (torch.argmax returns the first maximum).
// row_list/row_count: optional indirection (fallback rows chosen on the device).
// ------------------------------------------------------------------------------------------
constexpr int kExactRows = 8;
constexpr int kExactThreads = 256;

template <int G>
__global__ void __launch_bounds__(kExactThreads)
quantize_exact_kernel(const float *__restrict__ mu, const float *__restrict__ std,
                      const float *__restrict__ lstd, const float *__restrict__ cb,
                      const float *__restrict__ nlpb, int64_t N, const uint32_t *__restrict__ row_list,
                      const uint32_t *__restrict__ row_count, int64_t M, int64_t *__restrict__ idx_out,
                      float *__restrict__ zhat_out)
{
    __shared__ float s_mu[kExactRows][G], s_den[kExactRows][G], s_ls[kExactRows][G];
    __shared__ int64_t s_row[kExactRows];
    __shared__ float s_best[kExactRows][kExactThreads / 32];
    __shared__ int s_bidx[kExactRows][kExactThreads / 32];

    const int64_t total = row_count ? (int64_t)(*row_count) : M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int64_t base = (int64_t)blockIdx.x * kExactRows; base < total;
         base += (int64_t)gridDim.x * kExactRows) {
        const int rows = (int)min((int64_t)kExactRows, total - base);
        __syncthreads();
        if (tid < kExactRows) {
            int64_t r = -1;
            if (tid < rows) r = row_list ? (int64_t)row_list[base + tid] : base + tid;
            s_row[tid] = r;
        }
        __syncthreads();
        for (int i = tid; i < kExactRows * G; i += kExactThreads) {
            int rr = i / G, k = i % G;
            int64_t r = s_row[rr];
            float m = 0.f, s = 1.f, l = 0.f;
            if (r >= 0) {
                m = mu[r * G + k];
                s = std[r * G + k];
                l = lstd ? lstd[r * G + k] : logf(s);
            }
            s_mu[rr][k] = m;
            s_den[rr][k] = __fmul_rn(2.0f, __fmul_rn(s, s));
            s_ls[rr][k] = l;
        }
        __syncthreads();

        float best[kExactRows];
        int bidx[kExactRows];
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) { best[r] = -CUDART_INF_F; bidx[r] = 0x7fffffff; }

        for (int64_t n = tid; n < N; n += kExactThreads) {
            float x[G], nb[G];
            load_row<G>(cb + n * G, x);
            load_row<G>(nlpb + n * G, nb);
#pragma unroll
            for (int r = 0; r < kExactRows; ++r) {
                if (r >= rows) break;   // (block-uniform) a short row group does not pay for the empty slots
                float t[G];
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    float d = __fsub_rn(x[k], s_mu[r][k]);
                    float d2 = __fmul_rn(d, d);
                    float q = __fdiv_rn(-d2, s_den[r][k]);
                    float a = __fsub_rn(q, s_ls[r][k]);
                    float b = __fsub_rn(a, GQ_C0F);
                    t[k] = __fsub_rn(b, nb[k]);
                }
                float s = sum_torch_inner<G>(t);
                if (s > best[r]) { best[r] = s; bidx[r] = (int)n; }
            }
        }
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) {
            float b = best[r];
            int bi = bidx[r];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                float ob = __shfl_xor_sync(0xffffffffu, b, off);
                int oi = __shfl_xor_sync(0xffffffffu, bi, off);
                if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
            }
            if (lane == 0) { s_best[r][warp] = b; s_bidx[r][warp] = bi; }
        }
        __syncthreads();
        if (tid < rows) {
            float b = s_best[tid][0];
            int bi = s_bidx[tid][0];
            for (int w = 1; w < kExactThreads / 32; ++w) {
                float ob = s_best[tid][w];
                int oi = s_bidx[tid][w];
                if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
            }
            if (bi == 0x7fffffff) bi = 0;  // every score was -inf/NaN: torch.argmax gives 0
            int64_t r = s_row[tid];
            idx_out[r] = (int64_t)bi;
            if (zhat_out) {
#pragma unroll
                for (int k = 0; k < G; ++k) zhat_out[r * G + k] = cb[(int64_t)bi * G + k];
            }
        }
    }
}

template <int G>
static int launch_exact_g(const gq_codebook *cb, const float *mu, const float *std, const float *lstd,
                          int64_t M, const uint32_t *row_list, const uint32_t *row_count,
                          int64_t *idx, float *zhat, int max_blocks, cudaStream_t stream)
{
    int64_t need = (M + kExactRows - 1) / kExactRows;
    int blocks = (int)imin(need, max_blocks);
    if (blocks <= 0) return GQ_OK;
    quantize_exact_kernel<G><<<blocks, kExactThreads, 0, stream>>>(
        mu, std, lstd, cb->cb, cb->nlpb, cb->N, row_list, row_count, M, idx, zhat);
    count_launch();
    return check_launch("quantize_exact_kernel");
}

int launch_quantize_exact(const gq_codebook *cb, const float *mu, const float *std,
                          const float *lstd, int64_t M, const uint32_t *row_list,
                          const uint32_t *row_count, int64_t *idx, float *zhat, int max_blocks,
                          cudaStream_t stream)
{
    switch (cb->g) {
    case 1: return launch_exact_g<1>(cb, mu, std, lstd, M, row_list, row_count, idx, zhat, max_blocks, stream);
    case 2: return launch_exact_g<2>(cb, mu, std, lstd, M, row_list, row_count, idx, zhat, max_blocks, stream);
    case 4: return launch_exact_g<4>(cb, mu, std, lstd, M, row_list, row_count, idx, zhat, max_blocks, stream);
    case 8: return launch_exact_g<8>(cb, mu, std, lstd, M, row_list, row_count, idx, zhat, max_blocks, stream);
    case 16: return launch_exact_g<16>(cb, mu, std, lstd, M, row_list, row_count, idx, zhat, max_blocks, stream);
    case 32: return launch_exact_g<32>(cb, mu, std, lstd, M, row_list, row_count, idx, zhat, max_blocks, stream);
    default:
        set_error("group %d unsupported (supported: 1,2,4,8,16,32)", cb->g);
        return GQ_ERR_INVALID;
    }
}

// ------------------------------------------------------------------------------------------
// Sliced exact scan: the fallback of the tensor-core search.  Work item = (group of kExactRows
// listed rows, slice of the codebook); a block scores its slice in the reference op order and merges
// its per-row winners into packed 64-bit keys with atomicMax:
//   key = monotone_u32(score) << 32 | (0xffffffff - index)      (higher score wins, ties -> lower index)
// The block that completes a row group (arrival counter) turns the keys into idx / zhat.  A handful of
// fallback rows thus spreads over the whole GPU instead of serialising 65536 codewords in one block.
// keys[r] must be 0 when the row is listed (the listing kernels do that); done[] is zeroed per call.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_key(float s, int n)
{
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)n);
}

template <int G>
__global__ void __launch_bounds__(kExactThreads)
exact_sliced_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
                    const float *__restrict__ cb, const float *__restrict__ nlpb, int64_t N,
                    const uint32_t *__restrict__ row_list, const uint32_t *__restrict__ row_count, int slices,
                    unsigned long long *__restrict__ keys, uint32_t *__restrict__ done,
                    int64_t *__restrict__ idx_out, float *__restrict__ zhat_out, uint32_t *__restrict__ wide_dev,
                    uint32_t *__restrict__ fb_dev, uint32_t rows_total, const FinalLayout fin)
{
    pdl_wait();   // launched as a programmatic dependent of settle_kernel (gq_tc.cu run_tc)
    // Last kernel of a latent call whose rows are not already in final form: row indices -> the module's output
    // tensors (what a separate finalize launch did).  Rows on this kernel's own list carry idx = -1 and are written
    // when their scan completes, below; the two sets are disjoint, so nothing here waits for anything.
    if (fin.idx_final) {
        const int64_t total_el = fin.b * fin.l * fin.ncb;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_el; i += (int64_t)gridDim.x * blockDim.x) {
            int64_t bi, t;
            int j;
            if (fin.layout == GQ_LAYOUT_BCHW) { t = i % fin.l; j = (int)((i / fin.l) % fin.ncb); bi = i / (fin.l * fin.ncb); }
            else { j = (int)(i % fin.ncb); t = (i / fin.ncb) % fin.l; bi = i / (fin.l * fin.ncb); }
            const int64_t id = idx_out[(bi * fin.l + t) * fin.ncb + j];
            if (id >= 0) final_store_row<G>(fin.idx_final, fin.zhat_final, fin.l, fin.ncb, fin.layout, fin.usage_hist, cb, bi, t, j, id);
        }
    }
    // last kernel of the call: publish how many of its rows had a plain-image window over the cap
    // (GQ_MODE_AUTO feedback, read by the host without synchronising) and reset the device word
    if (fb_dev && blockIdx.x == 0 && threadIdx.x == 0) {
        fb_dev[0] = atomicExch(wide_dev, 0u);
        fb_dev[1] = rows_total;
        __threadfence_system();
    }
    __shared__ float s_mu[kExactRows][G], s_den[kExactRows][G], s_ls[kExactRows][G];
    __shared__ int64_t s_row[kExactRows];
    __shared__ unsigned long long s_key[kExactRows][kExactThreads / 32];
    __shared__ bool s_last;
    const int64_t total = (int64_t)(*row_count);
    const int64_t groups = (total + kExactRows - 1) / kExactRows;
    // A handful of rows (the usual case: one row of a 65 536-row batch missed its certificate) would leave most of the
    // grid idle with the host's slice count: cut the codebook finer so that every block gets a piece (down to one
    // codeword per thread).  Every block derives the same number from the same device-side row count.
    if (groups > 0)
        slices = (int)imin(imax((int64_t)slices, (int64_t)gridDim.x / groups), (N + kExactThreads - 1) / kExactThreads);
    const int64_t per = ((N + slices - 1) / slices + kExactThreads - 1) / kExactThreads * kExactThreads;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int64_t item = blockIdx.x; item < groups * slices; item += gridDim.x) {
        const int64_t grp = item / slices;
        const int sl = (int)(item % slices);
        const int64_t base = grp * kExactRows;
        const int rows = (int)imin((int64_t)kExactRows, total - base);
        __syncthreads();
        if (tid < kExactRows) s_row[tid] = tid < rows ? (int64_t)row_list[base + tid] : -1;
        __syncthreads();
        for (int i = tid; i < kExactRows * G; i += kExactThreads) {
            const int rr = i / G, k = i % G;
            const int64_t r = s_row[rr];
            float m = 0.f, sd = 1.f, l = 0.f;
            if (r >= 0) {
                m = mu[r * G + k];
                sd = std[r * G + k];
                l = lstd ? lstd[r * G + k] : logf(sd);
            }
            s_mu[rr][k] = m;
            s_den[rr][k] = __fmul_rn(2.0f, __fmul_rn(sd, sd));
            s_ls[rr][k] = l;
        }
        __syncthreads();
        unsigned long long best[kExactRows];
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) best[r] = 0ull;
        const int64_t lo = (int64_t)sl * per, hi = imin(N, lo + per);
        for (int64_t n = lo + tid; n < hi; n += kExactThreads) {
            float x[G], nb[G];
            load_row<G>(cb + n * G, x);
            load_row<G>(nlpb + n * G, nb);
#pragma unroll
            for (int r = 0; r < kExactRows; ++r) {
                if (r >= rows) break;   // (block-uniform) a short row group does not pay for the empty slots
                float t[G];
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    float d = __fsub_rn(x[k], s_mu[r][k]);
                    float d2 = __fmul_rn(d, d);
                    float q = __fdiv_rn(-d2, s_den[r][k]);
                    float a = __fsub_rn(q, s_ls[r][k]);
                    float b = __fsub_rn(a, GQ_C0F);
                    t[k] = __fsub_rn(b, nb[k]);
                }
                const float sc = sum_torch_inner<G>(t);
                if (sc == sc) {   // NaN never wins here; torch.argmax would pick it, documented divergence
                    const unsigned long long key = pack_key(sc, (int)n);
                    if (key > best[r]) best[r] = key;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kExactRows; ++r) {
            unsigned long long b = best[r];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(0xffffffffu, b, off);
                if (o > b) b = o;
            }
            if (lane == 0) s_key[r][warp] = b;
        }
        __syncthreads();
        if (tid < rows) {
            unsigned long long b = s_key[tid][0];
            for (int w = 1; w < kExactThreads / 32; ++w) if (s_key[tid][w] > b) b = s_key[tid][w];
            if (b) atomicMax(&keys[s_row[tid]], b);
        }
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&done[grp], 1u) == (uint32_t)(slices - 1));
        __syncthreads();
        if (s_last && tid < rows) {
            __threadfence();
            const int64_t r = s_row[tid];
            const unsigned long long key = atomicMax(&keys[r], 0ull);   // coherent read
            const int64_t bi = key ? (int64_t)(0xffffffffu - (uint32_t)(key & 0xffffffffull)) : 0;  // all NaN: torch gives 0
            if (fin.idx_final) {
                // (idx_out[r] stays -1: the pass at the head of the kernel must never pick this row up as well)
                const int64_t tok = r / fin.ncb;
                final_store_row<G>(fin.idx_final, fin.zhat_final, fin.l, fin.ncb, fin.layout, fin.usage_hist, cb, tok / fin.l,
                                   tok % fin.l, (int)(r % fin.ncb), bi);
            } else {
                idx_out[r] = bi;
                if (zhat_out) {
#pragma unroll
                    for (int k = 0; k < G; ++k) zhat_out[r * G + k] = cb[bi * G + k];
                }
            }
        }
    }
}

int launch_exact_sliced(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                        const uint32_t *row_list, const uint32_t *row_count, unsigned long long *keys,
                        uint32_t *done, int64_t *idx, float *zhat, cudaStream_t stream, const FinalLayout *fin)
{
    const int slices = (int)imax(1, imin(64, cb->N / 1024));
    const int blocks = device_sm_count() * 4;
    const FinalLayout fl = fin ? *fin : FinalLayout{nullptr, nullptr, 0, 0, 1, GQ_LAYOUT_BLC, nullptr};
#define GQ_XS(GG) launch_chained(exact_sliced_kernel<GG>, dim3(blocks), dim3(kExactThreads), 0, stream, mu, std, lstd, cb->cb, \
        cb->nlpb, cb->N, row_list, row_count, slices, keys, done, idx, zhat, cb->wide_dev, cb->fb_dev, \
        (uint32_t)imin(M, (int64_t)0xffffffffll), fl)
    switch (cb->g) {
    case 1: GQ_XS(1); break;
    case 2: GQ_XS(2); break;
    case 4: GQ_XS(4); break;
    case 8: GQ_XS(8); break;
    case 16: GQ_XS(16); break;
    case 32: GQ_XS(32); break;
    default: set_error("group %d unsupported", cb->g); return GQ_ERR_INVALID;
    }
#undef GQ_XS
    count_launch();
    return check_launch("exact_sliced_kernel");
}

// ------------------------------------------------------------------------------------------
// Codebook preparation (once per codebook, off the hot path).
//   nlpb[n,k] = nlp[n,k]*beta (one fp32 rounding, gaussian.py:143-144); nlp itself, when not
//   supplied, is Normal(0,1).log_prob(x) in torch's op order (gaussian.py:51-52).
//   bimg: fp16 [x^2 | x | 0-pad] rows of K = kpad (split-precision variant for g <= 4, see gq_common.cuh),
//   stored in UMMA K-major no-swizzle
//   core-matrix order: byte(n,k) = (n/8)*(kpad/8)*128 + (k/8)*128 + (n%8)*16 + (k%8)*2.
//   Rows n >= N (tile padding) are zero except k==2g slot... they are masked in the epilogue.
//   meta[0] = max_n ||x_n||^2 , meta[1] = sqrt(meta[0]), meta[2+j] = max_n |B[n,j]|,
//   meta[2+2*kMaxG] = sum_j meta[2+j].
// ------------------------------------------------------------------------------------------
__global__ void codebook_prep_kernel(const float *__restrict__ cb, const float *__restrict__ nlp_in,
                                     float *__restrict__ nlpb, __half *__restrict__ bimg,
                                     float *__restrict__ meta, int64_t N, int64_t n_pad, int g, int kpad,
                                     float beta, int split_image, int primary, int permute)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_pad) return;
    const int kch = kpad / 8;
    // fp16-accumulator images keep codeword n at image row gq_acc16_image_row(n) (gq_common.cuh)
    const int64_t ir = permute ? gq_acc16_image_row(n) : n;
    char *rowbase = reinterpret_cast<char *>(bimg) + (ir / 8) * (int64_t)kch * 128 + (ir % 8) * 16;
    float nrm2 = 0.f;
    const bool split = split_image != 0;
    for (int k = 0; k < kpad; ++k) {
        // column k of the image: plain image [x^2 | x | 0]; split image [hi(x^2|x) | lo(x^2|x) | hi(x^2|x) | 0],
        // paired with the row image [hi | hi | lo] (gq_tc.cu prep_row), lo = fp16(v - fp16(v))
        const int part = split ? k / (2 * g) : 0;
        const int j = split ? k % (2 * g) : k;
        float v = 0.f;
        if (n < N && (split ? part < 3 : k < 2 * g)) {
            if (j < g) { float x = cb[n * g + j]; v = __fmul_rn(x, x); if (part == 0) nrm2 += v; }
            else v = cb[n * g + (j - g)];
        }
        __half h = __float2half_rn(v);
        if (split && part == 1) h = __float2half_rn(__fsub_rn(v, __half2float(h)));
        *reinterpret_cast<__half *>(rowbase + (k / 8) * 128 + (k % 8) * 2) = h;
        if (primary && n < N && k < 2 * g) {
            // max over the *rounded* operand magnitudes and the exact ones
            float a = fmaxf(fabsf(v), fabsf(__half2float(h)));
            atomicMax(reinterpret_cast<int *>(meta + 2 + k), __float_as_int(a));
        }
    }
    if (primary && n < N) {
        atomicMax(reinterpret_cast<int *>(meta + 0), __float_as_int(nrm2));
        for (int k = 0; k < g; ++k) {
            float x = cb[n * g + k];
            float lp;
            if (nlp_in) lp = nlp_in[n * g + k];
            else {
                float d = __fsub_rn(x, 0.0f);
                float q = __fdiv_rn(-__fmul_rn(d, d), 2.0f);
                lp = __fsub_rn(__fsub_rn(q, 0.0f), GQ_C0F);
            }
            nlpb[n * g + k] = __fmul_rn(lp, beta);
        }
    }
}

__global__ void codebook_meta_finish_kernel(float *meta, int g)
{
    meta[1] = sqrtf(meta[0]) * 1.0000002f;
    float s = 0.f;
    for (int j = 0; j < 2 * g; ++j) s += meta[2 + j];
    meta[2 + 2 * kMaxG] = s * 1.000001f;
}

__global__ void column_keys_kernel(const float *__restrict__ cb, int64_t N, int g, int k, float *__restrict__ keys,
                                   uint32_t *__restrict__ vals)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) { keys[n] = cb[n * g + k]; vals[n] = (uint32_t)n; }
}

// sort_val / sort_idx: every codebook column sorted ascending with its codeword ids (cub radix sort, once per
// codebook, off the hot path; the scratch buffers are freed again)
static int build_sorted_columns(gq_codebook *cb, cudaStream_t stream)
{
    const int64_t N = cb->N;
    float *keys = nullptr;
    uint32_t *vals = nullptr;
    void *tmp = nullptr;
    size_t tmp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, cb->sort_val, vals, cb->sort_idx, (int)N, 0, 32, stream);
    if (e == cudaSuccess) e = cudaMalloc(&keys, (size_t)N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&vals, (size_t)N * 4);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16);
    for (int k = 0; k < cb->g && e == cudaSuccess; ++k) {
        column_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(cb->cb, N, cb->g, k, keys, vals);
        e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, cb->sort_val + (int64_t)k * N, vals, cb->sort_idx + (int64_t)k * N,
                                            (int)N, 0, 32, stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(keys);
    cudaFree(vals);
    cudaFree(tmp);
    if (e != cudaSuccess) { set_error("codebook column sort: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    return GQ_OK;
}

int launch_codebook_prep(gq_codebook *cb, const float *nlp_in, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(cb->meta, 0, sizeof(float) * (3 + 2 * kMaxG), stream);
    if (e != cudaSuccess) { set_error("memset meta: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    int64_t blocks = (cb->n_pad2 + 255) / 256;
    codebook_prep_kernel<<<(unsigned)blocks, 256, 0, stream>>>(cb->cb, nlp_in, cb->nlpb, cb->bimg, cb->meta,
                                                               cb->N, cb->n_pad2, cb->g, cb->kpad, cb->beta,
                                                               gq_split_image(cb->g) ? 1 : 0, 1,
                                                               gq_acc16(cb->g, gq_split_image(cb->g)) ? 1 : 0);
    codebook_meta_finish_kernel<<<1, 1, 0, stream>>>(cb->meta, cb->g);
    count_launch(2);
    if (cb->kpad_robust) {   // second, split-precision image for the robust mode (nlpb / meta already written)
        codebook_prep_kernel<<<(unsigned)blocks, 256, 0, stream>>>(cb->cb, nlp_in, cb->nlpb, cb->bimg_robust, cb->meta,
                                                                   cb->N, cb->n_pad2, cb->g, cb->kpad_robust, cb->beta, 1, 0, 0);
        count_launch(1);
    }
    int rc = check_launch("codebook_prep_kernel");
    if (rc) return rc;
    return build_sorted_columns(cb, stream);
}

// ------------------------------------------------------------------------------------------
// dequant gather (gaussian.py:162-178), fused with the (b l ncb g)->(b l (g ncb)) permute and
// the blc->bchw rearrange.  One thread per (token, sub-codebook j).
//   BLC : idx[(b*l + t)*ncb + j]          out[(b*l + t)*C + k*ncb + j]
//   BCHW: idx[(b*ncb + j)*l + t]          out[(b*C + k*ncb + j)*l + t]
// In BCHW consecutive threads take consecutive tokens, so both the index reads and the g
// strided stores are coalesced; in BLC with ncb == 1 the row is stored as float4s.
// ------------------------------------------------------------------------------------------
template <int G, typename IdxT>
__global__ void __launch_bounds__(256)
dequant_kernel(const IdxT *__restrict__ idx, const float *__restrict__ cb, float *__restrict__ out,
               int64_t b, int64_t l, int ncb, int64_t N, int layout, int32_t *err_flag)
{
    const int64_t total = b * l * ncb;
    const int C = ncb * G;
    // Warm the L2 with the codebook as one linear sweep (128-byte lines) before the random gather:
    // a cold table otherwise costs one scattered 64-byte DRAM access per row.
    if (total >= N) {
        const int64_t lines = (N * G * 4 + 127) / 128;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (int64_t)gridDim.x * blockDim.x)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(cb + i * 32));
    }
    if (layout == GQ_LAYOUT_BLC && ncb == 1 && G % 4 == 0) {
        // G/4 lanes per row, one 16-byte quarter each: loads and stores are both lane-contiguous
        constexpr int Q = (G >= 4) ? G / 4 : 1;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total * Q; e += (int64_t)gridDim.x * blockDim.x) {
            int64_t id = (int64_t)__ldg(idx + e / Q);
            if (id < 0 || id >= N) {
                if (err_flag) *err_flag = 1;
                id = 0;
            }
            reinterpret_cast<float4 *>(out)[e] = __ldg(reinterpret_cast<const float4 *>(cb) + id * Q + (e % Q));
        }
        return;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t bi, t;
        int j;
        if (layout == GQ_LAYOUT_BCHW) {
            t = i % l;
            j = (int)((i / l) % ncb);
            bi = i / (l * ncb);
        } else {
            j = (int)(i % ncb);
            t = (i / ncb) % l;
            bi = i / (l * ncb);
        }
        int64_t id = (int64_t)idx[i];
        if (id < 0 || id >= N) {
            if (err_flag) *err_flag = 1;
            id = 0;
        }
        float v[G];
        load_row<G>(cb + id * G, v);
        if (layout == GQ_LAYOUT_BCHW) {
            float *o = out + (bi * C + j) * l + t;
#pragma unroll
            for (int k = 0; k < G; ++k) o[(int64_t)k * ncb * l] = v[k];
        } else {
            float *o = out + (bi * l + t) * C + j;
            if (ncb == 1 && G % 4 == 0) {
#pragma unroll
                for (int k = 0; k < G / 4; ++k)
                    reinterpret_cast<float4 *>(o)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < G; ++k) o[k * ncb] = v[k];
            }
        }
    }
}

template <typename IdxT>
static int launch_dequant_t(const gq_codebook *cb, const IdxT *idx, float *out, int64_t b, int64_t l, int ncb,
                            int layout, int32_t *err_flag, cudaStream_t stream)
{
    const int64_t total = b * l * ncb;
    if (total == 0) return GQ_OK;
    int64_t items = total;
    if (layout == GQ_LAYOUT_BLC && ncb == 1 && cb->g % 4 == 0) items = total * (cb->g / 4);
    int blocks = (int)imin((items + 255) / 256, 148 * 16);
#define GQ_DQ(GG) dequant_kernel<GG, IdxT><<<blocks, 256, 0, stream>>>(idx, cb->cb, out, b, l, ncb, cb->N, layout, err_flag)
    switch (cb->g) {
    case 1: GQ_DQ(1); break;
    case 2: GQ_DQ(2); break;
    case 4: GQ_DQ(4); break;
    case 8: GQ_DQ(8); break;
    case 16: GQ_DQ(16); break;
    case 32: GQ_DQ(32); break;
    default: set_error("group %d unsupported", cb->g); return GQ_ERR_INVALID;
    }
#undef GQ_DQ
    count_launch();
    return check_launch("dequant_kernel");
}

// idx_bytes: 8 = int64 (the reference's wire format), 4 / 2 = packed stream (gq_tokens.cu)
int launch_dequant(const gq_codebook *cb, const void *idx, int idx_bytes, float *out, int64_t b, int64_t l, int ncb,
                   int layout, int32_t *err_flag, cudaStream_t stream)
{
    switch (idx_bytes) {
    case 8: return launch_dequant_t(cb, static_cast<const int64_t *>(idx), out, b, l, ncb, layout, err_flag, stream);
    case 4: return launch_dequant_t(cb, static_cast<const uint32_t *>(idx), out, b, l, ncb, layout, err_flag, stream);
    case 2: return launch_dequant_t(cb, static_cast<const uint16_t *>(idx), out, b, l, ncb, layout, err_flag, stream);
    default: set_error("index width %d unsupported", idx_bytes); return GQ_ERR_INVALID;
    }
}

}  // namespace gq
