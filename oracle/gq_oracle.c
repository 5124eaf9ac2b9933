/*
 * gq_oracle.c -- CPU restatement of the GaussianQuant hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker (or as the timed CPU baseline).
 * The product path (libgq_b200.so, sm_100a CUDA) never links or calls it.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the
 * restatement is pinned against outputs of the unmodified reference module imported from
 * /root/reference in the build container (tests/golden/make_golden.py wrote
 * tests/golden/NAME.npz; tests/test_oracle_golden.py replays them bit-for-bit).
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout).  All arithmetic is IEEE binary32 with no contraction
 * (-ffp-contract=off) unless a function says otherwise.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define GQ_MAX_G 32

/* math.log(math.sqrt(2*math.pi)) as torch sees it once the python double is cast to
 * the fp32 opmath type (torch/distributions/normal.py log_prob, called from
 * pit/quantization/gaussian.py:141-146). */
static const float C0F = (float)0.9189385332046727;

int gq_oracle_version(void) { return 1; }

/* Use n host threads (n <= 0: every online core).  torchrun exports OMP_NUM_THREADS=1 to its
 * workers; the CPU baseline has to override that to use the whole host. */
void gq_oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n <= 0) n = omp_get_num_procs();
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int gq_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* torch.sum(dim=-1) over a contiguous innermost dim of g <= 32 fp32 values on the CPU
 * (ATen SumKernel vectorized_inner_sum, Vectorized<float>::size() == 8): eight strided
 * lanes p[j] = t[j] + t[j+8] + ..., a scalar tail for g % 8, then
 * acc = 0 + tail... + p[0] + ... + p[7] in that order.  Probed bit-equal against
 * torch 2.11 for g in {4, 8, 16, 32} (see DESIGN.md "summation order").
 * Used by pit/quantization/gaussian.py:147 `torch.sum(log_ratios, dim=2)`. */
static inline float sum_torch_inner(const float *t, int g)
{
    float acc = 0.0f;
    int nv = g / 8;
    if (nv == 0) {
        for (int k = 0; k < g; ++k) acc = acc + t[k];
        /* adding the eight zero partials changes nothing but the sign of -0 */
        acc = acc + 0.0f;
        return acc;
    }
    float p[8];
    for (int j = 0; j < 8; ++j) {
        float a = t[j];
        for (int v = 1; v < nv; ++v) a = a + t[j + 8 * v];
        p[j] = a;
    }
    for (int k = nv * 8; k < g; ++k) acc = acc + t[k];
    for (int j = 0; j < 8; ++j) acc = acc + p[j];
    return acc;
}

/* One (row, codeword) log-ratio in the reference torch op order,
 * pit/quantization/gaussian.py:141-147 with torch.distributions.Normal.log_prob:
 *   t = x - mu ; t = t*t ; t = -t ; t = t / (2*(std*std)) ; t = t - log(std) ;
 *   t = t - c0 ; t = t - (nlp*beta) ; S = sum_k t
 * den[k] = 2*(std*std), lstd[k] = log(std) and nlpb[k] = nlp*beta are the row- and
 * codeword-invariant sub-expressions, each still one fp32 rounding as in torch. */
static inline float score_torch(const float *x, const float *nlpb, const float *mu,
                                const float *den, const float *lstd, int g)
{
    float t[GQ_MAX_G];
    for (int k = 0; k < g; ++k) {
        float d = x[k] - mu[k];
        float d2 = d * d;
        float q = (-d2) / den[k];
        float a = q - lstd[k];
        float b = a - C0F;
        t[k] = b - nlpb[k];
    }
    return sum_torch_inner(t, g);
}

/* Normal(0,1).log_prob(codebook) as the constructor computes it,
 * pit/quantization/gaussian.py:51-52: -((x-0)**2)/(2*1**2) - log(1) - c0. */
void gq_oracle_normal_log_prob(const float *codebook, int64_t n_elems, float *nlp)
{
    for (int64_t i = 0; i < n_elems; ++i) {
        float d = codebook[i] - 0.0f;
        float d2 = d * d;
        float q = (-d2) / 2.0f;
        float a = q - 0.0f;
        nlp[i] = a - C0F;
    }
}

/* Full [M,N] score matrix, torch order (small cases only).  gaussian.py:141-147. */
int gq_oracle_scores_torch(const float *mu, const float *std, const float *log_std,
                           const float *codebook, const float *nlp, int64_t M, int64_t N,
                           int g, float beta, float *out)
{
    if (g < 1 || g > GQ_MAX_G) return -1;
    float *nlpb = (float *)malloc(sizeof(float) * (size_t)N * g);
    if (!nlpb) return -2;
    for (int64_t i = 0; i < N * g; ++i) nlpb[i] = nlp[i] * beta;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < M; ++r) {
        float den[GQ_MAX_G];
        for (int k = 0; k < g; ++k) {
            float s = std[r * g + k];
            den[k] = 2.0f * (s * s);
        }
        for (int64_t n = 0; n < N; ++n)
            out[r * N + n] = score_torch(codebook + n * g, nlpb + n * g, mu + r * g, den,
                                         log_std + r * g, g);
    }
    free(nlpb);
    return 0;
}

/* Eval-branch quantizer, torch backend: gaussian.py:134-150.
 * idx[r] = first argmax_n S[r,n] (torch.argmax tie-break), zhat[r,:] = codebook[idx[r],:].
 * Optional outputs (may be NULL): best[r] = S[r,idx[r]], gap[r] = best - second best
 * (used by the tests to classify near-ties, SURVEY.md 8d "parity gates"). */
int gq_oracle_quantize(const float *mu, const float *std, const float *log_std,
                       const float *codebook, const float *nlp, int64_t M, int64_t N,
                       int g, float beta, int64_t *idx, float *zhat, float *best,
                       float *gap)
{
    if (g < 1 || g > GQ_MAX_G) return -1;
    if (N < 1) return -1;
    float *nlpb = (float *)malloc(sizeof(float) * (size_t)N * g);
    if (!nlpb) return -2;
    for (int64_t i = 0; i < N * g; ++i) nlpb[i] = nlp[i] * beta;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t r = 0; r < M; ++r) {
        float den[GQ_MAX_G];
        for (int k = 0; k < g; ++k) {
            float s = std[r * g + k];
            den[k] = 2.0f * (s * s);
        }
        float b1 = -INFINITY, b2 = -INFINITY;
        int64_t bi = 0;
        int first = 1;
        for (int64_t n = 0; n < N; ++n) {
            float s = score_torch(codebook + n * g, nlpb + n * g, mu + r * g, den,
                                  log_std + r * g, g);
            if (first) { b1 = s; bi = n; first = 0; }
            else if (s > b1) { b2 = b1; b1 = s; bi = n; }
            else if (s > b2) { b2 = s; }
        }
        idx[r] = bi;
        if (zhat) memcpy(zhat + r * g, codebook + bi * g, sizeof(float) * g);
        if (best) best[r] = b1;
        if (gap) gap[r] = b1 - b2;
    }
    free(nlpb);
    return 0;
}

/* Legacy operator semantics: gq_cuda_extension/gq_cuda/csrc/cuda/gq_cuda.cu:31-38 and
 * test/test_extension.py:5-6.  out[r,n] = sum_k( -((x-mu)/std)^2 + x*x*beta ), with beta a
 * double: `co*co` is an fp32 product, promoted to double, multiplied by beta and added to
 * the fp32 running value in double before rounding back.  nvcc's default -fmad=true
 * contracts `log_w -= iv*iv` into fma(-iv, iv, log_w) and the double multiply-add into one
 * fma(); contract != 0 reproduces that, contract == 0 is the uncontracted C semantics. */
int gq_oracle_scores_legacy(const float *mu, const float *std, const float *noise,
                            float *out, int64_t dim, int64_t M, int64_t N, double beta,
                            int contract)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < M; ++r) {
        for (int64_t n = 0; n < N; ++n) {
            float lw = 0.0f;
            for (int64_t i = 0; i < dim; ++i) {
                float co = noise[n * dim + i];
                float iv = (co - mu[r * dim + i]) / std[r * dim + i];
                float cc = co * co;
                if (contract) {
                    lw = fmaf(-iv, iv, lw);
                    lw = (float)fma((double)cc, beta, (double)lw);
                } else {
                    float iv2 = iv * iv;
                    lw = lw - iv2;
                    double p = (double)cc * beta;
                    lw = (float)((double)lw + p);
                }
            }
            out[r * N + n] = lw;
        }
    }
    return 0;
}

/* dequant gather on rows: gaussian.py:171-173 `index_select(prior_samples, 0, idx)`.
 * Returns -3 on an out-of-range index (torch raises IndexError). */
int gq_oracle_gather(const int64_t *idx, const float *codebook, int64_t M, int64_t N, int g,
                     float *out)
{
    for (int64_t r = 0; r < M; ++r) {
        if (idx[r] < 0 || idx[r] >= N) return -3;
        memcpy(out + r * g, codebook + idx[r] * g, sizeof(float) * g);
    }
    return 0;
}

/* Train-branch KL penalty on a [B, L, 2C] (blc) latent: gaussian.py:76-98.
 * Channel c of token (b,l): mu = z[b,l,c], logvar = clamp(z[b,l,C+c], lo, hi).
 * kl2 element = 0.7213 * (mu*mu + exp(logvar) - 1 - logvar)   (1.4426*0.5 is folded by
 * python before it meets the tensor, gaussian.py:86); group j of a token sums channels
 * {k*ncb + j : k < g}  (reshape(b,l,g,ncb).sum(2), gaussian.py:87-88).
 * stats[0..4] = { sum_all(kl2), min, max, sum_all(w*kl2), count } in double;
 * kl_loss = lam * stats[3] / B  is formed by the caller.  kl2_out (may be NULL) receives
 * the [B*L*ncb] fp32 per-group bits.  expf() here is glibc's, not torch's Sleef -- the
 * KL gate is 1e-5 relative (BASELINE.json), not bit-exact. */
int gq_oracle_kl(const float *z, int64_t B, int64_t L, int C, int g, float lv_lo, float lv_hi,
                 float target_bits, float tol, float lam_min, float lam_max, double *stats,
                 float *kl2_out)
{
    if (g < 1 || C % g) return -1;
    int ncb = C / g;
    double s_all = 0.0, s_w = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    for (int64_t t = 0; t < B * L; ++t) {
        const float *row = z + t * 2 * C;
        for (int j = 0; j < ncb; ++j) {
            float acc = 0.0f;
            for (int k = 0; k < g; ++k) {
                int c = k * ncb + j;
                float mu = row[c];
                float lv = row[C + c];
                lv = lv < lv_lo ? lv_lo : (lv > lv_hi ? lv_hi : lv);
                float var = expf(lv);
                float e = mu * mu + var;
                e = e - 1.0f;
                e = e - lv;
                acc = acc + 0.7213f * e;
            }
            float hi = target_bits + tol, lo = target_bits - tol;
            float w = 0.0f;
            if (acc > hi) w += lam_max;
            if (acc <= hi && acc >= lo) w += 1.0f;
            if (acc < lo) w += lam_min;
            s_all += (double)acc;
            s_w += (double)w * (double)acc;
            if (acc < mn) mn = acc;
            if (acc > mx) mx = acc;
            if (kl2_out) kl2_out[t * ncb + j] = acc;
        }
    }
    stats[0] = s_all;
    stats[1] = mn;
    stats[2] = mx;
    stats[3] = s_w;
    stats[4] = (double)(B * L * ncb);
    return 0;
}
