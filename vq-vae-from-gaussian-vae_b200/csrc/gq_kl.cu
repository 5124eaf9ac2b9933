// gq_kl.cu -- train-branch KL penalty of GaussianQuantRegularizer (gaussian.py:76-98) as one
// single-pass kernel (+ its closed-form backward).  HBM-bound: reads z (8 B per latent
// element) and eps (4 B), writes zhat (4 B) and one band weight per group.
//
// Thread <-> (token, sub-codebook j); it walks the g channels {k*ncb + j} of its group, so the
// group sum (reshape(b,l,g,ncb).sum(2), gaussian.py:87-88) never leaves registers.
//   BCHW: thread index = (bi*ncb + j)*l + t   -> every load/store is coalesced along t
//   BLC : thread index = (bi*l + t)*ncb + j   -> for fixed k the ncb threads of a token read
//         contiguous channels; with ncb == 1 each thread streams its 2x64 B with float4 loads
// Block partials {sum kl2, sum w*kl2, min, max} go through warp shuffles, then one double4
// per block; the last block to finish reduces them in block order (deterministic).
#include "gq_common.cuh"
#include "gq_internal.h"

namespace gq {

constexpr int kKlThreads = 256;
constexpr float kKlScale = 0.7213f;  // python folds 1.4426 * 0.5 before the tensor multiply

struct KlPartial { double s_all, s_w; float mn, mx; };

__device__ __forceinline__ void kl_addr(int layout, int64_t i, int64_t l, int ncb, int C, int64_t &zbase,
                                        int64_t &zstride, int64_t &obase, int64_t &ostride,
                                        int64_t &lv_off, int64_t &row)
{
    if (layout == GQ_LAYOUT_BCHW) {
        int64_t t = i % l;
        int j = (int)((i / l) % ncb);
        int64_t bi = i / (l * ncb);
        zbase = (bi * 2 * C + j) * l + t;
        zstride = (int64_t)ncb * l;
        lv_off = (int64_t)C * l;
        obase = (bi * C + j) * l + t;
        ostride = (int64_t)ncb * l;
        row = (bi * l + t) * ncb + j;
    } else {
        int j = (int)(i % ncb);
        int64_t tok = i / ncb;
        zbase = tok * 2 * C + j;
        zstride = ncb;
        lv_off = C;
        obase = tok * C + j;
        ostride = ncb;
        row = i;
    }
}

template <int G>
__global__ void __launch_bounds__(kKlThreads)
kl_fwd_kernel(const float *__restrict__ z, const float *__restrict__ eps, float *__restrict__ zhat,
              float *__restrict__ w_out, float *__restrict__ out_stats, int64_t b, int64_t l, int C,
              int layout, float lv_lo, float lv_hi, float hi_thr, float lo_thr, float lam, float lam_min,
              float lam_max, KlPartial *__restrict__ partials, unsigned int *__restrict__ done_counter)
{
    const int ncb = C / G;
    const int64_t total = b * l * ncb;
    float s_all = 0.f, s_w = 0.f, mn = CUDART_INF_F, mx = -CUDART_INF_F;

    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t zb, zs, ob, os, lvo, row;
        kl_addr(layout, i, l, ncb, C, zb, zs, ob, os, lvo, row);
        float mu[G], lv[G], ep[G];
        if (layout == GQ_LAYOUT_BLC && ncb == 1 && G % 4 == 0) {
            load_row<G>(z + zb, mu);
            load_row<G>(z + zb + lvo, lv);
            if (eps) load_row<G>(eps + ob, ep);
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) {
                mu[k] = __ldg(z + zb + k * zs);
                lv[k] = __ldg(z + zb + lvo + k * zs);
                if (eps) ep[k] = __ldg(eps + ob + k * os);
            }
        }
        float acc = 0.f;
        float zh[G];
#pragma unroll
        for (int k = 0; k < G; ++k) {
            float v = fminf(fmaxf(lv[k], lv_lo), lv_hi);
            float var = expf(v);
            float sd = expf(0.5f * v);
            float e = mu[k] * mu[k] + var;
            e = e - 1.0f;
            e = e - v;
            acc += kKlScale * e;
            zh[k] = eps ? fmaf(ep[k], sd, mu[k]) : mu[k];
        }
        if (zhat) {
            if (layout == GQ_LAYOUT_BLC && ncb == 1 && G % 4 == 0) {
#pragma unroll
                for (int k = 0; k < G / 4; ++k)
                    reinterpret_cast<float4 *>(zhat + ob)[k] =
                        make_float4(zh[4 * k], zh[4 * k + 1], zh[4 * k + 2], zh[4 * k + 3]);
            } else {
#pragma unroll
                for (int k = 0; k < G; ++k) zhat[ob + k * os] = zh[k];
            }
        }
        // ge/eq/le band weighting, gaussian.py:91-96
        float w = 0.f;
        if (acc > hi_thr) w += lam_max;
        if (acc <= hi_thr && acc >= lo_thr) w += 1.0f;
        if (acc < lo_thr) w += lam_min;
        if (w_out) w_out[row] = w;
        s_all += acc;
        s_w += w * acc;
        mn = fminf(mn, acc);
        mx = fmaxf(mx, acc);
    }

    // block reduction
    __shared__ double sh_all[kKlThreads / 32], sh_w[kKlThreads / 32];
    __shared__ float sh_mn[kKlThreads / 32], sh_mx[kKlThreads / 32];
    __shared__ bool is_last;
    double d_all = s_all, d_w = s_w;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        d_all += __shfl_xor_sync(0xffffffffu, d_all, off);
        d_w += __shfl_xor_sync(0xffffffffu, d_w, off);
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh_all[warp] = d_all; sh_w[warp] = d_w; sh_mn[warp] = mn; sh_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        KlPartial p{0.0, 0.0, CUDART_INF_F, -CUDART_INF_F};
        for (int w = 0; w < kKlThreads / 32; ++w) {
            p.s_all += sh_all[w];
            p.s_w += sh_w[w];
            p.mn = fminf(p.mn, sh_mn[w]);
            p.mx = fmaxf(p.mx, sh_mx[w]);
        }
        partials[blockIdx.x] = p;
        __threadfence();
        unsigned int prev = atomicAdd(done_counter, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        // the last block to finish folds the per-block partials: thread t takes partials t, t+256, ...
        // in order, then a fixed shuffle/shared tree -- the same association on every run
        __threadfence();
        double t_all = 0.0, t_w = 0.0;
        float t_mn = CUDART_INF_F, t_mx = -CUDART_INF_F;
        for (unsigned int k = threadIdx.x; k < gridDim.x; k += blockDim.x) {
            const volatile KlPartial *p = &partials[k];
            t_all += p->s_all;
            t_w += p->s_w;
            t_mn = fminf(t_mn, p->mn);
            t_mx = fmaxf(t_mx, p->mx);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            t_all += __shfl_xor_sync(0xffffffffu, t_all, off);
            t_w += __shfl_xor_sync(0xffffffffu, t_w, off);
            t_mn = fminf(t_mn, __shfl_xor_sync(0xffffffffu, t_mn, off));
            t_mx = fmaxf(t_mx, __shfl_xor_sync(0xffffffffu, t_mx, off));
        }
        __syncthreads();   // sh_* were read by thread 0 above
        if (lane == 0) { sh_all[warp] = t_all; sh_w[warp] = t_w; sh_mn[warp] = t_mn; sh_mx[warp] = t_mx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            KlPartial t{0.0, 0.0, CUDART_INF_F, -CUDART_INF_F};
            for (int w = 0; w < kKlThreads / 32; ++w) {
                t.s_all += sh_all[w];
                t.s_w += sh_w[w];
                t.mn = fminf(t.mn, sh_mn[w]);
                t.mx = fmaxf(t.mx, sh_mx[w]);
            }
            // kl_loss = lam * sum(w*kl2) / b  (gaussian.py:96-98)
            out_stats[0] = (float)((double)lam * t.s_w / (double)b);
            out_stats[1] = (float)(t.s_all / (double)total);
            out_stats[2] = t.mn;
            out_stats[3] = t.mx;
            out_stats[4] = (float)t.s_all;
            out_stats[5] = (float)t.s_w;
            out_stats[6] = 0.f;
            out_stats[7] = 0.f;
            *done_counter = 0u;  // re-arm for the next launch on this workspace
        }
    }
}

// Backward of {kl_loss, zhat} w.r.t. z.  With s = w*lam/b*grad_kl:
//   d/dmu     = s*2*0.7213*mu           + grad_zhat
//   d/dlogvar = s*0.7213*(exp(lv) - 1)  + grad_zhat*eps*0.5*exp(0.5*lv)    (0 outside the clamp)
template <int G>
__global__ void __launch_bounds__(kKlThreads)
kl_bwd_kernel(const float *__restrict__ z, const float *__restrict__ eps, const float *__restrict__ w,
              const float *__restrict__ grad_zhat, const float *__restrict__ grad_kl,
              float *__restrict__ grad_z, int64_t b, int64_t l, int C, int layout, float lv_lo, float lv_hi,
              float lam)
{
    const int ncb = C / G;
    const int64_t total = b * l * ncb;
    const float gk = grad_kl ? __ldg(grad_kl) : 0.f;
    const float sc = lam / (float)b * gk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        int64_t zb, zs, ob, os, lvo, row;
        kl_addr(layout, i, l, ncb, C, zb, zs, ob, os, lvo, row);
        const float s = sc * (w ? __ldg(w + row) : 1.0f);
        float mu[G], lr[G], ep[G], gz[G], gmu[G], glv[G];
        const bool vec = (layout == GQ_LAYOUT_BLC && ncb == 1 && G % 4 == 0);
        if (vec) {
            load_row<G>(z + zb, mu);
            load_row<G>(z + zb + lvo, lr);
            if (eps) load_row<G>(eps + ob, ep);
            if (grad_zhat) load_row<G>(grad_zhat + ob, gz);
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) {
                mu[k] = __ldg(z + zb + k * zs);
                lr[k] = __ldg(z + zb + lvo + k * zs);
                if (eps) ep[k] = __ldg(eps + ob + k * os);
                if (grad_zhat) gz[k] = __ldg(grad_zhat + ob + k * os);
            }
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float e = eps ? ep[k] : 0.f;
            const float g_ = grad_zhat ? gz[k] : 0.f;
            const float v = fminf(fmaxf(lr[k], lv_lo), lv_hi);
            const bool inside = (lr[k] >= lv_lo) && (lr[k] <= lv_hi);
            const float var = expf(v);
            const float sd = expf(0.5f * v);
            gmu[k] = s * (2.0f * kKlScale) * mu[k] + g_;
            glv[k] = inside ? (s * kKlScale * (var - 1.0f) + g_ * e * 0.5f * sd) : 0.f;
        }
        if (vec) {
#pragma unroll
            for (int k = 0; k < G / 4; ++k) {
                reinterpret_cast<float4 *>(grad_z + zb)[k] = make_float4(gmu[4 * k], gmu[4 * k + 1], gmu[4 * k + 2], gmu[4 * k + 3]);
                reinterpret_cast<float4 *>(grad_z + zb + lvo)[k] = make_float4(glv[4 * k], glv[4 * k + 1], glv[4 * k + 2], glv[4 * k + 3]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) {
                grad_z[zb + k * zs] = gmu[k];
                grad_z[zb + lvo + k * zs] = glv[k];
            }
        }
    }
}

static int kl_blocks(int64_t total)
{
    int64_t need = (total + kKlThreads - 1) / kKlThreads;
    return (int)imin(need, (int64_t)148 * 8);
}

int64_t kl_workspace_bytes(int64_t, int64_t, int)
{
    return (int64_t)sizeof(KlPartial) * 148 * 8 + 256;
}

int launch_kl_fwd(const float *z, const float *eps, float *zhat, float *w_out, float *out_stats,
                  int64_t b, int64_t l, int C, int g, int layout, float lv_lo, float lv_hi,
                  float target_bits, float tol, float lam, float lam_min, float lam_max, void *ws,
                  int64_t ws_bytes, cudaStream_t stream)
{
    if (ws_bytes < kl_workspace_bytes(b, l, C)) { set_error("kl workspace too small"); return GQ_ERR_WORKSPACE; }
    const int64_t total = b * l * (C / g);
    if (total <= 0) { set_error("empty latent"); return GQ_ERR_INVALID; }
    // workspace: [0,256) counter (zeroed once by the caller's allocation helper or here), then partials
    unsigned int *counter = reinterpret_cast<unsigned int *>(ws);
    KlPartial *partials = reinterpret_cast<KlPartial *>(reinterpret_cast<char *>(ws) + 256);
    cudaError_t e = cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    const int blocks = kl_blocks(total);
    const float hi = target_bits + tol, lo = target_bits - tol;
#define GQ_KLF(GG) kl_fwd_kernel<GG><<<blocks, kKlThreads, 0, stream>>>(z, eps, zhat, w_out, out_stats, b, l, C, \
        layout, lv_lo, lv_hi, hi, lo, lam, lam_min, lam_max, partials, counter)
    switch (g) {
    case 1: GQ_KLF(1); break;
    case 2: GQ_KLF(2); break;
    case 4: GQ_KLF(4); break;
    case 8: GQ_KLF(8); break;
    case 16: GQ_KLF(16); break;
    case 32: GQ_KLF(32); break;
    default: set_error("group %d unsupported", g); return GQ_ERR_INVALID;
    }
#undef GQ_KLF
    count_launch();
    return check_launch("kl_fwd_kernel");
}

int launch_kl_bwd(const float *z, const float *eps, const float *w, const float *grad_zhat,
                  const float *grad_kl, float *grad_z, int64_t b, int64_t l, int C, int g, int layout,
                  float lv_lo, float lv_hi, float lam, cudaStream_t stream)
{
    const int64_t total = b * l * (C / g);
    if (total <= 0) { set_error("empty latent"); return GQ_ERR_INVALID; }
    const int blocks = kl_blocks(total);
#define GQ_KLB(GG) kl_bwd_kernel<GG><<<blocks, kKlThreads, 0, stream>>>(z, eps, w, grad_zhat, grad_kl, grad_z, b, l, \
        C, layout, lv_lo, lv_hi, lam)
    switch (g) {
    case 1: GQ_KLB(1); break;
    case 2: GQ_KLB(2); break;
    case 4: GQ_KLB(4); break;
    case 8: GQ_KLB(8); break;
    case 16: GQ_KLB(16); break;
    case 32: GQ_KLB(32); break;
    default: set_error("group %d unsupported", g); return GQ_ERR_INVALID;
    }
#undef GQ_KLB
    count_launch();
    return check_launch("kl_bwd_kernel");
}

}  // namespace gq
