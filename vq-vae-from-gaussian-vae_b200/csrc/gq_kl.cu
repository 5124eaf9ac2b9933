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
#include "gq_ptx.cuh"

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

// Block partials {sum kl2, sum w*kl2, min, max} -> one KlPartial per block; the last block to finish
// folds all partials in a fixed order (deterministic) and writes the eight output statistics.
__device__ __forceinline__ void kl_finish(float s_all, float s_w, float mn, float mx, float *__restrict__ out_stats,
                                          int64_t b, int64_t total, float lam, KlPartial *__restrict__ partials,
                                          unsigned int *__restrict__ done_counter)
{
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
            float v = clamp_torch(lv[k], lv_lo, lv_hi);
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

    kl_finish(s_all, s_w, mn, mx, out_stats, b, total, lam, partials, done_counter);
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
            const float v = clamp_torch(lr[k], lv_lo, lv_hi);
            const bool inside = (lr[k] >= lv_lo) && (lr[k] <= lv_hi);
            const float var = expf(v);
            const float sd = expf(0.5f * v);
            gmu[k] = s * (2.0f * kKlScale) * mu[k] + g_;
            glv[k] = inside ? (s * kKlScale * (var - 1.0f) + g_ * e * 0.5f * sd) : (lr[k] != lr[k] ? lr[k] : 0.f);   // clamp backward: grad * mask, and NaN * 0 = NaN as in torch
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


// ---- coalesced fast paths -------------------------------------------------------------------------
// BLC with C = 16 (every shipped config): thread <-> (token, channel quad).  A token's latent row is
// [mu(16) | logvar(16)] = 128 B; the four lanes of a token read mu / logvar / eps as three 16-byte
// loads each, so a warp instruction covers eight 64-byte runs (sector-perfect), and the per-group sum
// crosses the four lanes with two shuffles.  Channel c belongs to group c % ncb (gaussian.py:87), so
// with ncb in {1,2,4} quad slot i always maps to group i % ncb.
__device__ __forceinline__ float kl_band_weight(float acc, float hi_thr, float lo_thr, float lam_min, float lam_max)
{
    float w = 0.f;
    if (acc > hi_thr) w += lam_max;
    if (acc <= hi_thr && acc >= lo_thr) w += 1.0f;
    if (acc < lo_thr) w += lam_min;
    return w;
}

template <int NCB>
__global__ void __launch_bounds__(kKlThreads)
kl_fwd_blc16_kernel(const float *__restrict__ z, const float *__restrict__ eps, float *__restrict__ zhat,
                    float *__restrict__ w_out, float *__restrict__ out_stats, int64_t b, int64_t l, float lv_lo,
                    float lv_hi, float hi_thr, float lo_thr, float lam, float lam_min, float lam_max,
                    KlPartial *__restrict__ partials, unsigned int *__restrict__ done_counter)
{
    const int64_t tokens = b * l;
    const int64_t total = tokens * NCB;
    float s_all = 0.f, s_w = 0.f, mn = CUDART_INF_F, mx = -CUDART_INF_F;
    // whole warps iterate together (the shuffles below need all four lanes of a token)
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < tokens * 4; i0 += nthreads) {
        const int64_t i = i0 + threadIdx.x;
        const bool live = i < tokens * 4;
        const int64_t tok = live ? i >> 2 : 0;
        const int q = (int)(i & 3);
        const float4 m4 = live ? __ldg(reinterpret_cast<const float4 *>(z + tok * 32) + q) : make_float4(0, 0, 0, 0);
        const float4 l4 = live ? __ldg(reinterpret_cast<const float4 *>(z + tok * 32 + 16) + q) : make_float4(0, 0, 0, 0);
        float4 e4 = make_float4(0, 0, 0, 0);
        if (eps && live) e4 = __ldg(reinterpret_cast<const float4 *>(eps + tok * 16) + q);
        const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, lr[4] = {l4.x, l4.y, l4.z, l4.w}, ep[4] = {e4.x, e4.y, e4.z, e4.w};
        float sl[4], zh[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float v = clamp_torch(lr[k], lv_lo, lv_hi);
            const float var = expf(v);
            const float sd = expf(0.5f * v);
            float e = mu[k] * mu[k] + var;
            e = e - 1.0f;
            e = e - v;
            sl[k] = kKlScale * e;
            zh[k] = eps ? fmaf(ep[k], sd, mu[k]) : mu[k];
        }
        if (zhat && live) reinterpret_cast<float4 *>(zhat + tok * 16)[q] = make_float4(zh[0], zh[1], zh[2], zh[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sl[k] += __shfl_xor_sync(0xffffffffu, sl[k], 1);
            sl[k] += __shfl_xor_sync(0xffffffffu, sl[k], 2);
        }
        if (live && q == 0) {
            float grp[NCB];
            if (NCB == 1) grp[0] = (sl[0] + sl[1]) + (sl[2] + sl[3]);
            else if (NCB == 2) { grp[0] = sl[0] + sl[2]; grp[NCB - 1] = sl[1] + sl[3]; }
            else {
#pragma unroll
                for (int j = 0; j < NCB; ++j) grp[j] = sl[j & 3];
            }
#pragma unroll
            for (int j = 0; j < NCB; ++j) {
                const float acc = grp[j];
                const float w = kl_band_weight(acc, hi_thr, lo_thr, lam_min, lam_max);
                if (w_out) w_out[tok * NCB + j] = w;
                s_all += acc;
                s_w += w * acc;
                mn = fminf(mn, acc);
                mx = fmaxf(mx, acc);
            }
        }
    }
    kl_finish(s_all, s_w, mn, mx, out_stats, b, total, lam, partials, done_counter);
}

template <int NCB>
__global__ void __launch_bounds__(kKlThreads)
kl_bwd_blc16_kernel(const float *__restrict__ z, const float *__restrict__ eps, const float *__restrict__ w,
                    const float *__restrict__ grad_zhat, const float *__restrict__ grad_kl,
                    float *__restrict__ grad_z, int64_t b, int64_t l, float lv_lo, float lv_hi, float lam)
{
    const int64_t tokens = b * l;
    const float gk = grad_kl ? __ldg(grad_kl) : 0.f;
    const float sc = lam / (float)b * gk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tokens * 4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tok = i >> 2;
        const int q = (int)(i & 3);
        const float4 m4 = __ldg(reinterpret_cast<const float4 *>(z + tok * 32) + q);
        const float4 l4 = __ldg(reinterpret_cast<const float4 *>(z + tok * 32 + 16) + q);
        float4 e4 = make_float4(0, 0, 0, 0), g4 = make_float4(0, 0, 0, 0);
        if (eps) e4 = __ldg(reinterpret_cast<const float4 *>(eps + tok * 16) + q);
        if (grad_zhat) g4 = __ldg(reinterpret_cast<const float4 *>(grad_zhat + tok * 16) + q);
        float wg[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) wg[k] = w ? __ldg(w + tok * NCB + (k % NCB)) : 1.0f;
        const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, lr[4] = {l4.x, l4.y, l4.z, l4.w};
        const float ep[4] = {e4.x, e4.y, e4.z, e4.w}, gz[4] = {g4.x, g4.y, g4.z, g4.w};
        float gmu[4], glv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float s = sc * wg[k];
            const float v = clamp_torch(lr[k], lv_lo, lv_hi);
            const bool inside = (lr[k] >= lv_lo) && (lr[k] <= lv_hi);
            const float var = expf(v);
            const float sd = expf(0.5f * v);
            gmu[k] = s * (2.0f * kKlScale) * mu[k] + gz[k];
            glv[k] = inside ? (s * kKlScale * (var - 1.0f) + gz[k] * ep[k] * 0.5f * sd) : (lr[k] != lr[k] ? lr[k] : 0.f);   // clamp backward: grad * mask, and NaN * 0 = NaN as in torch
        }
        reinterpret_cast<float4 *>(grad_z + tok * 32)[q] = make_float4(gmu[0], gmu[1], gmu[2], gmu[3]);
        reinterpret_cast<float4 *>(grad_z + tok * 32 + 16)[q] = make_float4(glv[0], glv[1], glv[2], glv[3]);
    }
}

// BCHW with l % 4 == 0 and C = 16: lane <-> (channel quarter, token quad).  A warp covers 8 consecutive
// token quads (128 bytes of every channel plane) x 4 channel quarters; each lane walks its 4 planes with
// 16-byte loads, and the per-group sums cross the channel quarters with two shuffles (lane bits 3, 4).
template <int NCB>
__global__ void __launch_bounds__(kKlThreads)
kl_fwd_bchw16_kernel(const float *__restrict__ z, const float *__restrict__ eps, float *__restrict__ zhat,
                     float *__restrict__ w_out, float *__restrict__ out_stats, int64_t b, int64_t l,
                     float lv_lo, float lv_hi, float hi_thr, float lo_thr, float lam, float lam_min, float lam_max,
                     KlPartial *__restrict__ partials, unsigned int *__restrict__ done_counter)
{
    constexpr int C = 16;
    const int64_t l4 = l / 4;
    const int64_t quads = b * l4;
    const int64_t total = b * l * NCB;
    const int lane = threadIdx.x & 31;
    const int part = lane >> 3;
    float s_all = 0.f, s_w = 0.f, mn = CUDART_INF_F, mx = -CUDART_INF_F;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t wq = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 8; wq < quads; wq += warps * 8) {
        const int64_t qi = wq + (lane & 7);
        const bool live = qi < quads;
        const int64_t bi = live ? qi / l4 : 0, t4 = live ? qi % l4 : 0;
        const float4 *zp = reinterpret_cast<const float4 *>(z + bi * 2 * C * l) + t4;
        const float4 *ep = eps ? reinterpret_cast<const float4 *>(eps + bi * C * l) + t4 : nullptr;
        float4 *op = zhat ? reinterpret_cast<float4 *>(zhat + bi * C * l) + t4 : nullptr;
        float acc[NCB][4];
#pragma unroll
        for (int j = 0; j < NCB; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = part * 4 + k;
            float4 m4 = make_float4(0, 0, 0, 0), v4 = m4, e4 = m4;
            if (live) {
                m4 = __ldg(zp + (int64_t)c * l4);
                v4 = __ldg(zp + (int64_t)(C + c) * l4);
                if (ep) e4 = __ldg(ep + (int64_t)c * l4);
            }
            const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, lr[4] = {v4.x, v4.y, v4.z, v4.w}, en[4] = {e4.x, e4.y, e4.z, e4.w};
            float zh[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float v = clamp_torch(lr[t], lv_lo, lv_hi);
                const float var = expf(v);
                const float sd = expf(0.5f * v);
                float e = mu[t] * mu[t] + var;
                e = e - 1.0f;
                e = e - v;
                acc[k % NCB][t] += kKlScale * e;      // channel part*4+k belongs to group k % NCB
                zh[t] = ep ? fmaf(en[t], sd, mu[t]) : mu[t];
            }
            if (op && live) op[(int64_t)c * l4] = make_float4(zh[0], zh[1], zh[2], zh[3]);
        }
#pragma unroll
        for (int j = 0; j < NCB; ++j)
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                acc[j][t] += __shfl_xor_sync(0xffffffffu, acc[j][t], 8);
                acc[j][t] += __shfl_xor_sync(0xffffffffu, acc[j][t], 16);
            }
        if (live && part == 0) {
            float wv[4][NCB];
#pragma unroll
            for (int t = 0; t < 4; ++t)
#pragma unroll
                for (int j = 0; j < NCB; ++j) {
                    const float a = acc[j][t];
                    const float w = kl_band_weight(a, hi_thr, lo_thr, lam_min, lam_max);
                    wv[t][j] = w;
                    s_all += a;
                    s_w += w * a;
                    mn = fminf(mn, a);
                    mx = fmaxf(mx, a);
                }
            if (w_out) {
                // rows (bi*l + 4*t4 + t)*NCB + j for t, j: 4*NCB consecutive floats
                float4 *wp = reinterpret_cast<float4 *>(w_out + (bi * l + t4 * 4) * NCB);
                const float *f = &wv[0][0];
#pragma unroll
                for (int q = 0; q < NCB; ++q) wp[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            }
        }
    }
    kl_finish(s_all, s_w, mn, mx, out_stats, b, total, lam, partials, done_counter);
}

// Backward, BCHW with l % 4 == 0 (any C, any group): thread <-> (image, channel, token quad).
__global__ void __launch_bounds__(kKlThreads)
kl_bwd_bchw4_kernel(const float *__restrict__ z, const float *__restrict__ eps, const float *__restrict__ w,
                    const float *__restrict__ grad_zhat, const float *__restrict__ grad_kl,
                    float *__restrict__ grad_z, int64_t b, int64_t l, int C, int ncb, float lv_lo, float lv_hi, float lam)
{
    const int64_t l4 = l / 4;
    const float gk = grad_kl ? __ldg(grad_kl) : 0.f;
    const float sc = lam / (float)b * gk;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < b * C * l4; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t4 = i % l4;
        const int c = (int)((i / l4) % C);
        const int64_t bi = i / (l4 * C);
        const int j = c % ncb;
        const float4 m4 = __ldg(reinterpret_cast<const float4 *>(z + (bi * 2 * C + c) * l) + t4);
        const float4 v4 = __ldg(reinterpret_cast<const float4 *>(z + (bi * 2 * C + C + c) * l) + t4);
        float4 e4 = make_float4(0, 0, 0, 0), g4 = make_float4(0, 0, 0, 0);
        if (eps) e4 = __ldg(reinterpret_cast<const float4 *>(eps + (bi * C + c) * l) + t4);
        if (grad_zhat) g4 = __ldg(reinterpret_cast<const float4 *>(grad_zhat + (bi * C + c) * l) + t4);
        float wg[4] = {1.f, 1.f, 1.f, 1.f};
        if (w) {
            if (ncb == 1) {
                const float4 w4 = __ldg(reinterpret_cast<const float4 *>(w + bi * l) + t4);
                wg[0] = w4.x; wg[1] = w4.y; wg[2] = w4.z; wg[3] = w4.w;
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) wg[t] = __ldg(w + (bi * l + t4 * 4 + t) * ncb + j);
            }
        }
        const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, lr[4] = {v4.x, v4.y, v4.z, v4.w};
        const float en[4] = {e4.x, e4.y, e4.z, e4.w}, gz[4] = {g4.x, g4.y, g4.z, g4.w};
        float gmu[4], glv[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float s = sc * wg[t];
            const float v = clamp_torch(lr[t], lv_lo, lv_hi);
            const bool inside = (lr[t] >= lv_lo) && (lr[t] <= lv_hi);
            const float var = expf(v);
            const float sd = expf(0.5f * v);
            gmu[t] = s * (2.0f * kKlScale) * mu[t] + gz[t];
            glv[t] = inside ? (s * kKlScale * (var - 1.0f) + gz[t] * en[t] * 0.5f * sd) : (lr[t] != lr[t] ? lr[t] : 0.f);   // clamp backward: grad * mask, and NaN * 0 = NaN as in torch
        }
        reinterpret_cast<float4 *>(grad_z + (bi * 2 * C + c) * l)[t4] = make_float4(gmu[0], gmu[1], gmu[2], gmu[3]);
        reinterpret_cast<float4 *>(grad_z + (bi * 2 * C + C + c) * l)[t4] = make_float4(glv[0], glv[1], glv[2], glv[3]);
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
    const float hi = target_bits + tol, lo = target_bits - tol;
    const int ncb = C / g;
    // coalesced fast paths (every shipped config); anything else takes the generic kernel
    if (layout == GQ_LAYOUT_BLC && C == 16 && (ncb == 1 || ncb == 2 || ncb == 4)) {
        const int fb = kl_blocks(b * l * 4);
#define GQ_KLF16(NN) kl_fwd_blc16_kernel<NN><<<fb, kKlThreads, 0, stream>>>(z, eps, zhat, w_out, out_stats, b, l, lv_lo, \
        lv_hi, hi, lo, lam, lam_min, lam_max, partials, counter)
        if (ncb == 1) GQ_KLF16(1); else if (ncb == 2) GQ_KLF16(2); else GQ_KLF16(4);
#undef GQ_KLF16
        count_launch();
        return check_launch("kl_fwd_blc16_kernel");
    }
    if (layout == GQ_LAYOUT_BCHW && l % 4 == 0 && C == 16 && (ncb == 1 || ncb == 2 || ncb == 4)) {
        const int fb = kl_blocks(b * l);      // four lanes (channel quarters) per token quad
#define GQ_KLFB(NN) kl_fwd_bchw16_kernel<NN><<<fb, kKlThreads, 0, stream>>>(z, eps, zhat, w_out, out_stats, b, l, lv_lo, \
        lv_hi, hi, lo, lam, lam_min, lam_max, partials, counter)
        if (ncb == 1) GQ_KLFB(1); else if (ncb == 2) GQ_KLFB(2); else GQ_KLFB(4);
#undef GQ_KLFB
        count_launch();
        return check_launch("kl_fwd_bchw16_kernel");
    }
    const int blocks = kl_blocks(total);
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
    const int ncb = C / g;
    if (layout == GQ_LAYOUT_BLC && C == 16 && (ncb == 1 || ncb == 2 || ncb == 4)) {
        const int fb = kl_blocks(b * l * 4);
#define GQ_KLB16(NN) kl_bwd_blc16_kernel<NN><<<fb, kKlThreads, 0, stream>>>(z, eps, w, grad_zhat, grad_kl, grad_z, b, l, \
        lv_lo, lv_hi, lam)
        if (ncb == 1) GQ_KLB16(1); else if (ncb == 2) GQ_KLB16(2); else GQ_KLB16(4);
#undef GQ_KLB16
        count_launch();
        return check_launch("kl_bwd_blc16_kernel");
    }
    if (layout == GQ_LAYOUT_BCHW && l % 4 == 0) {
        const int fb = kl_blocks(b * C * (l / 4));
        kl_bwd_bchw4_kernel<<<fb, kKlThreads, 0, stream>>>(z, eps, w, grad_zhat, grad_kl, grad_z, b, l, C, ncb, lv_lo, lv_hi, lam);
        count_launch();
        return check_launch("kl_bwd_bchw4_kernel");
    }
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

// Fold per-rank statistics (world x [kl_loss, bits_mean, bits_min, bits_max, sum_kl2, sum_w_kl2, -, -]) into the
// global batch's four scalars; sums in double, in rank order.  fminf / fmaxf skip NaN, so a NaN minimum or maximum of
// one shard is carried over explicitly (torch.min / max propagate it).
__global__ void kl_fold_kernel(const float *__restrict__ all_stats, int world, double lam, double global_batch,
                               double total_groups, float *__restrict__ out4)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s_kl2 = 0.0, s_wkl2 = 0.0;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    bool nan_mn = false, nan_mx = false;
    for (int r = 0; r < world; ++r) {
        const float *st = all_stats + 8 * r;
        s_kl2 += (double)st[4];
        s_wkl2 += (double)st[5];
        nan_mn |= st[2] != st[2];
        nan_mx |= st[3] != st[3];
        mn = fminf(mn, st[2]);
        mx = fmaxf(mx, st[3]);
    }
    out4[0] = (float)(lam * s_wkl2 / global_batch);
    out4[1] = (float)(s_kl2 / total_groups);
    out4[2] = nan_mn ? CUDART_NAN_F : mn;
    out4[3] = nan_mx ? CUDART_NAN_F : mx;
}

// The same fold with the exchange inside: ONE warp stores this rank's eight statistics into slot [rank] of every
// rank's peer-mapped buffer (NVLink peer stores, 32 bytes each), signals / awaits every rank (peer_signal_and_wait)
// and folds the world-size rows it now holds -- global_kl_stats as a single kernel, no collective library.
__global__ void __launch_bounds__(64)
peer_kl_fold_kernel(const float *__restrict__ stats, void *const *__restrict__ peer_bufs, int world, int rank, int64_t data_off,
                    int64_t flag_off, uint32_t seq, double lam, double global_batch, double total_groups,
                    float *__restrict__ out4, int32_t *__restrict__ err_flag)
{
    const int p = threadIdx.x;
    if (p < world) {
        const uint4 a = reinterpret_cast<const uint4 *>(stats)[0], b = reinterpret_cast<const uint4 *>(stats)[1];
        uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(peer_bufs[p]) + data_off + 32 * (int64_t)rank);
        dst[0] = a;
        dst[1] = b;
        __threadfence_system();
        if (!ptx::peer_signal_and_wait(peer_bufs, flag_off, world, rank, p, seq) && err_flag) *err_flag = 2;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const volatile float *all_stats = reinterpret_cast<const volatile float *>(reinterpret_cast<const char *>(peer_bufs[rank]) + data_off);
    double s_kl2 = 0.0, s_wkl2 = 0.0;
    float mn = CUDART_INF_F, mx = -CUDART_INF_F;
    bool nan_mn = false, nan_mx = false;
    for (int r = 0; r < world; ++r) {
        const volatile float *st = all_stats + 8 * r;
        const float v2 = st[2], v3 = st[3];
        s_kl2 += (double)st[4];
        s_wkl2 += (double)st[5];
        nan_mn |= v2 != v2;
        nan_mx |= v3 != v3;
        mn = fminf(mn, v2);
        mx = fmaxf(mx, v3);
    }
    out4[0] = (float)(lam * s_wkl2 / global_batch);
    out4[1] = (float)(s_kl2 / total_groups);
    out4[2] = nan_mn ? CUDART_NAN_F : mn;
    out4[3] = nan_mx ? CUDART_NAN_F : mx;
}

int launch_peer_kl_fold(const float *stats, void *const *peer_bufs, int world, int rank, int64_t data_off, int64_t flag_off,
                        uint32_t seq, double lam, double global_batch, double total_groups, float *out4, int32_t *err_flag,
                        cudaStream_t stream)
{
    peer_kl_fold_kernel<<<1, 64, 0, stream>>>(stats, peer_bufs, world, rank, data_off, flag_off, seq, lam, global_batch,
                                              total_groups, out4, err_flag);
    count_launch();
    return check_launch("peer_kl_fold_kernel");
}

int launch_kl_fold(const float *all_stats, int world, double lam, double global_batch, double total_groups, float *out4,
                   cudaStream_t stream)
{
    kl_fold_kernel<<<1, 32, 0, stream>>>(all_stats, world, lam, global_batch, total_groups, out4);
    count_launch();
    return check_launch("kl_fold_kernel");
}

}  // namespace gq
