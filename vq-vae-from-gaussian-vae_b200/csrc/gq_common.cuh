// gq_common.cuh -- shared device helpers for libgq_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <math_constants.h>
#include "../../include/gq_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libgq_b200 is written for sm_100a (B200) only"
#endif

namespace gq {

// math.log(math.sqrt(2*math.pi)) after torch casts the python double to fp32
// (torch.distributions.Normal.log_prob via pit/quantization/gaussian.py:141-146).
__device__ __constant__ static const float kC0 = 0.9189385332046727f;
#define GQ_C0F 0.9189385332046727f

constexpr int kMaxG = 32;

__host__ __device__ inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t imax(int64_t a, int64_t b) { return a > b ? a : b; }

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);

// torch.sum over a contiguous innermost dim of G <= 32 fp32 values on the CPU: eight strided
// lanes, then lanes added in order (ATen vectorized_inner_sum, Vectorized<float>::size()==8).
// The oracle (oracle/gq_oracle.c:sum_torch_inner) is pinned bit-for-bit to the reference
// with this order; the device code must reproduce it exactly, hence the _rn intrinsics
// (never contracted into FMAs).
template <int G>
__device__ __forceinline__ float sum_torch_inner(const float (&t)[G])
{
    if constexpr (G < 8) {
        float acc = 0.0f;
#pragma unroll
        for (int k = 0; k < G; ++k) acc = __fadd_rn(acc, t[k]);
        return __fadd_rn(acc, 0.0f);
    } else {
        constexpr int NV = G / 8;
        float p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float a = t[j];
#pragma unroll
            for (int v = 1; v < NV; ++v) a = __fadd_rn(a, t[j + 8 * v]);
            p[j] = a;
        }
        float acc = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = __fadd_rn(acc, p[j]);
        return acc;
    }
}

// One (row, codeword) log-ratio in the reference torch op order, gaussian.py:141-147:
//   t = x - mu ; t = t*t ; t = -t ; t = t / (2*(std*std)) ; t = t - log(std) ; t = t - c0 ;
//   t = t - (nlp*beta) ; S = sum_k t
// den = 2*(std*std), lstd = log(std), nlpb = nlp*beta are the invariant sub-expressions.
template <int G>
__device__ __forceinline__ float score_exact(const float (&x)[G], const float (&nlpb)[G],
                                             const float (&mu)[G], const float (&den)[G],
                                             const float (&lstd)[G])
{
    float t[G];
#pragma unroll
    for (int k = 0; k < G; ++k) {
        float d = __fsub_rn(x[k], mu[k]);
        float d2 = __fmul_rn(d, d);
        float q = __fdiv_rn(-d2, den[k]);
        float a = __fsub_rn(q, lstd[k]);
        float b = __fsub_rn(a, GQ_C0F);
        t[k] = __fsub_rn(b, nlpb[k]);
    }
    return sum_torch_inner<G>(t);
}

// torch.clamp(x, lo, hi): NaN propagates (fminf / fmaxf would silently return the bound, turning a NaN from the
// encoder's logvar head into a finite loss, zhat and gradient; the reference yields NaN, which trainers detect)
__device__ __forceinline__ float clamp_torch(float x, float lo, float hi)
{
    return (x != x) ? x : fminf(fmaxf(x, lo), hi);
}

template <int G>
__device__ __forceinline__ void load_row(const float *__restrict__ p, float (&v)[G])
{
    if constexpr (G % 4 == 0) {
#pragma unroll
        for (int k = 0; k < G / 4; ++k) {
            float4 f = __ldg(reinterpret_cast<const float4 *>(p) + k);
            v[4 * k + 0] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < G; ++k) v[k] = __ldg(p + k);
    }
}

}  // namespace gq

// Operand-image geometry of the tensor-core search.  Groups g <= 4 use a split-precision image: with
// K = 2g <= 8 the plain fp16 image leaves most of a K = 16 slice empty, and a 4-dimensional codebook of
// 65 536 points is so dense that hundreds of codewords fall inside the fp16 error window of the best
// one.  Three fp16 products per term (hi*hi + hi*lo + lo*hi) bring the search to ~2^-17 relative
// accuracy for the cost of one extra K = 16 slice at g = 4 and none at g <= 2.
//
// g = 8 and g = 16 normally use the plain image (their windows are tight enough for typical latents), but
// rows with very small variances blow the fp16 window up and would fall back to the CUDA-core exact
// scan.  For those groups a second, split-precision image ("robust" mode, K = 6g = 48 / 96) is kept and
// a call switches to it when the previous calls on the codebook saw many such rows (gq_codebook_set_mode).
__host__ __device__ constexpr bool gq_split_image(int g) { return g <= 4; }                     // always split
__host__ __device__ constexpr bool gq_has_robust_image(int g) { return g == 8 || g == 16; }    // optional split
// fp16 accumulators ("ACC16"): the plain fp16 image of groups 8 and 16 is contracted into fp16 TMEM
// accumulators, read back packed (two scores per register) and max-reduced with packed 3-input fp16
// maxima -- half the ALU work, registers and tensor-memory read time of the fp32 epilogue, which is what
// bounded the search kernel.  The accumulator rounding (one RN per tcgen05.mma, measured) is covered by a
// term of the per-row window that is relative to the running maximum (kWinRel16, prep_row in gq_tc.cu).
// The codebook image of such a group stores its rows permuted inside every 64-codeword half tile so that
// the two halves of a packed register belong to two different 32-codeword chunks (gq_acc16_image_row).
// Split-precision images need the fp32 accumulator and keep it.  -DGQ_ACC16=0 restores the fp32 epilogue.
#ifndef GQ_ACC16
#define GQ_ACC16 1
#endif
__host__ __device__ constexpr bool gq_acc16(int g, bool split) { return GQ_ACC16 != 0 && !split && (g == 8 || g == 16); }
// Experimental (-DGQ_ACC16_TS=1): the fp16-accumulator search with its row operand in tensor memory (TS-form
// tcgen05.mma, gq_tc.cu quantize_tc16_kernel).  The TS instruction is faster (87 vs 114 cycles, scripts/ubench/
// mma_rate.cu) but leaves tensor memory for only three accumulator stages shared by the two row blocks; with the
// column-half epilogue the kernel measured 0.268 ms against 0.205 ms for the SS form (profiles/README.md), correct on
// the whole GPU suite.  Off by default.
#ifndef GQ_ACC16_TS
#define GQ_ACC16_TS 0
#endif
__host__ __device__ constexpr bool gq_acc16_ts(int g, bool split) { return GQ_ACC16_TS != 0 && gq_acc16(g, split); }
// image row (position in the MMA N dimension) of codeword n: half tile h = (n % 128) / 64, chunk parity
// p = (n % 64) / 32, lane i = n % 32  ->  column 64 h + 2 i + p, i.e. packed register i = (chunk 2h, chunk 2h+1)
__host__ __device__ constexpr int64_t gq_acc16_image_row(int64_t n)
{
    return (n & ~(int64_t)63) + 2 * (n & 31) + ((n >> 5) & 1);
}
__host__ __device__ constexpr int gq_image_k(int g, bool split) { return split ? 6 * g : 2 * g; }
__host__ __device__ constexpr int gq_image_kpad(int g, bool split) { return (gq_image_k(g, split) + 15) / 16 * 16; }
__host__ __device__ constexpr int gq_image_kpad(int g) { return gq_image_kpad(g, gq_split_image(g)); }

namespace gq {
extern bool g_chained_launch;   // gq_api.cu; gq_set_chained_launch()

// ---- programmatic dependent launch ----------------------------------------------------------------
// The step is a chain of dependent kernels on one stream.  Each is launched with the
// programmatic-stream-serialization attribute and starts with griddepcontrol.wait, so the GPU can set
// the next grid up while the previous one drains instead of paying a launch gap per boundary; nothing
// produced by the predecessor is touched before the wait.  griddepcontrol.launch_dependents is issued
// where letting the successor's CTAs become resident cannot starve the kernel's own later waves.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                  Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_chained_launch ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// One row's result into the module's output tensors (FinalLayout, gq_internal.h): index element and the G channel
// values of the codeword.  r = (bi*l + t)*ncb + j.
template <int G>
__device__ __forceinline__ void final_store_row(int64_t *__restrict__ idx_final, float *__restrict__ zhat_final, int64_t l, int ncb,
                                                int layout, unsigned long long *__restrict__ usage_hist,
                                                const float *__restrict__ cb, int64_t bi, int64_t t, int j, int64_t id)
{
    const int C = ncb * G;
    if (layout == GQ_LAYOUT_BCHW) idx_final[(bi * ncb + j) * l + t] = id;
    else idx_final[(bi * l + t) * ncb + j] = id;
    if (usage_hist) atomicAdd(usage_hist + id, 1ull);
    float v[G];
    load_row<G>(cb + id * G, v);
    if (layout == GQ_LAYOUT_BCHW) {
        float *o = zhat_final + (bi * C + j) * l + t;
#pragma unroll
        for (int k = 0; k < G; ++k) o[(int64_t)k * ncb * l] = v[k];
    } else {
        float *o = zhat_final + (bi * l + t) * C + j;
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

}  // namespace gq

// opaque handle behind gq_codebook_t
struct gq_codebook {
    int64_t N;        // codewords
    int g;            // group (dims per codeword)
    int kpad;         // K of the fp16 operand image (gq_image_kpad(g))
    int64_t n_pad;    // N rounded up to the 128-codeword search tile
    int64_t n_pad2;   // rows of the operand images: N rounded up to 256 (the fp16-accumulator kernel contracts 256-codeword tiles), zero beyond N
    float beta;
    void *base;       // the single device allocation
    float *cb;        // [N,g] fp32 codebook
    float *nlpb;      // [N,g] fp32 beta*normal_log_prob (one rounding, as torch)
    __half *bimg;     // fp16 operand image, UMMA K-major no-swizzle core-matrix order
    float *meta;      // device floats: [0]=max ||x||^2, [1]=max ||x||, [2..2+2g) per-column max |B|, [2+2*kMaxG] their sum
    int device;
    // robust mode (g = 8, 16): split-precision image and the feedback that drives GQ_MODE_AUTO
    int kpad_robust;          // 0 when the group has no robust image
    __half *bimg_robust;
    int mode;                 // gq_mode
    uint32_t *wide_dev;       // device word: rows of the calls in flight whose PLAIN window exceeded the cap
    volatile uint32_t *fb_host;   // pinned, mapped: [0] = such rows in the last completed call, [1] = its row count
    uint32_t *fb_dev;         // device view of fb_host
    unsigned long long *usage_hist;   // optional device uint64[N]: gq_quantize_latent accumulates its indices into it
    // per-dimension sorted view of the codebook (pruned exact scan of the rows the search cannot certify, gq_tc.cu):
    // sort_val[k*N + p] = p-th smallest x_k, sort_idx[k*N + p] = the codeword it belongs to
    float *sort_val;
    uint32_t *sort_idx;
};
