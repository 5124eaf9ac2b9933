// gq_tc.cu -- the fused GaussianQuant quantizer for B200 (sm_100a).
//
// What it replaces: the eval branch of GaussianQuantRegularizer.forward, both backends
// (pit/quantization/gaussian.py:124-150): score every (row, codeword) pair, argmax per row,
// gather the winning codeword.  The reference materialises an [M, N] score matrix in HBM;
// here it lives only in tensor memory.
//
// Pipeline (three launches + one optional fallback launch, all on the caller's stream):
//   1. rows_prep_kernel   per row: A[r,:] = s_r*[beta/2 - 1/(2v) | mu/v] as fp16 in UMMA
//                         core-matrix order, and a RIGOROUS window W_r bounding twice the
//                         worst-case |fp16 tensor-core score - reference fp32 score| gap.
//   2. quantize_tc_kernel persistent, warp-specialised: bulk-async-copy (TMA engine) streams
//                         the prepared [x^2 | x] codebook image through a shared-memory ring,
//                         one thread issues tcgen05.mma (fp16 x fp16 -> fp32, M=128, N=128,
//                         K=2g) into four TMEM accumulators, eight epilogue warps pull tiles
//                         with tcgen05.ld (the whole 32x128 tile of a warp into registers, so
//                         the TMEM stage is released before any math) and keep a running row
//                         maximum with FMNMX3; a 32-column chunk whose maximum comes within
//                         W_r of the running maximum is recorded as (max, chunk id).
//   3. rescore_kernel     one warp per row: drop chunk records below (final max - W_r), score
//                         the 32 codewords of every surviving chunk (lane = codeword) in the
//                         reference's exact fp32 op order, first-max tie-break, write idx
//                         (int64) and zhat.  Rows whose window is too wide, that overflow
//                         their list, or that contain non-finite values go to
//   4. quantize_exact_kernel (gq_kernels.cu) through a device-side row list.
//
// Why the result is exactly the reference argmax: with E_r >= |G~[n]/s - (S_ref[n] - Q_r)| for
// all n (Q_r constant in n), the reference argmax n* satisfies
// G~[n*] >= G~[m] - 2 s E_r for every m, in particular for the running maxima the epilogue
// compares against (each is some G~[m]); so the chunk holding n* is always recorded (its maximum
// is >= G~[n*]), survives the final filter, and n* wins the exact re-scoring of that chunk.
// DESIGN.md ("error window") derives E_r.
#include "gq_common.cuh"
#include "gq_internal.h"
#include "gq_ptx.cuh"

#include <utility>
#include <vector>

namespace gq {

// ---- optional per-launch timing of the search kernel (bench.py's roofline leg) ----------------
static bool g_prof = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
void prof_enable(int on) { g_prof = on != 0; }
int prof_read(double *total_ms, int *count)
{
    double tot = 0.0;
    int n = 0;
    for (auto &pr : g_prof_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
            tot += ms;
            ++n;
        }
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    g_prof_events.clear();
    if (total_ms) *total_ms = tot;
    if (count) *count = n;
    return GQ_OK;
}

using namespace gq::ptx;

// ---- geometry ------------------------------------------------------------------------------
constexpr int kTileN = 128;          // codewords per B tile / MMA N
constexpr int kBlockRows = 128;      // rows per MMA (M)
constexpr int kUnitRows = 256;       // rows per CTA work unit (two row blocks)
constexpr int kStages = 6;           // B ring depth
constexpr int kAccStages = 2;        // TMEM accumulator stages per row block
constexpr int kTmemCols = 512;       // 2 row blocks x 2 stages x 128 fp32 columns
constexpr int kThreads = 384;        // 4 control warps + 8 epilogue warps
constexpr int kEpiWarp0 = 4;
constexpr int kCap = 32;             // chunk records per (row, segment)
constexpr int kChunk = 32;           // columns per recorded chunk
constexpr int kMinStepsPerCta = 8;

struct __align__(8) SegHdr { float runmax; uint32_t cnt; };
struct __align__(8) Cand { float score; uint32_t idx; };  // idx = chunk id (codeword / 32)

struct TcPlan {
    int64_t M, M_pad, N;
    int units;           // ceil(M / 256)
    int tiles;           // B tiles per unit = n_pad / 128
    int64_t total_steps; // units * tiles
    int ctas;
    int steps_per_cta;
    int nseg;            // max CTAs touching one unit
};

static TcPlan make_plan_sms(int64_t N, int64_t n_pad, int64_t M, int sms)
{
    TcPlan p;
    p.M = M;
    p.units = (int)((M + kUnitRows - 1) / kUnitRows);
    p.M_pad = (int64_t)p.units * kUnitRows;
    p.N = N;
    p.tiles = (int)(n_pad / kTileN);
    p.total_steps = (int64_t)p.units * p.tiles;
    int64_t max_ctas = (p.total_steps + kMinStepsPerCta - 1) / kMinStepsPerCta;
    p.ctas = (int)imin(sms, imax(1, max_ctas));
    p.steps_per_cta = (int)((p.total_steps + p.ctas - 1) / p.ctas);
    p.ctas = (int)((p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta);
    p.nseg = (p.tiles + p.steps_per_cta - 1) / p.steps_per_cta + 1;
    return p;
}

static TcPlan make_plan(const gq_codebook *cb, int64_t M)
{
    return make_plan_sms(cb->N, cb->n_pad, M, device_sm_count());
}

struct TcWorkspace {
    __half *a_img;      // [M_pad, kpad] fp16, core-matrix order
    float *wrow;        // [M_pad] scaled window (negative => row goes to the exact fallback)
    SegHdr *seg;        // [M_pad, nseg]
    Cand *cand;         // [M_pad, nseg, kCap]
    uint32_t *fb_rows;  // [M] fallback row list
    uint32_t *counters; // [0] fallback count
    int64_t bytes;
};

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static TcWorkspace carve(const gq_codebook *cb, const TcPlan &p, void *base)
{
    TcWorkspace w;
    char *ptr = reinterpret_cast<char *>(base);
    int64_t off = 0;
    w.a_img = reinterpret_cast<__half *>(ptr + off);
    off = align_up(off + p.M_pad * cb->kpad * 2, 256);
    w.wrow = reinterpret_cast<float *>(ptr + off);
    off = align_up(off + p.M_pad * 4, 256);
    w.seg = reinterpret_cast<SegHdr *>(ptr + off);
    off = align_up(off + p.M_pad * p.nseg * (int64_t)sizeof(SegHdr), 256);
    w.cand = reinterpret_cast<Cand *>(ptr + off);
    off = align_up(off + p.M_pad * p.nseg * (int64_t)kCap * (int64_t)sizeof(Cand), 256);
    w.fb_rows = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + p.M * 4, 256);
    w.counters = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + 256, 256);
    w.bytes = off;
    return w;
}

int64_t tc_workspace_bytes(const gq_codebook *cb, int64_t M)
{
    if (M <= 0) return 256;
    TcPlan p = make_plan(cb, M);
    return carve(cb, p, nullptr).bytes;
}

// ---- 1. row preparation -----------------------------------------------------------------------
// meta (from codebook prep): [0] = max_n ||x_n||^2, [1] = max_n ||x_n||, [2+j] = max_n |B[n,j]|.
// Window (unscaled score units), all terms upper bounds:
//   P     = min( sum_j |A_j| Bmax_j ,  max_k|a_k| * X2 + ||b||_2 * X1 )     >= sum_j |A_j B_nj|
//   E_tc  = 1.01 * 2^-10 * P                    fp16 rounding of A and of B (2^-11 each), the
//                                               product of both, fp32 tensor-core accumulation
//         + 2^-25 * (sum_j Bmax_j / s + sum_j |A_j|)      fp16 subnormal floors
//   E_ref = 2^-20 * sum_k[ (Xmax_k+|mu_k|)^2/(2 v_k) + |log std_k| + 2 + beta*X2max_k ]
//                                               rounding of the reference's own fp32 evaluation
//   W     = 2 * (E_tc + E_ref)
// A is built in double so that its own error (2^-52) is irrelevant next to the fp16 rounding.
template <int G, int KPAD>
__global__ void __launch_bounds__(128)
rows_prep_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
                 const float *__restrict__ meta, __half *__restrict__ a_img, float *__restrict__ wrow,
                 int64_t M, int64_t M_pad, float beta, float w_max)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M_pad) return;
    constexpr int KCH = KPAD / 8;
    char *rowbase = reinterpret_cast<char *>(a_img) + (r / 8) * (int64_t)KCH * 128 + (r % 8) * 16;
    double A[2 * G];
    bool bad = (r >= M);
    double amax = 0.0, bnorm2 = 0.0, e_ref = 0.0, p_naive = 0.0, maxA = 0.0;
    if (!bad) {
        float m[G], s[G];
        load_row<G>(mu + r * G, m);
        load_row<G>(std + r * G, s);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            double v = (double)s[k] * (double)s[k];
            double a = 0.5 * (double)beta - 0.5 / v;
            double b = (double)m[k] / v;
            A[k] = a;
            A[G + k] = b;
            amax = fmax(amax, fabs(a));
            bnorm2 += b * b;
            maxA = fmax(maxA, fmax(fabs(a), fabs(b)));
            double x2m = (double)meta[2 + k], xm = (double)meta[2 + G + k];
            p_naive += fabs(a) * x2m + fabs(b) * xm;
            double ls = lstd ? (double)lstd[r * G + k] : log((double)s[k]);
            double q = (xm + fabs((double)m[k]));
            e_ref += q * q / (2.0 * v) + fabs(ls) + 2.0 + (double)beta * x2m;
        }
        if (!(maxA > 0.0) || !isfinite(maxA) || !isfinite(e_ref)) bad = true;
    }
    double scale = 1.0;
    float w_scaled = -1.0f;
    if (!bad) {
        int ex;
        frexp(maxA, &ex);             // maxA = f * 2^ex, f in [0.5, 1)
        scale = ldexp(1.0, 8 - ex);   // max |A| * scale in [128, 256)
        double p_norm = amax * (double)meta[0] + sqrt(bnorm2) * (double)meta[1];
        double P = fmin(p_naive, p_norm);
        double sumB = 0.0, sumA = 0.0;
#pragma unroll
        for (int j = 0; j < 2 * G; ++j) { sumB += (double)meta[2 + j]; sumA += fabs(A[j]); }
        double e_tc = 1.01 * ldexp(P, -10) + ldexp(sumB / scale + sumA, -25);
        double W = 2.0 * (e_tc + ldexp(e_ref, -20));
        if (!(W <= (double)w_max) || !isfinite(W * scale)) bad = true;
        else w_scaled = (float)(W * scale * 1.001);
    }
#pragma unroll
    for (int k = 0; k < KPAD; ++k) {
        float v = 0.f;
        if (!bad && k < 2 * G) v = (float)(A[k] * scale);
        *reinterpret_cast<__half *>(rowbase + (k / 8) * 128 + (k % 8) * 2) = __float2half_rn(v);
    }
    wrow[r] = w_scaled;
}

// ---- 2. the tensor-core search ------------------------------------------------------------------
template <int KPAD>
struct TcSmem {
    static constexpr int kBTileBytes = kTileN * KPAD * 2;
    static constexpr int kABytes = kUnitRows * KPAD * 2;
    static constexpr int kBOff = 0;
    static constexpr int kAOff = kBOff + kStages * kBTileBytes;
    static constexpr int kBarOff = kAOff + 2 * kABytes;
    // barriers: full_b[kStages], empty_b[kStages], a_full[2], a_empty[2], acc_full[4], acc_empty[4]
    static constexpr int kNumBars = 2 * kStages + 4 + 8;
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kBytes = kSlotOff + 16;
};

struct Piece { int unit, t0, t1, seg; };

// CTA c owns steps [c*S, min((c+1)*S, total)); a step is (unit, tile) in unit-major order.
__device__ __forceinline__ bool next_piece(int64_t &step, int64_t end, int tiles, int spc, int cta, Piece &p)
{
    if (step >= end) return false;
    p.unit = (int)(step / tiles);
    p.t0 = (int)(step % tiles);
    int64_t room = end - step;
    p.t1 = (int)imin(tiles, p.t0 + room);
    p.seg = cta - (int)(((int64_t)p.unit * tiles) / spc);
    step += p.t1 - p.t0;
    return true;
}

// Running-max epilogue for one chunk of 32 consecutive columns held in v[0..31].
// Fast path: 16 FMNMX3/FMNMX and one compare.  Slow path (chunk maximum within the window of the
// running maximum, ~10 times per row over 2048 chunks): raise the running max / threshold and
// record (chunk max, chunk id); the rescoring kernel re-derives the candidates of the chunk.
__device__ __forceinline__ float chunk_max32(const float *v)
{
    float a0 = max3(v[0], v[1], v[2]), a1 = max3(v[3], v[4], v[5]), a2 = max3(v[6], v[7], v[8]);
    float a3 = max3(v[9], v[10], v[11]), a4 = max3(v[12], v[13], v[14]), a5 = max3(v[15], v[16], v[17]);
    float a6 = max3(v[18], v[19], v[20]), a7 = max3(v[21], v[22], v[23]), a8 = max3(v[24], v[25], v[26]);
    float a9 = max3(v[27], v[28], v[29]), a10 = fmaxf(v[30], v[31]);
    float b0 = max3(a0, a1, a2), b1 = max3(a3, a4, a5), b2 = max3(a6, a7, a8), b3 = fmaxf(a9, a10);
    return fmaxf(max3(b0, b1, b2), b3);
}

__device__ __forceinline__ void epi_record(float m, uint32_t chunk, float w, float &runmax, float &thr,
                                           uint32_t &cnt, Cand *__restrict__ list)
{
    if (m > runmax) { runmax = m; thr = m - w; }
    if (cnt < (uint32_t)kCap) list[cnt] = Cand{m, chunk};
    ++cnt;
}

template <int KPAD, bool DUMP>
__global__ void __launch_bounds__(kThreads, 1)
quantize_tc_kernel(const __half *__restrict__ a_img, const __half *__restrict__ b_img,
                   const float *__restrict__ wrow, SegHdr *__restrict__ seg_hdr, Cand *__restrict__ cand,
                   int64_t N, int tiles, int64_t total_steps, int steps_per_cta, int nseg,
                   float *__restrict__ dump)
{
    using L = TcSmem<KPAD>;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + L::kBarOff;
    // barrier addresses (shared-window u32), computed once
    auto full_b = [&](int i) { return bar_base + 8u * i; };
    auto empty_b = [&](int i) { return bar_base + 8u * (kStages + i); };
    auto a_full = [&](int i) { return bar_base + 8u * (2 * kStages + i); };
    auto a_empty = [&](int i) { return bar_base + 8u * (2 * kStages + 2 + i); };
    auto acc_full = [&](int i) { return bar_base + 8u * (2 * kStages + 4 + i); };   // [rb*2 + stage]
    auto acc_empty = [&](int i) { return bar_base + 8u * (2 * kStages + 8 + i); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kSlotOff);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int64_t step_begin = (int64_t)cta * steps_per_cta;
    const int64_t step_end = imin(step_begin + steps_per_cta, total_steps);

    if (warp == 1 && lane == 0) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
        for (int i = 0; i < 2 * kStages + 4; ++i) mbar_init(&bars[i], 1);           // full_b, empty_b, a_full, a_empty
        for (int i = 0; i < 4; ++i) mbar_init(&bars[2 * kStages + 4 + i], 1);       // acc_full
        for (int i = 0; i < 4; ++i) mbar_init(&bars[2 * kStages + 8 + i], 4);       // acc_empty: 4 epilogue warps
        fence_mbar_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: bulk async copies (TMA engine) =====
        if (lane == 0) {
            int64_t step = step_begin;
            Piece p;
            int stage = 0;
            uint32_t phase = 0;
            int pi = 0;
            while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
                const int ab = pi & 1;
                mbar_wait_a(a_empty(ab), ((pi >> 1) & 1) ^ 1);
                mbar_expect_tx_a(a_full(ab), L::kABytes);
                bulk_g2s_a(smem_base + L::kAOff + ab * L::kABytes,
                           reinterpret_cast<const char *>(a_img) + (int64_t)p.unit * L::kABytes, L::kABytes,
                           a_full(ab));
                for (int t = p.t0; t < p.t1; ++t) {
                    mbar_wait_a(empty_b(stage), phase ^ 1);
                    mbar_expect_tx_a(full_b(stage), L::kBTileBytes);
                    bulk_g2s_a(smem_base + L::kBOff + stage * L::kBTileBytes,
                               reinterpret_cast<const char *>(b_img) + (int64_t)t * L::kBTileBytes,
                               L::kBTileBytes, full_b(stage));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                ++pi;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread drives the tensor core =====
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_f32(kBlockRows, kTileN);
            constexpr uint32_t kLbo = 128;              // K-adjacent core matrices
            constexpr uint32_t kSbo = (KPAD / 8) * 128; // next 8-row group
            const uint32_t a_base = smem_base + L::kAOff;
            const uint32_t b_base = smem_base + L::kBOff;
            int64_t step = step_begin;
            Piece p;
            int stage = 0;
            uint32_t phase = 0;
            int as = 0;
            uint32_t aphase = 0;
            int pi = 0;
            while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
                const int ab = pi & 1;
                mbar_wait_a(a_full(ab), (pi >> 1) & 1);
                for (int t = p.t0; t < p.t1; ++t) {
                    mbar_wait_a(full_b(stage), phase);
                    tc_fence_after();
#pragma unroll
                    for (int rb = 0; rb < 2; ++rb) {
                        mbar_wait_a(acc_empty(rb * 2 + as), aphase ^ 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)((rb * kAccStages + as) * kTileN);
                        const uint32_t a_addr = a_base + ab * L::kABytes + rb * (kBlockRows * KPAD * 2);
                        const uint32_t b_addr = b_base + stage * L::kBTileBytes;
#pragma unroll
                        for (int j = 0; j < KPAD / 16; ++j) {
                            uint64_t ad = umma_desc_kmajor_noswz(a_addr + j * 256, kLbo, kSbo);
                            uint64_t bd = umma_desc_kmajor_noswz(b_addr + j * 256, kLbo, kSbo);
                            umma_f16_ss(d, ad, bd, idesc, j > 0 ? 1u : 0u);
                        }
                        umma_commit_a(acc_full(rb * 2 + as));
                    }
                    umma_commit_a(empty_b(stage));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                    as ^= 1;
                    if (as == 0) aphase ^= 1;
                }
                umma_commit_a(a_empty(ab));
                ++pi;
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===== epilogue: TMEM -> registers, running row max, chunk records =====
        const int ew = warp - kEpiWarp0;
        const int rb = ew >> 2;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const bool padded = (N % kTileN) != 0;
        int64_t step = step_begin;
        Piece p;
        int as = 0;
        uint32_t aphase = 0;
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            const int64_t row = (int64_t)p.unit * kUnitRows + rb * kBlockRows + q * 32 + lane;
            const float w = wrow[row];
            float runmax = -CUDART_INF_F;
            float thr = (w >= 0.f) ? -CUDART_INF_F : CUDART_INF_F;
            uint32_t cnt = 0;
            Cand *list = cand + (row * nseg + p.seg) * (int64_t)kCap;
            for (int t = p.t0; t < p.t1; ++t) {
                mbar_wait_a(acc_full(rb * 2 + as), aphase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                                       (uint32_t)((rb * kAccStages + as) * kTileN);
                float v[128];
                tmem_ld_32x128_sync(taddr, v);
                // the tile now lives in registers: hand the TMEM stage back before doing any math
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(acc_empty(rb * 2 + as));
                as ^= 1;
                if (as == 0) aphase ^= 1;

                if constexpr (DUMP) {
                    // debug build only: spill the raw tensor-core scores [M_pad, tiles*128]
                    float *drow = dump + row * ((int64_t)tiles * kTileN) + (int64_t)t * kTileN;
#pragma unroll
                    for (int i = 0; i < 128; ++i) drow[i] = v[i];
                }
                const uint32_t chunk0 = (uint32_t)t * (kTileN / kChunk);
                if (padded && t == tiles - 1) {
                    // last tile of a codebook that is not a multiple of 128: columns >= N hold the
                    // zero padding rows of the operand image and must not raise the running max
                    const int n_in_tile = (int)(N - (int64_t)t * kTileN);
#pragma unroll
                    for (int c = 0; c < kTileN / kChunk; ++c) {
                        float m = -CUDART_INF_F;
#pragma unroll
                        for (int i = 0; i < kChunk; ++i)
                            if (c * kChunk + i < n_in_tile) m = fmaxf(m, v[c * kChunk + i]);
                        if (m >= thr && c * kChunk < n_in_tile) epi_record(m, chunk0 + c, w, runmax, thr, cnt, list);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < kTileN / kChunk; ++c) {
                        const float m = chunk_max32(v + c * kChunk);
                        if (m >= thr) epi_record(m, chunk0 + c, w, runmax, thr, cnt, list);
                    }
                }
            }
            seg_hdr[row * nseg + p.seg] = SegHdr{runmax, cnt};
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- 3. exact re-scoring of the recorded chunks: one warp per row, lane = codeword -----------------
constexpr int kRescoreWarps = 4;

template <int G>
__global__ void __launch_bounds__(kRescoreWarps * 32)
rescore_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
               const float *__restrict__ cb, const float *__restrict__ nlpb, const float *__restrict__ wrow,
               const SegHdr *__restrict__ seg_hdr, const Cand *__restrict__ cand, int64_t M, int64_t N, int tiles,
               int steps_per_cta, int nseg, int64_t *__restrict__ idx_out, float *__restrict__ zhat_out,
               uint32_t *__restrict__ fb_rows, uint32_t *__restrict__ fb_count, uint32_t *__restrict__ stats)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * kRescoreWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kRescoreWarps;
    uint32_t st_rows = 0, st_resc = 0, st_fb = 0, st_ovf = 0;

    for (int64_t r = warp_global; r < M; r += nwarps) {
        const float w = wrow[r];
        bool fallback = !(w >= 0.f);
        bool overflow = false;
        int best_idx = 0x7fffffff;
        float best = -CUDART_INF_F;
        if (!fallback) {
            const int unit = (int)(r / kUnitRows);
            const int64_t s0 = (int64_t)unit * tiles;
            const int first = (int)(s0 / steps_per_cta);
            const int last = (int)((s0 + tiles - 1) / steps_per_cta);
            const int npieces = last - first + 1;
            float gmax = -CUDART_INF_F;
            for (int s = 0; s < npieces; ++s) {
                SegHdr h = seg_hdr[r * nseg + s];
                gmax = fmaxf(gmax, h.runmax);
                if (h.cnt > (uint32_t)kCap) overflow = true;
            }
            if (overflow || !(gmax > -CUDART_INF_F)) fallback = true;
            else {
                const float thr = gmax - w;
                float m[G], den[G], ls[G];
                load_row<G>(mu + r * G, m);
                {
                    float sd[G];
                    load_row<G>(std + r * G, sd);
#pragma unroll
                    for (int k = 0; k < G; ++k) {
                        den[k] = __fmul_rn(2.0f, __fmul_rn(sd[k], sd[k]));
                        ls[k] = lstd ? __ldg(lstd + r * G + k) : logf(sd[k]);
                    }
                }
                for (int s = 0; s < npieces; ++s) {
                    const uint32_t cnt = seg_hdr[r * nseg + s].cnt;
                    const Cand *list = cand + (r * nseg + s) * (int64_t)kCap;
                    for (uint32_t i = 0; i < cnt; ++i) {
                        const Cand c = list[i];            // warp-uniform
                        if (!(c.score >= thr)) continue;
                        const int64_t n = (int64_t)c.idx * kChunk + lane;
                        st_resc += 1;
                        if (n < N) {
                            float x[G], nb[G];
                            load_row<G>(cb + n * G, x);
                            load_row<G>(nlpb + n * G, nb);
                            const float sc = score_exact<G>(x, nb, m, den, ls);
                            if (sc > best || (sc == best && (int)n < best_idx)) { best = sc; best_idx = (int)n; }
                        }
                    }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
                    if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
                }
                if (best_idx == 0x7fffffff || !(best > -CUDART_INF_F)) fallback = true;  // NaN / -inf rows: let the scan decide
            }
        }
        if (fallback) {
            if (lane == 0) {
                uint32_t slot = atomicAdd(fb_count, 1u);
                fb_rows[slot] = (uint32_t)r;
            }
            st_fb += 1;
        } else {
            if (lane == 0) idx_out[r] = (int64_t)best_idx;
            if (zhat_out && lane < G) zhat_out[r * G + lane] = __ldg(cb + (int64_t)best_idx * G + lane);
        }
        st_rows += 1;
        st_ovf += overflow ? 1u : 0u;
    }
    if (stats && lane == 0) {
        atomicAdd(&stats[0], st_rows);
        atomicAdd(&stats[1], st_resc * 32u);
        atomicAdd(&stats[2], st_fb);
        atomicAdd(&stats[3], st_ovf);
    }
}

// ---- host side ------------------------------------------------------------------------------------
template <int G, int KPAD>
static int run_tc(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                  int64_t *idx, float *zhat, const TcPlan &p, const TcWorkspace &w, uint32_t *stats,
                  float w_max, cudaStream_t stream)
{
    using L = TcSmem<KPAD>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(quantize_tc_kernel<KPAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             L::kBytes);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
        attr_set = true;
    }
    cudaError_t e = cudaMemsetAsync(w.counters, 0, 256, stream);
    if (e != cudaSuccess) { set_error("memset: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }

    rows_prep_kernel<G, KPAD><<<(unsigned)((p.M_pad + 127) / 128), 128, 0, stream>>>(
        mu, std, lstd, cb->meta, w.a_img, w.wrow, M, p.M_pad, cb->beta, w_max);
    int rc = check_launch("rows_prep_kernel");
    if (rc) return rc;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (g_prof) {
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }
    quantize_tc_kernel<KPAD, false><<<p.ctas, kThreads, L::kBytes, stream>>>(
        w.a_img, cb->bimg, w.wrow, w.seg, w.cand, cb->N, p.tiles, p.total_steps, p.steps_per_cta, p.nseg, nullptr);
    if (g_prof) {
        cudaEventRecord(ev1, stream);
        g_prof_events.emplace_back(ev0, ev1);
    }
    rc = check_launch("quantize_tc_kernel");
    if (rc) return rc;

    {
        int64_t need = (M + kRescoreWarps - 1) / kRescoreWarps;
        int blocks = (int)imin(need, (int64_t)device_sm_count() * 16);
        rescore_kernel<G><<<blocks, kRescoreWarps * 32, 0, stream>>>(
            mu, std, lstd, cb->cb, cb->nlpb, w.wrow, w.seg, w.cand, M, cb->N, p.tiles, p.steps_per_cta, p.nseg, idx,
            zhat, w.fb_rows, w.counters, stats);
    }
    rc = check_launch("rescore_kernel");
    if (rc) return rc;
    count_launch(3);

    // exact scan over the rows the search could not certify (device-side count, no host sync)
    int fb_blocks = device_sm_count() * 2;
    return launch_quantize_exact(cb, mu, std, lstd, M, w.fb_rows, w.counters, idx, zhat, fb_blocks, stream);
}

// Debug/diagnostic path (tests only): runs row prep + the search with the DUMP epilogue and
// returns the raw fp16 tensor-core scores [M_pad, n_pad] plus the scaled windows [M_pad].
template <int G, int KPAD>
static int run_tc_dump(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                       const TcPlan &p, const TcWorkspace &w, float *dump, float *wrow_out, float w_max,
                       cudaStream_t stream)
{
    using L = TcSmem<KPAD>;
    cudaError_t e = cudaFuncSetAttribute(quantize_tc_kernel<KPAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::kBytes);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    rows_prep_kernel<G, KPAD><<<(unsigned)((p.M_pad + 127) / 128), 128, 0, stream>>>(
        mu, std, lstd, cb->meta, w.a_img, w.wrow, M, p.M_pad, cb->beta, w_max);
    int rc = check_launch("rows_prep_kernel");
    if (rc) return rc;
    quantize_tc_kernel<KPAD, true><<<p.ctas, kThreads, L::kBytes, stream>>>(
        w.a_img, cb->bimg, w.wrow, w.seg, w.cand, cb->N, p.tiles, p.total_steps, p.steps_per_cta, p.nseg, dump);
    rc = check_launch("quantize_tc_kernel<dump>");
    if (rc) return rc;
    e = cudaMemcpyAsync(wrow_out, w.wrow, sizeof(float) * p.M_pad, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) { set_error("memcpy: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    count_launch(2);
    return GQ_OK;
}

static float g_w_max = 3.0f;
void set_tc_window_max(float w) { g_w_max = w; }

int launch_quantize_tc(const gq_codebook *cb, const float *mu, const float *std, const float *lstd,
                       int64_t M, int64_t *idx, float *zhat, void *ws, int64_t ws_bytes, uint32_t *stats,
                       cudaStream_t stream)
{
    if (M == 0) return GQ_OK;
    if (M >= ((int64_t)1 << 31)) { set_error("M too large"); return GQ_ERR_INVALID; }
    if (device_cc_major() != 10) {
        set_error("gq_quantize needs an sm_100 device (tcgen05/TMEM); found cc major %d", device_cc_major());
        return GQ_ERR_UNSUPPORTED;
    }
    TcPlan p = make_plan(cb, M);
    TcWorkspace w = carve(cb, p, ws);
    if (ws == nullptr || ws_bytes < w.bytes) {
        set_error("workspace too small: need %lld bytes, got %lld", (long long)w.bytes, (long long)ws_bytes);
        return GQ_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) { set_error("workspace must be 256-byte aligned"); return GQ_ERR_INVALID; }
    switch (cb->g) {
    case 1: return run_tc<1, 16>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 2: return run_tc<2, 16>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 4: return run_tc<4, 16>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 8: return run_tc<8, 16>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 16: return run_tc<16, 32>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 32: return run_tc<32, 64>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    default: set_error("group %d unsupported", cb->g); return GQ_ERR_INVALID;
    }
}

int debug_tc_scores(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                    void *ws, int64_t ws_bytes, float *dump, float *wrow_out, cudaStream_t stream)
{
    if (device_cc_major() != 10) { set_error("needs sm_100"); return GQ_ERR_UNSUPPORTED; }
    TcPlan p = make_plan(cb, M);
    TcWorkspace w = carve(cb, p, ws);
    if (ws == nullptr || ws_bytes < w.bytes) { set_error("workspace too small"); return GQ_ERR_WORKSPACE; }
    switch (cb->g) {
    case 4: return run_tc_dump<4, 16>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
    case 8: return run_tc_dump<8, 16>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
    case 16: return run_tc_dump<16, 32>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
    default: set_error("debug dump supports g in {4,8,16}"); return GQ_ERR_INVALID;
    }
}

// Host-only: the stream-K decomposition for (N, g, M) on `sms` SMs, for CPU-side tests.
// out[0..5] = {units, tiles, ctas, steps_per_cta, nseg, M_pad}
void debug_plan(int64_t N, int64_t M, int sms, int64_t *out)
{
    TcPlan p = make_plan_sms(N, (N + kTileN - 1) / kTileN * kTileN, M, sms);
    out[0] = p.units; out[1] = p.tiles; out[2] = p.ctas; out[3] = p.steps_per_cta; out[4] = p.nseg; out[5] = p.M_pad;
}

}  // namespace gq
