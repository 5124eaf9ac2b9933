// gq_tc.cu -- the fused GaussianQuant quantizer for B200 (sm_100a).
//
// What it replaces: the eval branch of GaussianQuantRegularizer.forward, both backends
// (pit/quantization/gaussian.py:124-150): score every (row, codeword) pair, argmax per row,
// gather the winning codeword.  The reference materialises an [M, N] score matrix in HBM;
// here it lives only in tensor memory.
//
// Pipeline (all launches on the caller's stream):
//   1. rows_prep_kernel   per row: A[r,:] = s_r*[beta/2 - 1/(2v) | mu/v] as fp16, a RIGOROUS
//                         window W_r bounding twice the worst-case |fp16 tensor-core score -
//                         reference fp32 score| gap, and the (much smaller) margin of the fp32
//                         prefilter used by the rescoring kernel.
//   2. quantize_tc_kernel persistent, warp-specialised, 608 threads: a producer warp streams the
//                         prepared [x^2 | x] codebook image through a shared-memory ring (bulk
//                         async copies, TMA engine); one issuer warp per 128-row block issues
//                         tcgen05.mma (fp16 x fp16, M=128, N=128, K=2g, both operands from shared
//                         memory: the row operand is written there once per work unit) into that
//                         row block's two private TMEM accumulator stages (4 x 128 columns = all of
//                         tensor memory).  Plain fp16 images of groups 8 / 16 (the headline): FP16
//                         accumulators, read back packed (tcgen05.ld ... pack::16b, two scores per
//                         register) by epilogue warps that own one stage each and keep the running
//                         row maximum with packed three-input fp16 maxima (VHMNMX).  Split-precision
//                         images: fp32 accumulators, 32x64 half tiles, FMNMX3.  Either way the stage
//                         is handed back before any math, a 32-codeword chunk whose maximum comes
//                         within the row's window of the running maximum is recorded as (max,
//                         chunk id), and per-piece survivors are compacted into a 64-byte summary.
//                         (quantize_tc16_kernel: experimental TS-form variant, -DGQ_ACC16_TS=1.)
//   3. rescore_kernel     one warp per row: drop chunk records below the window under the final
//                         maximum, score the 32 codewords of every surviving chunk (lane =
//                         codeword) with an fp32 FMA form; if exactly one codeword lies within the
//                         prefilter margin of the best it is the answer, otherwise the row goes to
//   4. settle_kernel      (a few percent of the rows) which re-scores the contenders in the
//                         reference's exact fp32 op order (first-max tie-break).  Rows the search
//                         cannot certify (window too wide, lists overflowed, non-finite values) take
//                         the pruned exact scan along their sharpest dimension in the same kernel
//                         (pruned_exact_row); what that cannot decide goes to
//   5. exact_sliced_kernel (gq_kernels.cu), an exact-order full scan, through a device-side row list; for latent
//      calls whose rows are not already in final form it also writes the module's output tensors (FinalLayout).
//      Small calls (many segments per row) finish steps 3-4 inside rescore_many_kernel.
//
// Why the result is exactly the reference argmax: with E_r >= |G~[n]/s - (S_ref[n] - Q_r)| for
// all n (Q_r constant in n), the reference argmax n* satisfies
// G~[n*] >= G~[m] - 2 s E_r for every m, in particular for the running maxima the epilogue
// compares against (each is some G~[m]); so the chunk holding n* is always recorded (its maximum
// is >= G~[n*]), survives the final filter, and n* survives the prefilter (same argument with
// the prefilter's own bound) and wins the exact re-scoring.  DESIGN.md ("error window").
#include "gq_common.cuh"
#include "gq_internal.h"
#include "gq_ptx.cuh"

#include <utility>
#include <vector>

namespace gq {

// ---- optional per-launch timing of the search kernel (bench.py's roofline leg) ----------------
static int g_prof_period = 0;     // 0 = off, n = bracket every n-th search launch with an event pair
static unsigned g_prof_calls = 0;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
void prof_enable(int on) { g_prof_period = on > 0 ? on : 0; g_prof_calls = 0; }
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
#ifndef GQ_RING
#define GQ_RING 16
#endif
constexpr int kMaxStages = GQ_RING;  // codebook-tile ring depth, as far as ~190 KB of shared memory go (TcSmem::kStagesL)
constexpr int kAccStages = 4;        // TMEM accumulator stages: two private ones per row block (rb, rb + 2)
constexpr int kTmemCols = 512;       // 4 x 128 accumulator columns (the row operands live in shared memory)
#ifndef GQ_EPI_HALVES
#define GQ_EPI_HALVES 2
#endif
constexpr int kHalves = GQ_EPI_HALVES;   // warps per lane quarter draining one accumulator tile (128 / kHalves columns each)
constexpr int kNumEpiWarps = 8 * kHalves;
// Warp roles: 16 epilogue warps, then the service warps: the bulk-copy producer and kIssuers MMA issuer
// warps (the first of them also allocates tensor memory).  Four issuers = one per accumulator stage
// (row block, A/B): an issuer's per-tile chain -- barrier probe, descriptor arithmetic, two tcgen05.mma, two
// tcgen05.commit, loop bookkeeping, ~400 cycles of serial latency on the uniform datapath -- then has 512
// tensor cycles to hide behind instead of 256 (scripts/tc_prof.py timeline, profiles/README.md).
// Registers are allocated per scheduler (16 K each): 21 warps put six on one scheduler, i.e. at most 80
// registers per thread.  The fp16-accumulator kernel needs 72 and runs four issuers; the fp32-accumulator
// kernels (64 data registers per thread) keep two.
#ifndef GQ_ISSUERS
#define GQ_ISSUERS 2
#endif
__host__ __device__ constexpr int tc_issuers(bool acc16) { return acc16 ? GQ_ISSUERS : 2; }
__host__ __device__ constexpr int tc_threads(bool acc16) { return (kNumEpiWarps + 1 + tc_issuers(acc16)) * 32; }
static_assert(GQ_ISSUERS == 2 || GQ_ISSUERS == 4, "one issuer per row block, or one per accumulator stage");
constexpr int kEpiWarp0 = 0;         // epilogue warps 0..15
constexpr int kProducerWarp = kNumEpiWarps, kIssuerWarp0 = kNumEpiWarps + 1, kAllocWarp = kIssuerWarp0;
constexpr int kCap = 32;             // chunk records per (row, segment)
#ifndef GQ_CHUNK
#define GQ_CHUNK 32
#endif
constexpr int kChunk = GQ_CHUNK;     // columns per recorded chunk (32 or 16); the rescoring kernels gather kChunk codewords per record
static_assert(kChunk == 32 || kChunk == 16, "chunk width");
constexpr int kMinStepsPerCta = 8;
// rescore_kernel settling / pruning its rows itself instead of handing them to settle_kernel (one launch less) was
// measured at batch 64: 0.339 ms per step against 0.292 -- the rare rows, run out of line under the 64-register cap,
// turn their warps into stragglers of a kernel in which every warp walks ~14 rows.  Off; small calls
// (rescore_many_kernel) do finish their rows inline, where the launch is what costs.
#ifndef GQ_RESCORE_FINISHES
#define GQ_RESCORE_FINISHES 0
#endif
constexpr bool kRescoreFinishes = GQ_RESCORE_FINISHES != 0;
#ifndef GQ_FRONTEND_X4
#define GQ_FRONTEND_X4 1     // group 16, plain image: four lanes per row in the fused front-end (0: one thread per row)
#endif
#ifndef GQ_PROD_BATCH
#define GQ_PROD_BATCH 2
#endif
constexpr int kProdBatch = GQ_PROD_BATCH;   // codebook tiles the fp16-accumulator kernel's producer warp refills per loop trip
#ifndef GQ_ALT_TILES
#define GQ_ALT_TILES 1
#endif
constexpr bool kAltTiles = GQ_ALT_TILES != 0;   // fp16-accumulator epilogue: a warp owns one accumulator stage (0: a column half of both)

struct __align__(8) Cand { float score; uint32_t idx; };  // idx = chunk id (codeword / 32)

// fp16 accumulators add an error that is relative to the score itself: the second rounding is
// <= 2^-11 |G~|, the first <= 2^-11 |p1| with |p1| <= min(|G~| + s U2, s amax X2) (prep_row).  With
// d = 2^-11 (1 + 2^-9) the distance D = R - G~[n*] between a running maximum R and the reference argmax
// satisfies D <= 2 s E0 + d (2|R| + s B1) + d (2|G~[n*]| + s B1) and |G~[n*]| <= |R| + D, hence
// D <= w0 + kWinRel16 |R| with w0 = (2 s E0 + 2 d s B1) / (1 - 2d) (stored per row) and
// kWinRel16 = 4 d / (1 - 2d) < 2^-9 * 1.01.  fp32 accumulators: 0.
constexpr float kWinRel16 = 1.01f / 512.0f;
// candidate threshold below a (running or final) maximum: everything the reference argmax can hide behind
__device__ __forceinline__ float win_thr(float rmax, float w, float rel) { return rmax - fmaf(rel, fabsf(rmax), w); }

// Per (row, segment) summary written by the epilogue thread when its piece ends: the segment's running
// maximum and its records already filtered against that maximum (score >= runmax - w), so that the
// rescoring kernel fetches everything it needs about a segment with one 64-byte read.
// cnt: number of surviving records (<= kSurv), or kSegOverflow if the record list or the survivor
// slots overflowed (the row then goes to the exact scan).
constexpr int kSurv = 7;
constexpr uint32_t kSegOverflow = 0xffffffffu;
struct __align__(16) SegHdr { float runmax; uint32_t cnt; Cand surv[kSurv]; };
static_assert(sizeof(SegHdr) == 64, "segment summary must be 64 bytes");

struct TcPlan {
    int64_t M, M_pad, N;
    int unit_rows;       // rows per work unit: 256 (fp32 accumulators) or 128 (fp16 accumulators)
    int tile_cols;       // codewords per tile: 128 or 256
    int units;           // M_pad / unit_rows
    int tiles;           // codebook tiles per unit = ceil(N / tile_cols)
    int64_t total_steps; // units * tiles
    int ctas;
    int steps_per_cta;
    int nseg;            // storage segments per row: kHalves x (max CTAs touching one unit)
};

// Geometry of a search: the fp32-accumulator kernel works on units of 256 rows (two 128-row blocks per CTA)
// and 128-codeword tiles; the fp16-accumulator kernel on units of 128 rows (one row block per CTA, its two
// TMEM stages double-buffer 256-codeword tiles).  M_pad is a multiple of 256 either way, so the workspace
// layout does not depend on which kernel a call uses (the search mode may change from call to call).
static TcPlan make_plan_sms(int64_t N, int64_t M, int sms, bool acc16)
{
    TcPlan p;
    p.M = M;
    p.unit_rows = acc16 ? kBlockRows : kUnitRows;
    p.tile_cols = acc16 ? 2 * kTileN : kTileN;
    p.M_pad = (M + kUnitRows - 1) / kUnitRows * kUnitRows;
    p.units = (int)(p.M_pad / p.unit_rows);
    p.N = N;
    p.tiles = (int)((N + p.tile_cols - 1) / p.tile_cols);
    p.total_steps = (int64_t)p.units * p.tiles;
    int64_t max_ctas = (p.total_steps + kMinStepsPerCta - 1) / kMinStepsPerCta;
    p.ctas = (int)imin(sms, imax(1, max_ctas));
    p.steps_per_cta = (int)((p.total_steps + p.ctas - 1) / p.ctas);
    p.ctas = (int)((p.total_steps + p.steps_per_cta - 1) / p.steps_per_cta);
    p.nseg = kHalves * ((p.tiles + p.steps_per_cta - 1) / p.steps_per_cta + 1);
    return p;
}

static bool use_robust_image(const gq_codebook *cb);

// the plan of the NEXT call on this codebook (its search mode decides the kernel)
static bool plan_acc16(const gq_codebook *) { return false; }   // every kernel in use works on 256-row units and 128-codeword tiles
static TcPlan make_plan(const gq_codebook *cb, int64_t M) { return make_plan_sms(cb->N, M, device_sm_count(), plan_acc16(cb)); }

struct TcWorkspace {
    __half *a_img;      // [M_pad, kpad] fp16, row-major
    float *wrow;        // [M_pad] scaled window (negative => row goes to the exact fallback)
    float *margin;      // [M_pad] prefilter margin (unscaled score units)
    SegHdr *seg;        // [M_pad, nseg]
    Cand *cand;         // [M_pad, nseg, kCap]
    uint32_t *fb_rows;  // [M] rows the rescoring kernels could not certify (pruned exact scan inside settle_kernel); count = counters[0]
    uint32_t *fb_rows2; // [M] rows left for the sliced full scan (from the pruned scan and from settle_kernel); count = counters[2]
    uint32_t *st_rows;  // [M] rows whose prefilter left several contenders (settle_kernel); bit 31: use the full record list
    uint32_t *st_chunks; // [M][8] per listed row: {number of surviving chunks (<= 7, else ~0), chunk ids}: saves settle the re-derivation
    uint32_t *rec_cnt;  // [M_pad, nseg] records written per (row, segment) (may exceed kCap: list overflow)
    unsigned long long *fb_keys;  // [M] packed (score, ~index) argmax keys of the sliced exact scan
    uint32_t *counters; // [0] fallback count, [1] settle count, [64..] per-row-group arrival counters of the scan
    int64_t counter_bytes;
    int64_t bytes;
};

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

static TcWorkspace carve(const gq_codebook *cb, const TcPlan &p, void *base)
{
    TcWorkspace w;
    char *ptr = reinterpret_cast<char *>(base);
    int64_t off = 0;
    w.a_img = reinterpret_cast<__half *>(ptr + off);
    off = align_up(off + p.M_pad * imax(cb->kpad, cb->kpad_robust) * 2, 256);
    w.wrow = reinterpret_cast<float *>(ptr + off);
    off = align_up(off + p.M_pad * 4, 256);
    w.margin = reinterpret_cast<float *>(ptr + off);
    off = align_up(off + p.M_pad * 4, 256);
    w.seg = reinterpret_cast<SegHdr *>(ptr + off);
    off = align_up(off + p.M_pad * p.nseg * (int64_t)sizeof(SegHdr), 256);
    w.cand = reinterpret_cast<Cand *>(ptr + off);
    off = align_up(off + p.M_pad * p.nseg * (int64_t)kCap * (int64_t)sizeof(Cand), 256);
    w.fb_rows = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + p.M * 4, 256);
    w.fb_rows2 = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + p.M * 4, 256);
    w.st_rows = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + p.M * 4, 256);
    w.st_chunks = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + p.M * 32, 256);
    w.rec_cnt = reinterpret_cast<uint32_t *>(ptr + off);
    off = align_up(off + p.M_pad * p.nseg * 4, 256);
    w.fb_keys = reinterpret_cast<unsigned long long *>(ptr + off);
    off = align_up(off + p.M * 8, 256);
    w.counters = reinterpret_cast<uint32_t *>(ptr + off);
    w.counter_bytes = align_up(256 + (p.M / 8 + 2) * 4, 256);
    off += w.counter_bytes;
    w.bytes = off;
    return w;
}

int64_t tc_workspace_bytes(const gq_codebook *cb, int64_t M)
{
    if (M <= 0) return 256;
    // the search mode (and with it the kernel geometry) may change between calls: size for the larger plan
    const int sms = device_sm_count();
    int64_t b = carve(cb, make_plan_sms(cb->N, M, sms, false), nullptr).bytes;
    return b;
}

// ---- 1. row preparation -----------------------------------------------------------------------
// meta (from codebook prep): [0] = max_n ||x_n||^2, [1] = max_n ||x_n||, [2+j] = max_n |B[n,j]|,
// [2+2*kMaxG] = sum_j max_n |B[n,j]|.
// Window (unscaled score units), all terms upper bounds:
//   P     = min( sum_j |A_j| Bmax_j ,  max_k|a_k| * X2 + ||b||_2 * X1 )     >= sum_j |A_j B_nj|
//   E_tc  = 1.01 * 2^-10 * P                    fp16 rounding of A and of B (2^-11 each), the
//                                               product of both, fp32 tensor-core accumulation
//         + 2^-25 * (sum_j Bmax_j / s + sum_j |A_j|)      fp16 subnormal floors
//         + 2^-22 * sum_k[(beta/2 + 1/(2v_k)) X2max_k + |b_k| Xmax_k]   A itself is evaluated in fp32
//   E_ref = 2^-20 * R,  R = sum_k[ (Xmax_k+|mu_k|)^2/(2 v_k) + |log std_k| + 2 max(1,beta) + beta*X2max_k ]
//                                               rounding of the reference's own fp32 evaluation
//                                               (<= 7 roundings per term + <= 8 per summation path)
//   W     = 2 * (E_tc + E_ref)                  (x 1.01: the bound itself is evaluated in fp32)
//   E_pre = 3 * 2^-20 * R                       rounding of the rescoring kernel's fp32 FMA prefilter
//   margin = 2 * (E_pre + E_ref) * 1.01         (E_ref, E_pre doubled for g = 32)
template <int G, int KPAD, bool SPLIT>
__device__ __forceinline__ void prep_row(const float (&m)[G], const float (&s)[G], const float (&ls)[G],
                                         const float *__restrict__ meta, float beta, float w_max, bool valid,
                                         __half *__restrict__ a_row, float *__restrict__ w_out,
                                         float *__restrict__ margin_out, SegHdr *__restrict__ seg_row,
                                         uint32_t *__restrict__ rec_row, int nseg, uint32_t *__restrict__ wide_dev)
{
    // Segments no CTA will touch must read as empty.  Only the compact record counters are cleared
    // (nseg contiguous words per row); a summary whose counter is zero is stale memory from an earlier
    // call and is ignored by the readers -- clearing the 64-byte summaries themselves cost a partial
    // sector write per segment, a third of this kernel's traffic.
    (void)seg_row;
    if (valid) {
        for (int sgi = 0; sgi < nseg; ++sgi) rec_row[sgi] = 0u;
    }
    float A[2 * G];
    bool bad = !valid;
    float amax = 0.f, bnorm2 = 0.f, R = 0.f, p_naive = 0.f, maxA = 0.f, sumA = 0.f, eA = 0.f;
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const float v = s[k] * s[k];
        const float r = 0.5f / v;                  // 1/(2v)
        const float a = 0.5f * beta - r;
        const float b = m[k] / v;
        A[k] = a;
        A[G + k] = b;
        const float x2m = meta[2 + k], xm = meta[2 + G + k];
        amax = fmaxf(amax, fabsf(a));
        bnorm2 = fmaf(b, b, bnorm2);
        maxA = fmaxf(maxA, fmaxf(fabsf(a), fabsf(b)));
        p_naive += fabsf(a) * x2m + fabsf(b) * xm;
        sumA += fabsf(a) + fabsf(b);
        eA += (0.5f * beta + r) * x2m + fabsf(b) * xm;
        const float q = xm + fabsf(m[k]);
        R += q * q * r + fabsf(ls[k]) + 2.0f * fmaxf(1.0f, beta) + beta * x2m;
    }
    if (!(maxA > 0.f) || !isfinite(maxA) || !isfinite(R) || !isfinite(bnorm2)) bad = true;
    float scale = 1.f, w_scaled = -1.f, margin = 0.f;
    bool wide = false;
    constexpr bool ACC16 = gq_acc16(G, SPLIT);
    if (!bad) {
        int ex;
        frexpf(maxA, &ex);               // maxA = f * 2^ex, f in [0.5, 1)
        scale = ldexpf(1.0f, 8 - ex);    // max |A| * scale in [128, 256)
        const float u2 = sqrtf(bnorm2) * 1.000001f * meta[1];    // >= |sum_k b_k x_k|
        const float p_norm = amax * meta[0] + u2;
        const float P = fminf(p_naive, p_norm);
        if constexpr (ACC16) {
            // fp16 accumulators: every value the tensor core writes to TMEM (the x^2 partial sum and the score)
            // is bounded by s P (1 + 2^-10) -- keep that below 2^15 so nothing overflows fp16; the operands
            // themselves stay below 2^15 as well
            int exp_p;
            frexpf(P * 1.002f, &exp_p);
            scale = ldexpf(1.0f, imin(15 - exp_p, 15 - ex));
            if (!isfinite(P) || !(scale > 0.f) || !isfinite(scale)) bad = true;
        }
        const float sumB = meta[2 + 2 * kMaxG];
        // plain image: fp16 rounding of both operands dominates.  Split image (g <= 4): what is left is
        // the rounding of the two lo parts (2^-22 each), the dropped lo*lo product (2^-22) and the tensor
        // core's fp32 accumulation of up to 32 products, bounded by 2^-18 P (32 truncating adds of 2^-23
        // relative each; measured worst case on B200 is 25x below the resulting window: tests/diag/tc_diag.py
        // dump_g4 reports max |error| / (W/2) = 0.02 with the 2^-17 variant of this bound).
        const float e_round = SPLIT ? 1.01f * (ldexpf(P, -18) + ldexpf(P, -20)) : 1.01f * ldexpf(P, -10);
        const float e_floor = ldexpf(sumB / scale + sumA, SPLIT ? -24 : -25);   // fp16 subnormal floors (hi and lo parts)
        // fp16 accumulators: the absolute part of the accumulator-rounding bound (kWinRel16 above), B1 = a bound
        // on the x^2 partial sum that does not involve the score: min(U2, amax X2); plus the fp16 subnormal
        // spacing of the stored values themselves (2^-25 in scaled units, twice)
        const float e_acc = ACC16 ? 1.004f * ldexpf(fminf(u2, amax * meta[0]), -11) + ldexpf(1.0f / scale, -24) : 0.f;
        const float e_tc = e_round + e_floor + ldexpf(eA, -22) + e_acc;
        constexpr float kE = (G > 16) ? 2.0f : 1.0f;   // deeper fp32 summation trees for g = 32
        const float W = 2.02f * (e_tc + kE * ldexpf(R, -20));
        margin = 2.02f * kE * 4.0f * ldexpf(R, -20);
        // The cap only routes rows whose candidate lists would get long to the exact scan (correctness never
        // depends on it).  With the split image a window this wide means an extremely sharp row (variance
        // ~1e-4 in some dimension) whose neighbours are far apart in score, so the cap can be 32x higher:
        // measured on the stress latents, 8 % -> 1 % of the rows left for the exact scan.
        const float w_cap = SPLIT ? 32.0f * w_max : w_max;
        if (!(W <= w_cap) || !isfinite(W * scale)) bad = true;
        else w_scaled = W * scale * 1.001f;
        if constexpr (gq_has_robust_image(G)) {
            // GQ_MODE_AUTO feedback: rows whose PLAIN-image window exceeds the cap (they take the exact scan
            // in fast mode), counted in either mode with one atomic per warp
            const float w_plain = 2.02f * (1.01f * ldexpf(P, -10) + ldexpf(sumB / scale + sumA, -25) + ldexpf(eA, -22) + kE * ldexpf(R, -20)
                                           + (gq_acc16(G, false) ? 1.004f * ldexpf(fminf(u2, amax * meta[0]), -11) : 0.f));
            wide = !(w_plain <= w_max);
        }
    }
    if constexpr (gq_has_robust_image(G)) {
        const unsigned act = __activemask();
        const unsigned mw = __ballot_sync(act, wide);
        if (mw && (threadIdx.x & 31) == (unsigned)(__ffs(mw) - 1)) atomicAdd(wide_dev, (uint32_t)__popc(mw));
    }
    // fp16 row, 16-byte stores.  Plain image: [A | 0]; split image: [hi(A) | hi(A) | lo(A) | 0].
    auto image_value = [&](int k) -> float {
        if (bad) return 0.f;
        if (!SPLIT) return k < 2 * G ? A[k < 2 * G ? k : 0] * scale : 0.f;
        const int part = k / (2 * G), j = k % (2 * G);
        if (part >= 3) return 0.f;
        const float v = A[j] * scale;                                   // exact: scale is a power of two
        const float hi = __half2float(__float2half_rn(v));
        return part == 2 ? __fsub_rn(v, hi) : hi;                       // v - hi is exact in fp32
    };
    uint4 *dst = reinterpret_cast<uint4 *>(a_row);
#pragma unroll
    for (int c = 0; c < KPAD / 8; ++c) {
        __half2 h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(image_value(c * 8 + 2 * i), image_value(c * 8 + 2 * i + 1));
        dst[c] = *reinterpret_cast<uint4 *>(h);
    }
    *w_out = w_scaled;
    *margin_out = margin;
}

template <int G, int KPAD, bool SPLIT>
__global__ void __launch_bounds__(128)
rows_prep_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
                 const float *__restrict__ meta, __half *__restrict__ a_img, float *__restrict__ wrow,
                 float *__restrict__ margin, SegHdr *__restrict__ seg_hdr, uint32_t *__restrict__ rec_cnt, int nseg,
                 int64_t M, int64_t M_pad, float beta, float w_max, uint32_t *__restrict__ counters, int64_t counter_words,
                 uint32_t *__restrict__ wide_dev)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < counter_words) counters[r] = 0u;   // device-side lists / arrival counters of this call (counter_words <= M_pad)
    if (r >= M_pad) return;
    float m[G], s[G], ls[G];
    const bool valid = r < M;
    if (valid) {
        load_row<G>(mu + r * G, m);
        load_row<G>(std + r * G, s);
        if (lstd) load_row<G>(lstd + r * G, ls);
        else {
#pragma unroll
            for (int k = 0; k < G; ++k) ls[k] = logf(s[k]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < G; ++k) { m[k] = 0.f; s[k] = 1.f; ls[k] = 0.f; }
    }
    prep_row<G, KPAD, SPLIT>(m, s, ls, meta, beta, w_max, valid, a_img + r * KPAD, wrow + r, margin + r,
                      seg_hdr + r * nseg, rec_cnt + r * nseg, nseg, wide_dev);
}

// ---- 2. the tensor-core search ------------------------------------------------------------------
template <int KPAD>
struct TcSmem {
    static constexpr int kBTileBytes = kTileN * KPAD * 2;
    // Ring depth.  The two row blocks drift apart by several tiles (each has its own accumulator handshake),
    // and a slot is only free once BOTH have consumed it, so the prefetch distance that is left is the depth
    // minus the drift; a bulk copy needs ~700 cycles from issue to landing (scripts/tc_prof.py).  Eight
    // slots left the issuers waiting for tiles; sixteen cover it.  Even, so that a slot keeps its stage parity.
    static constexpr int kFit = (190 * 1024 - 2 * (kBlockRows * KPAD * 2)) / kBTileBytes;
    static constexpr int kStagesL = (kFit < kMaxStages ? kFit : kMaxStages) & ~1;
    static constexpr int kBOff = 0;
    // row operand of each of the two row blocks, same canonical K-major core-matrix layout as a B tile
    static constexpr int kABytes = kBlockRows * KPAD * 2;
    static constexpr int kAOff = kBOff + kStagesL * kBTileBytes;
    static constexpr int kBarOff = kAOff + 2 * kABytes;
    // barriers: full_b[kStages], empty_b[kStages], a_full[2], acc_full[kAccStages], acc_empty[kAccStages]
    static constexpr int kNumBars = 2 * kStagesL + 2 + 2 * kAccStages;
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

__device__ __forceinline__ void tmem_ld_cols_sync(uint32_t taddr, float (&v)[64]) { tmem_ld_32x64_sync(taddr, v); }
__device__ __forceinline__ void tmem_ld_cols_sync(uint32_t taddr, float (&v)[128]) { tmem_ld_32x128_sync(taddr, v); }

// Maximum of 32 consecutive columns: 14 FMNMX3 + 2 FMNMX.
__device__ __forceinline__ float chunk_max32(const float *v)
{
#ifdef GQ_EXP_NOREDUCE
    return fmaxf(v[0], v[31]);   // timing experiment only (results are wrong): how fast is the pipeline without the reduction?
#endif
    if constexpr (kChunk == 16) {   // 7 FMNMX3 + 1 FMNMX
        const float c0 = max3(v[0], v[1], v[2]), c1 = max3(v[3], v[4], v[5]), c2 = max3(v[6], v[7], v[8]);
        const float c3 = max3(v[9], v[10], v[11]), c4 = max3(v[12], v[13], v[14]);
        return fmaxf(max3(c0, c1, c2), max3(c3, c4, v[15]));
    }
    float a0 = max3(v[0], v[1], v[2]), a1 = max3(v[3], v[4], v[5]), a2 = max3(v[6], v[7], v[8]);
    float a3 = max3(v[9], v[10], v[11]), a4 = max3(v[12], v[13], v[14]), a5 = max3(v[15], v[16], v[17]);
    float a6 = max3(v[18], v[19], v[20]), a7 = max3(v[21], v[22], v[23]), a8 = max3(v[24], v[25], v[26]);
    float a9 = max3(v[27], v[28], v[29]), a10 = fmaxf(v[30], v[31]);
    float b0 = max3(a0, a1, a2), b1 = max3(a3, a4, a5), b2 = max3(a6, a7, a8), b3 = fmaxf(a9, a10);
    return fmaxf(max3(b0, b1, b2), b3);
}

// Same over the first n (<= 32, possibly <= 0) registers only; -inf when n <= 0.
__device__ __forceinline__ float chunk_max32_masked(const float *v, int n)
{
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < kChunk; ++i)
        if (i < n) m = fmaxf(m, v[i]);
    return m;
}

// Slow path (~10 times per row over 2048 chunks): raise the running max / threshold and record
// (chunk max, chunk id); the rescoring kernel re-derives the candidates of the chunk.
template <bool ACC16>
__device__ __forceinline__ void epi_record_t(float m, uint32_t chunk, float w, float &runmax, float &thr,
                                             uint32_t &cnt, Cand *__restrict__ list)
{
    constexpr float rel = ACC16 ? kWinRel16 : 0.f;
    if (m > runmax) { runmax = m; thr = win_thr(m, w, rel); }
    if (cnt < (uint32_t)kCap) list[cnt] = Cand{m, chunk};
    ++cnt;
}

// One row of the row-operand image (KPAD fp16, row-major in global memory) into the shared-memory
// operand of its row block: canonical K-major no-swizzle layout, core matrix = 8 rows x 16 bytes,
// K-adjacent core matrices 128 bytes apart, 8-row groups (KPAD / 8) * 128 bytes apart -- the layout
// of a codebook tile, so both descriptors share LBO / SBO.
template <int KPAD>
__device__ __forceinline__ void store_a_row_smem(uint32_t a_smem, int r, const __half *__restrict__ a_row)
{
    const uint32_t base = a_smem + (uint32_t)(r >> 3) * (uint32_t)((KPAD / 8) * 128) + (uint32_t)(r & 7) * 16u;
#pragma unroll
    for (int j = 0; j < KPAD / 8; ++j) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(a_row) + j);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)j * 128u), "r"(q.x), "r"(q.y), "r"(q.z), "r"(q.w) : "memory");
    }
}

#ifdef GQ_TC_TRACE
// diagnostic build only (scripts/tc_prof.py): clock64 timeline of CTA 0, iterations [kTraceFrom, +kTraceN)
#ifndef GQ_TRACE_FROM
#define GQ_TRACE_FROM 300   // first traced iteration (0 for small calls: scripts/tc_prof.py build -DGQ_TRACE_FROM=0)
#endif
constexpr int kTraceFrom = GQ_TRACE_FROM, kTraceN = 24;
__device__ unsigned int g_tc_trace[32 * kTraceN * 4];
#define GQ_TRACE_DECL() uint32_t tr_it_ = 0; const long long tr_t0_ = clock64()
#define GQ_TRACE(slot) do { if (cta == 0 && lane == 0 && tr_it_ >= kTraceFrom && tr_it_ < kTraceFrom + kTraceN) \
        g_tc_trace[(warp * kTraceN + (tr_it_ - kTraceFrom)) * 4 + (slot)] = (unsigned int)(clock64() - tr_t0_); } while (0)
#define GQ_TRACE_NEXT() (++tr_it_)
#define GQ_TRACE_SET(x) (tr_it_ = (x))
#else
#define GQ_TRACE_DECL()
#define GQ_TRACE(slot)
#define GQ_TRACE_NEXT()
#define GQ_TRACE_SET(x)
#endif

// Maximum of the 64 scores a thread holds as 32 packed fp16 pairs: low halves = chunk 2h, high halves =
// chunk 2h+1 of the warp's half tile (gq_acc16_image_row), so the result is (chunk max, chunk max).
// 15 VHMNMX + 1 HMNMX2.
__device__ __forceinline__ uint32_t chunk_max_packed(const uint32_t *v)
{
#ifdef GQ_EXP_NOREDUCE
    return hmax2(v[0], v[31]);   // timing experiment only (results are wrong)
#endif
    uint32_t a[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = hmax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    a[10] = hmax2(v[30], v[31]);
    const uint32_t b0 = hmax3(a[0], a[1], a[2]), b1 = hmax3(a[3], a[4], a[5]), b2 = hmax3(a[6], a[7], a[8]);
    return hmax3(hmax3(b0, b1, b2), a[9], a[10]);
}
__device__ __forceinline__ float half_lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xffffu))); }
__device__ __forceinline__ float half_hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }
// (thr, thr) as fp16, rounded DOWN: a packed compare against it never misses a value >= thr
__device__ __forceinline__ uint32_t pack_thr(float thr)
{
    const uint32_t h = __half_as_ushort(__float2half_rd(thr));
    return h | (h << 16);
}
// nonzero iff either half of m is >= the matching half of t
__device__ __forceinline__ uint32_t hge2_any(uint32_t m, uint32_t t)
{
    return __hge2_mask(*reinterpret_cast<const __half2 *>(&m), *reinterpret_cast<const __half2 *>(&t));
}

template <int KPAD, bool DUMP, bool ACC16>
__global__ void __launch_bounds__(tc_threads(ACC16), 1)
quantize_tc_kernel(const __half *__restrict__ a_img, const __half *__restrict__ b_img,
                   const float *__restrict__ wrow, SegHdr *__restrict__ seg_hdr, Cand *__restrict__ cand,
                   uint32_t *__restrict__ rec_cnt, int64_t N, int tiles, int64_t total_steps, int steps_per_cta,
                   int nseg, float *__restrict__ dump)
{
    using L = TcSmem<KPAD>;
    constexpr int kStages = L::kStagesL;
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t smem_base;
    // pin the shared-window base in a register (otherwise it is rematerialised from SR_CgaCtaId
    // inside every wait loop)
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_base) : "r"(smem_u32(smem)));
    const uint32_t bar_base = smem_base + L::kBarOff;
    auto full_b = [&](int i) { return bar_base + 8u * i; };
    auto empty_b = [&](int i) { return bar_base + 8u * (kStages + i); };
    auto a_full = [&](int i) { return bar_base + 8u * (2 * kStages + i); };                      // per row block
    // Accumulator stages are private to a row block (rb owns stages rb and rb + 2): a classic double
    // buffer per consumer group.  acc_full(s): the MMA into stage s is complete.  acc_empty(s): the
    // eight epilogue warps of the owning row block have pulled stage s into registers.
    auto acc_full = [&](int st) { return bar_base + 8u * (2 * kStages + 2 + st); };
    auto acc_empty = [&](int st) { return bar_base + 8u * (2 * kStages + 2 + kAccStages + st); };
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kSlotOff);

    // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so that everything the
    // service warps derive from it -- stage counters, barrier and descriptor addresses -- lives in uniform
    // registers instead of being moved there with R2UR in front of every tcgen05 instruction)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int64_t step_begin = (int64_t)cta * steps_per_cta;
    const int64_t step_end = imin(step_begin + steps_per_cta, total_steps);
    mbar_probe_declare();
    // One persistent wave: the successor's CTAs cannot become resident before these exit anyway, so its
    // launch may be set up at once.  Barrier init, the TMEM allocation and the first codebook tiles (static
    // data) do not depend on the predecessor and overlap its tail; only the epilogue warps consume what
    // it wrote (row operands, windows, initialised segment headers) and wait for it below.
    pdl_launch_dependents();
    GQ_TRACE_DECL();

    if (warp == kIssuerWarp0 && lane == 0) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);                                  // full_b
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[kStages + i], 2);                        // empty_b: one arrival per row block
        for (int i = 0; i < 2; ++i) mbar_init(&bars[2 * kStages + i], 4);                          // a_full: 4 warps
        for (int i = 0; i < kAccStages; ++i) mbar_init(&bars[2 * kStages + 2 + i], 1);                           // acc_full
        for (int i = 0; i < kAccStages; ++i) mbar_init(&bars[2 * kStages + 2 + kAccStages + i], (ACC16 && kAltTiles) ? 4 : 4 * kHalves);     // acc_empty: the warps draining the stage
        fence_mbar_init();
    }
    if (warp == kAllocWarp) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kProducerWarp) {
        // ===== producer: bulk async copies (TMA engine) of codebook tiles =====
        // The whole warp runs the loop (warp-uniform control flow keeps addresses on the uniform
        // datapath); one elected lane arms the barrier and issues the copy.
        int64_t step = step_begin;
        Piece p;
        int stage = 0;
        uint32_t phase = 0;
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            for (int t = p.t0; t < p.t1; ++t) {
                GQ_TRACE(0);
                mbar_wait_a(empty_b(stage), phase ^ 1);
                GQ_TRACE(1);
                if (elect_one()) {
#ifdef GQ_EXP_HALFCOPY
                    // timing experiment only (results are wrong): every second tile is not copied at all
                    if (t & 1) mbar_arrive_a(full_b(stage)); else {
#endif
                    mbar_expect_tx_a(full_b(stage), L::kBTileBytes);
                    bulk_g2s_a(smem_base + L::kBOff + stage * L::kBTileBytes,
                               reinterpret_cast<const char *>(b_img) + (int64_t)t * L::kBTileBytes,
                               L::kBTileBytes, full_b(stage));
#ifdef GQ_EXP_HALFCOPY
                    }
#endif
                }
                __syncwarp();
                GQ_TRACE(2);
                GQ_TRACE(3);
                GQ_TRACE_NEXT();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= kIssuerWarp0 && warp < kIssuerWarp0 + tc_issuers(ACC16)) {
        // ===== MMA issuers.  Issuer iw serves row block rb = iw & 1; with four issuers it takes every second
        // tile of that row block -- the ones that land in accumulator stage rb + 2 * (iw >> 1).  The whole warp
        // runs the loop (warp-uniform control flow keeps the addresses on the uniform datapath), one elected
        // lane issues to the tensor core.  Tiles are numbered by a per-CTA counter c that runs across pieces:
        // codebook-tile ring slot c % kStages, accumulator stage parity c & 1. =====
        const int iw = warp - kIssuerWarp0;
        const int rb = iw & 1;
        constexpr int kStride = tc_issuers(ACC16) / 2;       // tiles between two uses of this issuer
        const int sub0 = kStride == 2 ? (iw >> 1) : 0;
        constexpr uint32_t idesc = umma_idesc_f16(kBlockRows, kTileN, /*d_f32=*/!ACC16);
        constexpr uint32_t kLbo = 128;              // K-adjacent core matrices
        constexpr uint32_t kSbo = (KPAD / 8) * 128; // next 8-row group
        // descriptor words: everything but the 14-bit start address is loop invariant
        constexpr uint32_t desc_hi = ((kSbo >> 4) & 0x3FFFu) | (1u << 14);           // SBO, version = 1
        const uint32_t desc_lo0 = (((smem_base + L::kBOff) >> 4) & 0x3FFFu) | (((kLbo >> 4) & 0x3FFFu) << 16);
        const uint32_t adesc_lo0 = (((smem_base + L::kAOff + rb * L::kABytes) >> 4) & 0x3FFFu) | (((kLbo >> 4) & 0x3FFFu) << 16);
        static_assert(kStages % 2 == 0, "ring slots keep their accumulator parity");
        int64_t step = step_begin;
        Piece p;
        uint32_t c = 0;       // tile counter of this CTA at the start of the piece
        int pi = 0;
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            mbar_wait_a(a_full(rb), pi & 1);
            tc_fence_after();
            const uint32_t c_end = c + (uint32_t)(p.t1 - p.t0);
            uint32_t ci = kStride == 2 ? c + ((c ^ (uint32_t)sub0) & 1u) : c;   // first tile of the piece that is ours
            for (; ci < c_end; ci += kStride) {
                const uint32_t stage = ci % (uint32_t)kStages;
                const uint32_t phase = (ci / (uint32_t)kStages) & 1u;
                const uint32_t as = (uint32_t)rb + 2u * (ci & 1u);
                const uint32_t ep = ((ci >> 1) & 1u) ^ 1u;   // u-th use of a stage (u = ci / 2) waits for its (u-1)-th drain
                GQ_TRACE_SET(ci);
                GQ_TRACE(0);
                mbar_wait2_a(full_b(stage), phase, acc_empty(as), ep);
                GQ_TRACE(1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + as * (uint32_t)kTileN;
                    const uint32_t desc_lo = desc_lo0 + stage * (uint32_t)(L::kBTileBytes >> 4);
#pragma unroll
                    for (int j = 0; j < KPAD / 16; ++j) {
                        // one K = 16 slice = two K-adjacent core matrices = 256 bytes in either operand
                        const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)(adesc_lo0 + j * 16);
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo + j * 16);
                        umma_f16_ss(d, ad, bd, idesc, j > 0 ? 1u : 0u);
                    }
                    GQ_TRACE(2);
                    umma_commit_a(acc_full(as));
                    umma_commit_a(empty_b(stage));   // 2 arrivals (one per row block) free the B slot
                }
                __syncwarp();
                GQ_TRACE(3);
                GQ_TRACE_NEXT();
            }
            c = c_end;
            ++pi;
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + kNumEpiWarps) {
        // ===== epilogue: TMEM -> registers, running row max, chunk records =====
        // 16 warps: (row block, column half, lane quarter).  Thread <-> (row, column half): each row is
        // watched by two threads with independent running maxima / record lists ("storage segments"
        // 2*piece + half); four epilogue warps per scheduler hide the barrier and TMEM-load latencies.
        // (Tried and measured slower, see DESIGN.md 4.4: the two warps of a row taking alternate tiles
        // instead of column halves; a software-pipelined loop of 32-column loads; four issuer warps.)
        const int ew = warp - kEpiWarp0;
        const int rb = (ew >> 2) & 1;
        const int half = ew >> 3;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const bool padded = (N % kTileN) != 0;
        constexpr int kHalfCols = kTileN / kHalves;
        int64_t step = step_begin;
        Piece p;
        if constexpr (ACC16 && kAltTiles) {
        // ---- fp16 accumulators: (row block, TILE PARITY, lane quarter).  With packed loads a whole 128-column
        // accumulator tile is 64 registers, so a warp can own ONE accumulator stage (`half` = 0: stage A = the
        // even tiles of the CTA's running tile count, 1: stage B) instead of a column half of both.  Its serial
        // chain per use -- barrier, tcgen05.ld, release, two independent 32-register max trees, threshold test --
        // then has two tile periods to hide behind, and the next barrier it needs has usually completed long
        // before it asks (the timeline of the column-half layout showed ~290 cycles of chain per 256-cycle
        // tile period).  Thread <-> (row, tile parity): storage segment 2 * piece + parity.
        static_assert(kHalves == 2 && kChunk == 32, "alternate-tile epilogue");
        constexpr uint32_t kEmptyOff2 = 8u * kAccStages;
        const uint32_t as = (uint32_t)rb + 2u * (uint32_t)half;
        const uint32_t fbar = acc_full(as);
        const uint32_t obar = acc_full(rb + 2 * (half ^ 1));   // the row block's other stage
        const uint32_t taddr = tmem_base + lane_base + as * (uint32_t)kTileN;
        int pad_tile;
        asm volatile("mov.b32 %0, %1;" : "=r"(pad_tile) : "r"(padded ? tiles - 1 : -1));
        uint32_t c = 0;   // tiles this CTA has processed before the current piece
        pdl_wait();       // the predecessor's rows / windows / segment headers are complete and visible from here on
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            const int64_t row = (int64_t)p.unit * kUnitRows + rb * kBlockRows + q * 32 + lane;
            const int sseg = p.seg * kHalves + half;
            const uint32_t c_end = c + (uint32_t)(p.t1 - p.t0);
            if (half == 0) {
                // Row operand into shared memory.  Every MMA of the previous piece that read this row block's
                // operand must have completed.  This warp consumed that piece's even tiles; if its last tile was
                // an odd one, wait for that accumulator as well (a pure wait: the stage-B warps drain it).
                if (c > 0u && ((c - 1u) & 1u)) mbar_wait_a(obar, ((c - 1u) >> 1) & 1u);
                store_a_row_smem<KPAD>(smem_base + L::kAOff + rb * L::kABytes, q * 32 + lane, a_img + row * KPAD);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_full(rb));
            }
            const float w = wrow[row];
            float runmax = -CUDART_INF_F;
            float thr = (w >= 0.f) ? -CUDART_INF_F : CUDART_INF_F;
#ifdef GQ_EXP_NOREC
            thr = CUDART_INF_F;   // timing experiment only: the record path is never taken (results are wrong)
#endif
            uint32_t thr2 = pack_thr(thr);
            uint32_t cnt = 0;
            const uint32_t seg_index = (uint32_t)(row * nseg + sseg);   // < 2^32, checked on the host
            Cand *list = cand + (int64_t)seg_index * kCap;
            uint32_t ci = c + ((c ^ (uint32_t)half) & 1u);              // first tile of the piece with our parity
            int t = p.t0 + (int)(ci - c);
            for (; ci < c_end; ci += 2u, t += 2) {
                const uint32_t par = (ci >> 1) & 1u;                    // this is use number ci / 2 of the stage
                GQ_TRACE_SET(ci);
                GQ_TRACE(0);
                mbar_probe_wait_a(fbar, par);
                GQ_TRACE(1);
                tc_fence_after();
                // register i < 32: (codeword i, codeword 32 + i) of the tile; register 32 + i: (64 + i, 96 + i)
                uint32_t pv[64];
                tmem_ld_32x128p_sync(taddr, pv);
                GQ_TRACE(2);
                tc_fence_before();
                if (lane == 0) mbar_arrive_a(fbar + kEmptyOff2);
                if constexpr (DUMP) {
                    float *drow = dump + row * ((int64_t)tiles * kTileN) + (int64_t)t * kTileN;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        drow[i] = half_lo(pv[i]); drow[32 + i] = half_hi(pv[i]);
                        drow[64 + i] = half_lo(pv[32 + i]); drow[96 + i] = half_hi(pv[32 + i]);
                    }
                }
                if (t == pad_tile) {
                    // columns >= N hold the zero rows of the operand image: replace them with -inf
                    const int n_in = (int)(N - (int64_t)t * kTileN);
#pragma unroll
                    for (int i = 0; i < 64; ++i) {
                        const int lo_cw = (i & 32) * 2 + (i & 31);
                        if (lo_cw >= n_in) pv[i] = (pv[i] & 0xffff0000u) | 0xfc00u;
                        if (lo_cw + 32 >= n_in) pv[i] = (pv[i] & 0x0000ffffu) | 0xfc000000u;
                    }
                }
                const uint32_t ma = chunk_max_packed(pv);        // (chunk 4t, chunk 4t + 1)
                mbar_probe_a(fbar, par ^ 1u);                    // the stage's next use: answered long before we ask again
                const uint32_t mb = chunk_max_packed(pv + 32);   // (chunk 4t + 2, chunk 4t + 3)
                if (hge2_any(hmax2(ma, mb), thr2)) {
                    const float m0 = half_lo(ma), m1 = half_hi(ma), m2 = half_lo(mb), m3 = half_hi(mb);
                    const uint32_t c0 = (uint32_t)t * (kTileN / kChunk);
                    if (m0 >= thr) epi_record_t<ACC16>(m0, c0, w, runmax, thr, cnt, list);
                    if (m1 >= thr) epi_record_t<ACC16>(m1, c0 + 1u, w, runmax, thr, cnt, list);
                    if (m2 >= thr) epi_record_t<ACC16>(m2, c0 + 2u, w, runmax, thr, cnt, list);
                    if (m3 >= thr) epi_record_t<ACC16>(m3, c0 + 3u, w, runmax, thr, cnt, list);
                    thr2 = pack_thr(thr);
                }
                GQ_TRACE(3);
            }
            c = c_end;
            // piece done: filter this thread's own records against its final maximum and publish the survivors
            {
                SegHdr *hp = &seg_hdr[seg_index];
                uint32_t ns = 0;
                bool ovf = cnt > (uint32_t)kCap;
                const float keep = win_thr(runmax, w, kWinRel16);
                const uint32_t nrec = cnt < (uint32_t)kCap ? cnt : (uint32_t)kCap;
                for (uint32_t i = 0; i < nrec; ++i) {
                    const Cand cd = list[i];
                    if (cd.score >= keep) {
                        if (ns < (uint32_t)kSurv) hp->surv[ns] = cd;
                        else ovf = true;
                        ++ns;
                    }
                }
                hp->runmax = runmax;
                hp->cnt = ovf ? kSegOverflow : ns;
                rec_cnt[seg_index] = cnt;
            }
        }
        } else {
        // This row block's two private stages A = rb and B = rb + 2: barrier address, TMEM address and
        // the phase parity of the next use.  acc_empty(s) sits at a constant offset from acc_full(s).
        constexpr uint32_t kEmptyOff = 8u * kAccStages;
        const uint32_t fa = acc_full(rb), fb = acc_full(rb + 2);
        const uint32_t tcol0 = tmem_base + lane_base + (uint32_t)(half * kHalfCols);
        const uint32_t ta = tcol0 + rb * kTileN, tb = tcol0 + (rb + 2) * kTileN;
        uint32_t pa = 0, pb = 0;
        int cur = 0;   // 0: the next use takes stage A, 1: stage B
        int pad_tile;   // pinned in a register (otherwise re-derived from the kernel parameters every tile)
        asm volatile("mov.b32 %0, %1;" : "=r"(pad_tile) : "r"(padded ? tiles - 1 : -1));
        pdl_wait();     // the predecessor's rows / windows / segment headers are complete and visible from here on
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            const int64_t row = (int64_t)p.unit * kUnitRows + rb * kBlockRows + q * 32 + lane;
            const int sseg = p.seg * kHalves + half;
            if (half == 0) {
                // Row operand into shared memory.  Every MMA of the previous piece that read this row
                // block's operand has completed: this warp consumed that piece's last accumulator.
                // The tensor core reads through the async proxy: fence before signalling.
                store_a_row_smem<KPAD>(smem_base + L::kAOff + rb * L::kABytes, q * 32 + lane, a_img + row * KPAD);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_full(rb));
            }

            const float w = wrow[row];
            float runmax = -CUDART_INF_F;
            float thr = (w >= 0.f) ? -CUDART_INF_F : CUDART_INF_F;
#ifdef GQ_EXP_NOREC
            thr = CUDART_INF_F;   // timing experiment only: the record path is never taken (results are wrong)
#endif
            [[maybe_unused]] uint32_t thr2 = pack_thr(thr);   // fp16 accumulators: (thr, thr) rounded down
            uint32_t cnt = 0;
            const uint32_t seg_index = (uint32_t)(row * nseg + sseg);   // < 2^32, checked on the host
            int t = p.t0;

            // One accumulator use: wait, pull the warp's 32x64 half tile into registers, release the
            // TMEM stage at once (tcgen05.wait::ld is warp-collective, so lane 0 may arrive for the
            // warp), then reduce two 32-column chunks against the running maximum.  `nbar/npar` name
            // the barrier of the use after this one (the row block's other stage).
            auto use_tile = [&](const uint32_t fbar, const uint32_t taddr, const uint32_t par, const uint32_t nbar,
                                const uint32_t npar, const bool pad) {
                GQ_TRACE(0);
                mbar_probe_wait_a(fbar, par);
                GQ_TRACE(1);
                tc_fence_after();
                if constexpr (ACC16) {
                    static_assert(!ACC16 || (kHalfCols == 64 && kChunk == 32), "fp16-accumulator epilogue: 64-column half tiles, 32-codeword chunks");
                    // 64 fp16 scores as 32 packed registers: register i = (codeword 64*half + i, codeword 64*half + 32 + i)
                    // of this tile, i.e. (chunk 2*half, chunk 2*half + 1) -- the codebook image is stored in that order
                    uint32_t pv[32];
                    tmem_ld_32x64p_sync(taddr, pv);
                    GQ_TRACE(2);
                    tc_fence_before();
                    if (lane == 0) mbar_arrive_a(fbar + kEmptyOff);
                    if constexpr (DUMP) {
                        float *drow = dump + row * ((int64_t)tiles * kTileN) + (int64_t)t * kTileN + half * kHalfCols;
#pragma unroll
                        for (int i = 0; i < 32; ++i) { drow[i] = half_lo(pv[i]); drow[32 + i] = half_hi(pv[i]); }
                    }
                    if (pad) {
                        // columns >= N hold the zero rows of the operand image: replace them with -inf
                        const int n_in_half = (int)(N - (int64_t)t * kTileN) - half * kHalfCols;
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (i >= n_in_half) pv[i] = (pv[i] & 0xffff0000u) | 0xfc00u;
                            if (32 + i >= n_in_half) pv[i] = (pv[i] & 0x0000ffffu) | 0xfc000000u;
                        }
                    }
                    mbar_probe_a(nbar, npar);
                    const uint32_t m2 = chunk_max_packed(pv);
                    if (hge2_any(m2, thr2)) {
                        const float m0 = half_lo(m2), m1 = half_hi(m2);
                        const uint32_t c0 = (uint32_t)t * (kTileN / kChunk) + (uint32_t)(half * 2);
                        if (m0 >= thr) epi_record_t<ACC16>(m0, c0, w, runmax, thr, cnt, cand + (int64_t)seg_index * kCap);
                        if (m1 >= thr) epi_record_t<ACC16>(m1, c0 + 1u, w, runmax, thr, cnt, cand + (int64_t)seg_index * kCap);
                        thr2 = pack_thr(thr);
                    }
                } else {
                float v[kHalfCols];
                tmem_ld_cols_sync(taddr, v);
                GQ_TRACE(2);
                tc_fence_before();
                if (lane == 0) mbar_arrive_a(fbar + kEmptyOff);
                if constexpr (DUMP) {
                    // debug build only: spill the raw tensor-core scores [M_pad, tiles*128]
                    float *drow = dump + row * ((int64_t)tiles * kTileN) + (int64_t)t * kTileN + half * kHalfCols;
#pragma unroll
                    for (int i = 0; i < kHalfCols; ++i) drow[i] = v[i];
                }
                if (pad) {
                    // last tile of a codebook that is not a multiple of 128: columns >= N hold the
                    // zero padding rows of the operand image and must not raise the running max
                    const int n_in_half = (int)(N - (int64_t)t * kTileN) - half * kHalfCols;
                    mbar_probe_a(nbar, npar);
#pragma unroll
                    for (int c = 0; c < kHalfCols / kChunk; ++c) {
                        const float m = chunk_max32_masked(v + c * kChunk, n_in_half - c * kChunk);
                        if (m >= thr && c * kChunk < n_in_half) epi_record_t<ACC16>(m, (uint32_t)t * (kTileN / kChunk) + (uint32_t)(half * (kHalfCols / kChunk) + c), w, runmax, thr, cnt, cand + (int64_t)seg_index * kCap);
                    }
                } else {
                    // Probe the NEXT use's barrier half way through the reduction: the answer (a predicate
                    // only mbar_probe_wait_a reads) is there by the time the next use starts, so a
                    // satisfied wait costs nothing instead of ~100 idle cycles per tile.  Once true it
                    // stays true: the phase cannot advance again before this warp drains the stage.
                    constexpr int kC = kHalfCols / kChunk;
                    float m[kC];
#pragma unroll
                    for (int c = 0; c < kC; ++c) {
                        m[c] = chunk_max32(v + c * kChunk);
                        if (c == kC / 2 - 1) mbar_probe_a(nbar, npar);
                    }
                    float mm = m[0];
#pragma unroll
                    for (int c = 1; c < kC; ++c) mm = fmaxf(mm, m[c]);
                    if (mm >= thr) {
#pragma unroll
                        for (int c = 0; c < kC; ++c)
                            if (m[c] >= thr) epi_record_t<ACC16>(m[c], (uint32_t)t * (kTileN / kChunk) + (uint32_t)(half * (kHalfCols / kChunk) + c), w, runmax, thr, cnt, cand + (int64_t)seg_index * kCap);
                    }
                }
                }   // fp32 accumulators
                ++t;
                GQ_TRACE(3);
                GQ_TRACE_NEXT();
            };
            // The row block's uses alternate between its stages A = rb and B = rb + 2 (`cur` says which
            // comes next and carries over from piece to piece).  The loop is unrolled over one A/B pair:
            // no per-tile stage bookkeeping, half the loop branches (taken branches cost ~20 cycles of
            // latency each, and this loop's serial latency is what bounds the kernel).  A padded last
            // tile (pad_tile = tiles - 1, so it can only end a piece) is peeled off.
            const bool has_pad = pad_tile >= p.t0 && pad_tile < p.t1;
            const int t_fast = has_pad ? p.t1 - 1 : p.t1;
            // (a use of stage X flips X's parity; the "next" parity handed to the probe is the other
            // stage's current one.  Over two A/B pairs both parities return to where they started, so
            // the main loop carries no parity updates at all.)
            if (cur && t < t_fast) { use_tile(fb, tb, pb, fa, pa, false); pb ^= 1u; cur = 0; }
            while (t + 3 < t_fast) {   // two A/B pairs per trip: measured 3 % faster than one
                use_tile(fa, ta, pa, fb, pb, false);
                use_tile(fb, tb, pb, fa, pa ^ 1u, false);
                use_tile(fa, ta, pa ^ 1u, fb, pb ^ 1u, false);
                use_tile(fb, tb, pb ^ 1u, fa, pa, false);
            }
            while (t + 1 < t_fast) {
                use_tile(fa, ta, pa, fb, pb, false);
                pa ^= 1u;
                use_tile(fb, tb, pb, fa, pa, false);
                pb ^= 1u;
            }
            if (t < t_fast) { use_tile(fa, ta, pa, fb, pb, false); pa ^= 1u; cur = 1; }
            if (has_pad) {
                if (cur) { use_tile(fb, tb, pb, fa, pa, true); pb ^= 1u; cur = 0; }
                else { use_tile(fa, ta, pa, fb, pb, true); pa ^= 1u; cur = 1; }
            }
            Cand *list = cand + (int64_t)seg_index * kCap;
            // piece done: filter this thread's own records against its final maximum and publish the
            // survivors (typically one or two) in the 64-byte segment summary
            {
                SegHdr *hp = &seg_hdr[seg_index];
                uint32_t ns = 0;
                bool ovf = cnt > (uint32_t)kCap;
                const float keep = win_thr(runmax, w, ACC16 ? kWinRel16 : 0.f);
                const uint32_t nrec = cnt < (uint32_t)kCap ? cnt : (uint32_t)kCap;
                for (uint32_t i = 0; i < nrec; ++i) {
                    const Cand c = list[i];
                    if (c.score >= keep) {
                        if (ns < (uint32_t)kSurv) hp->surv[ns] = c;
                        else ovf = true;
                        ++ns;
                    }
                }
                hp->runmax = runmax;
                hp->cnt = ovf ? kSegOverflow : ns;
                rec_cnt[seg_index] = cnt;
            }
        }
        }   // column-half epilogue
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- 2b. the search with fp16 accumulators and the row operand in TENSOR MEMORY ------------------------------
// What bounds the kernel above is neither arithmetic nor latency any more but the operand fetch of
// tcgen05.mma: scripts/ubench/mma_rate.cu measures a kind::f16 instruction (K = 16) with both operands in
// shared memory at 114 cycles for M = N = 128 -- 56 % of the tensor rate, and exactly the 4 x 114 = 457 cycles per
// (256-row, 128-codeword) step the SS kernel runs at -- whatever the shared-memory layout; shared memory moves
// ~90 B/cycle there (A 4 KB + B 4 KB per instruction, plus the tile copies).  With the row operand in tensor
// memory (TS form) the same instruction takes 87 cycles and reads half the bytes.  The price is tensor memory:
// 2 x 16 columns of row operand leave room for THREE 128-column accumulator stages, used round-robin by the
// (tile, row block) sequence -- MMA-tile g = 2 c + rb of the CTA's running tile count c goes to stage g % 3 -- instead
// of two private ones per row block.  fp16 accumulators keep the epilogue (packed loads, packed 3-input maxima)
// short enough for that: a stage's round trip (MMA pair 175 cycles, wake-up, tcgen05.ld, release, issuer wake-up,
// issue) is ~440 cycles, three stages deliver a tile every ~150.
// Warps: 16 epilogue (row block, column half, lane quarter), producer, one MMA issuer per row block.
template <int KPAD>
struct Tc16Smem {
    static constexpr int kBTileBytes = kTileN * KPAD * 2;
    static constexpr int kFit = (200 * 1024) / kBTileBytes;
    static constexpr int kStages = (kFit < 16 ? kFit : 16) & ~1;
    static constexpr int kBOff = 0;
    static constexpr int kBarOff = kBOff + kStages * kBTileBytes;
    // barriers: full_b[kStages], empty_b[kStages], a_full[2], acc_full[3][2], acc_empty[3][2]
    static constexpr int kNumBars = 2 * kStages + 2 + 12;
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kBytes = kSlotOff + 16;
};
constexpr int kAcc16Stages = 3;
constexpr int kTmemACol = kAcc16Stages * kTileN;      // row operands: columns 384 + rb * KPAD / 2
constexpr int kProducerWarp16 = kNumEpiWarps, kIssuerWarp16 = kNumEpiWarps + 1;
constexpr int kThreads16 = (kNumEpiWarps + 3) * 32;

// One row of the row-operand image (KPAD fp16) into the row's TMEM lane: two K-elements per 32-bit column.
template <int KPAD>
__device__ __forceinline__ void store_a_row_tmem(uint32_t taddr, const __half *__restrict__ a_row)
{
    constexpr int NW = KPAD / 2;
    static_assert(NW == 8 || NW == 16, "K = 16 or 32");
    uint32_t r[NW];
#pragma unroll
    for (int c = 0; c < NW / 4; ++c) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(a_row) + c);
        r[4 * c + 0] = q.x; r[4 * c + 1] = q.y; r[4 * c + 2] = q.z; r[4 * c + 3] = q.w;
    }
    if constexpr (NW == 8) tmem_st_32x8(taddr, r);
    else tmem_st_32x16(taddr, r);
    tmem_st_wait();
}

template <int KPAD, bool DUMP>
__global__ void __launch_bounds__(kThreads16, 1)
quantize_tc16_kernel(const __half *__restrict__ a_img, const __half *__restrict__ b_img,
                     const float *__restrict__ wrow, SegHdr *__restrict__ seg_hdr, Cand *__restrict__ cand,
                     uint32_t *__restrict__ rec_cnt, int64_t N, int tiles, int64_t total_steps, int steps_per_cta,
                     int nseg, float *__restrict__ dump)
{
    using L = Tc16Smem<KPAD>;
    constexpr int kStages = L::kStages;
    static_assert(kHalves == 2 && kChunk == 32, "column-half epilogue, 32-codeword chunks");
    extern __shared__ __align__(128) uint8_t smem[];
    uint32_t smem_base;
    asm volatile("mov.u32 %0, %1;" : "=r"(smem_base) : "r"(smem_u32(smem)));
    const uint32_t bar_base = smem_base + L::kBarOff;
    auto full_b = [&](int i) { return bar_base + 8u * i; };
    auto empty_b = [&](int i) { return bar_base + 8u * (kStages + i); };
    auto a_full = [&](int i) { return bar_base + 8u * (2 * kStages + i); };                         // per row block
    // A stage is used by the two row blocks ALTERNATELY (MMA-tiles g, g + 3, g + 6, ...), and the two issuers run
    // independently, so a parity wait on a per-stage barrier would alias: a row block waiting for its use g + 6 cannot
    // tell "g + 3 has not completed yet" from "g + 6 has".  Every (stage, row block) therefore has its own pair:
    // acc_full(s, rb) is committed by rb's issuer and waited for by rb's warps (consecutive uses of s BY rb, g / 6 = 0, 1, ...);
    // acc_empty(s, rb) collects the arrivals of rb's eight warps and is waited for by the OTHER row block's issuer,
    // the stage's next user.  Each waiter sees consecutive phases of its barrier.
    auto acc_full = [&](int st, int rb) { return bar_base + 8u * (2 * kStages + 2 + 2 * st + rb); };
    constexpr uint32_t kEmptyOff = 8u * 2 * kAcc16Stages;                                           // acc_empty(s, rb) = acc_full(s, rb) + kEmptyOff
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L::kSlotOff);

    // (broadcast from lane 0: tells the compiler the warp index is warp-uniform, so that the service warps'
    // counters, barrier and descriptor addresses live in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int cta = blockIdx.x;
    const int64_t step_begin = (int64_t)cta * steps_per_cta;
    const int64_t step_end = imin(step_begin + steps_per_cta, total_steps);
    mbar_probe_declare();
    pdl_launch_dependents();
    GQ_TRACE_DECL();

    if (warp == kIssuerWarp16 && lane == 0) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::kBarOff);
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[i], 1);                                  // full_b
        for (int i = 0; i < kStages; ++i) mbar_init(&bars[kStages + i], 2);                        // empty_b: one commit per row block
        for (int i = 0; i < 2; ++i) mbar_init(&bars[2 * kStages + i], 4);                          // a_full: 4 warps
        for (int i = 0; i < 2 * kAcc16Stages; ++i) mbar_init(&bars[2 * kStages + 2 + i], 1);           // acc_full(s, rb)
        for (int i = 0; i < 2 * kAcc16Stages; ++i) mbar_init(&bars[2 * kStages + 2 + 2 * kAcc16Stages + i], 4 * kHalves);   // acc_empty(s, rb): 8 warps
        fence_mbar_init();
    }
    if (warp == kIssuerWarp16) {
        __syncwarp();
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kProducerWarp16) {
        // ===== producer: bulk async copies (TMA engine) of codebook tiles =====
        int64_t step = step_begin;
        Piece p;
        int stage = 0;
        uint32_t phase = 0;
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            for (int t = p.t0; t < p.t1; ++t) {
                GQ_TRACE(0);
                mbar_wait_a(empty_b(stage), phase ^ 1);
                GQ_TRACE(1);
                if (elect_one()) {
                    mbar_expect_tx_a(full_b(stage), L::kBTileBytes);
                    bulk_g2s_a(smem_base + L::kBOff + stage * L::kBTileBytes,
                               reinterpret_cast<const char *>(b_img) + (int64_t)t * L::kBTileBytes, L::kBTileBytes, full_b(stage));
                }
                __syncwarp();
                GQ_TRACE(2);
                GQ_TRACE(3);
                GQ_TRACE_NEXT();
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kIssuerWarp16 || warp == kIssuerWarp16 + 1) {
        // ===== MMA issuers, one per row block.  The whole warp runs the loop (warp-uniform control flow keeps the
        // addresses on the uniform datapath), one elected lane issues to the tensor core. =====
        const int rb = warp - kIssuerWarp16;
        constexpr uint32_t idesc = umma_idesc_f16(kBlockRows, kTileN, /*d_f32=*/false);
        constexpr uint32_t kLbo = 128;              // K-adjacent core matrices
        constexpr uint32_t kSbo = (KPAD / 8) * 128; // next 8-row group
        constexpr uint32_t desc_hi = ((kSbo >> 4) & 0x3FFFu) | (1u << 14);           // SBO, version = 1
        const uint32_t desc_lo0 = (((smem_base + L::kBOff) >> 4) & 0x3FFFu) | (((kLbo >> 4) & 0x3FFFu) << 16);
        const uint32_t a_t = tmem_base + (uint32_t)(kTmemACol + rb * (KPAD / 2));
        int64_t step = step_begin;
        Piece p;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t g = (uint32_t)rb;   // MMA-tile number of this row block's next tile: g = 2 c + rb; stage g % 3
        int pi = 0;
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            mbar_wait_a(a_full(rb), pi & 1);
            tc_fence_after();
            for (int t = p.t0; t < p.t1; ++t) {
                const int as = (int)(g % 3u);
                GQ_TRACE(0);
                // the stage's previous user was the other row block, MMA-tile g - 3, its use number (g - 3) / 6 of the stage
                // (none for g < 3: parity 1 passes on a fresh barrier)
                mbar_wait2_a(full_b(stage), phase, acc_full(as, rb ^ 1) + kEmptyOff, ((g + 3u) / 6u + 1u) & 1u);
                GQ_TRACE(1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d = tmem_base + (uint32_t)(as * kTileN);
                    const uint32_t desc_lo = desc_lo0 + (uint32_t)stage * (L::kBTileBytes >> 4);
#pragma unroll
                    for (int j = 0; j < KPAD / 16; ++j) {
                        // one K = 16 slice: 8 columns of the row operand, two K-adjacent core matrices (256 bytes) of the tile
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)(desc_lo + j * 16);
                        umma_f16_ts(d, a_t + (uint32_t)(j * 8), bd, idesc, j > 0 ? 1u : 0u);
                    }
                    GQ_TRACE(2);
                    umma_commit_a(acc_full(as, rb));
                    umma_commit_a(empty_b(stage));   // 2 arrivals (one per row block) free the slot
                }
                __syncwarp();
                GQ_TRACE(3);
                GQ_TRACE_NEXT();
                g += 2u;
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            ++pi;
        }
    } else if (warp < kNumEpiWarps) {
        // ===== epilogue: (row block, column half, lane quarter).  Thread <-> (row, column half): two threads per row
        // with independent running maxima / record lists, storage segment 2 * piece + half. =====
        const int rb = (warp >> 2) & 1;
        const int half = warp >> 3;
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const uint32_t tcol0 = tmem_base + lane_base + (uint32_t)(half * 64);
        const bool padded = (N % kTileN) != 0;
        int pad_tile;
        asm volatile("mov.b32 %0, %1;" : "=r"(pad_tile) : "r"(padded ? tiles - 1 : -1));
        int64_t step = step_begin;
        Piece p;
        // stage / barrier parity of the next tile to LOAD: MMA-tile g = 2 c + rb, stage g % 3, use g / 6 of acc_full(stage, rb)
        int ld_as = rb, ld_g6 = rb;
        uint32_t ld_par = 0;
        auto ld_advance = [&]() {
            ld_as += 2; if (ld_as >= kAcc16Stages) ld_as -= kAcc16Stages;
            ld_g6 += 2; if (ld_g6 >= 6) { ld_g6 -= 6; ld_par ^= 1u; }
        };
        pdl_wait();           // the predecessor's rows / windows / segment headers are complete and visible from here on
        while (next_piece(step, step_end, tiles, steps_per_cta, cta, p)) {
            const int64_t row = (int64_t)p.unit * kUnitRows + rb * kBlockRows + q * 32 + lane;
            const int sseg = p.seg * kHalves + half;
            if (half == 0) {
                // Row operand into tensor memory.  Every MMA of the previous piece that read this row block's operand
                // has completed: this warp consumed that piece's last accumulator.
                store_a_row_tmem<KPAD>(tmem_base + lane_base + (uint32_t)(kTmemACol + rb * (KPAD / 2)), a_img + row * KPAD);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(a_full(rb));
            }
            const float w = wrow[row];
            float runmax = -CUDART_INF_F;
            float thr = (w >= 0.f) ? -CUDART_INF_F : CUDART_INF_F;
#ifdef GQ_EXP_NOREC
            thr = CUDART_INF_F;   // timing experiment only: the record path is never taken (results are wrong)
#endif
            uint32_t thr2 = pack_thr(thr);
            uint32_t cnt = 0;
            const uint32_t seg_index = (uint32_t)(row * nseg + sseg);   // < 2^32, checked on the host
            Cand *list = cand + (int64_t)seg_index * kCap;

            // Software pipeline over two register buffers: the NEXT tile's tcgen05.ld is issued before the current
            // tile is reduced (if its accumulator is complete by then -- asked with an early test_wait probe whose
            // answer is read only after the current buffer has been waited for and its stage released), so the load
            // latency and the barrier round trip hide behind the ~250-cycle reduction instead of adding to it.
            // 64 fp16 scores as 32 packed registers: register i = (codeword 64*half + i, codeword 64*half + 32 + i)
            // of the tile, i.e. (chunk 2*half, chunk 2*half + 1) -- the codebook image is stored in that order.
            auto issue_load = [&](uint32_t (&buf)[32], uint32_t &fbar_out) {
                fbar_out = acc_full(ld_as, rb);
                tc_fence_after();
                tmem_ld_32x64p(tcol0 + (uint32_t)(ld_as * kTileN), buf);
                ld_advance();
            };
            auto process = [&](uint32_t (&pv)[32], int t) {
                if constexpr (DUMP) {
                    float *drow = dump + row * ((int64_t)tiles * kTileN) + (int64_t)t * kTileN + half * 64;
#pragma unroll
                    for (int i = 0; i < 32; ++i) { drow[i] = half_lo(pv[i]); drow[32 + i] = half_hi(pv[i]); }
                }
                if (t == pad_tile) {
                    // columns >= N hold the zero rows of the operand image: replace them with -inf
                    const int n_in_half = (int)(N - (int64_t)t * kTileN) - half * 64;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        if (i >= n_in_half) pv[i] = (pv[i] & 0xffff0000u) | 0xfc00u;
                        if (32 + i >= n_in_half) pv[i] = (pv[i] & 0x0000ffffu) | 0xfc000000u;
                    }
                }
                const uint32_t m2 = chunk_max_packed(pv);
                if (hge2_any(m2, thr2)) {
                    const float m0 = half_lo(m2), m1 = half_hi(m2);
                    const uint32_t c0 = (uint32_t)t * (kTileN / kChunk) + (uint32_t)(half * 2);
                    if (m0 >= thr) epi_record_t<true>(m0, c0, w, runmax, thr, cnt, list);
                    if (m1 >= thr) epi_record_t<true>(m1, c0 + 1u, w, runmax, thr, cnt, list);
                    thr2 = pack_thr(thr);
                }
            };
            // one pipeline step: `cur` holds tile t (its load is in flight), `nxt` receives tile t + 1 if there is one
            auto pipe_step = [&](uint32_t (&cur)[32], uint32_t fbar_cur, uint32_t (&nxt)[32], uint32_t &fbar_nxt, int t) {
                const bool more = t + 1 < p.t1;
                GQ_TRACE(0);
                if (more) mbar_probe_a(acc_full(ld_as, rb), ld_par);     // asynchronous: answered while we wait for `cur`
                tmem_ld_wait_p(cur);
                GQ_TRACE(1);
                tc_fence_before();
                if (lane == 0) mbar_arrive_a(fbar_cur + kEmptyOff);
                bool issued = false;
                if (more && mbar_probe_result()) { issue_load(nxt, fbar_nxt); issued = true; }
                GQ_TRACE(2);
                process(cur, t);
                if (more && !issued) { mbar_wait_a(acc_full(ld_as, rb), ld_par); issue_load(nxt, fbar_nxt); }
                GQ_TRACE(3);
                GQ_TRACE_NEXT();
            };
            uint32_t bufA[32], bufB[32];
            uint32_t fbarA = 0, fbarB = 0;
            mbar_wait_a(acc_full(ld_as, rb), ld_par);
            issue_load(bufA, fbarA);
            for (int t = p.t0; t < p.t1; t += 2) {
                pipe_step(bufA, fbarA, bufB, fbarB, t);
                if (t + 1 < p.t1) pipe_step(bufB, fbarB, bufA, fbarA, t + 1);
            }
            // piece done: filter this thread's own records against its final maximum and publish the survivors
            {
                SegHdr *hp = &seg_hdr[seg_index];
                uint32_t ns = 0;
                bool ovf = cnt > (uint32_t)kCap;
                const float keep = win_thr(runmax, w, kWinRel16);
                const uint32_t nrec = cnt < (uint32_t)kCap ? cnt : (uint32_t)kCap;
                for (uint32_t i = 0; i < nrec; ++i) {
                    const Cand cd = list[i];
                    if (cd.score >= keep) {
                        if (ns < (uint32_t)kSurv) hp->surv[ns] = cd;
                        else ovf = true;
                        ++ns;
                    }
                }
                hp->runmax = runmax;
                hp->cnt = ovf ? kSegOverflow : ns;
                rec_cnt[seg_index] = cnt;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kIssuerWarp16) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// The search kernel of an operand image and its launch geometry.
template <int G, int KPAD, bool SPLIT, bool DUMP>
struct SearchKernel {
    static constexpr bool kWide = gq_acc16_ts(G, SPLIT);   // fp16 accumulators + row operand in tensor memory
    static constexpr int kSmem = kWide ? Tc16Smem<KPAD>::kBytes : TcSmem<KPAD>::kBytes;
    static constexpr int kBlock = kWide ? kThreads16 : tc_threads(gq_acc16(G, SPLIT));
    static auto fn()
    {
        if constexpr (kWide) return quantize_tc16_kernel<KPAD, DUMP>;
        else return quantize_tc_kernel<KPAD, DUMP, gq_acc16(G, SPLIT)>;
    }
};

// ---- 3. re-scoring of the recorded chunks: one warp per row, lane = codeword ----------------------
constexpr int kRescoreWarps = 8;

// fp32 FMA form of the score (up to a per-row constant), the rescoring prefilter.  The row's parameters
// are read from shared memory (16-byte broadcast loads) instead of 2 G registers, and -- for G % 4 == 0 --
// the arithmetic runs on packed fp32 pairs (FADD2 / FMUL2 / FFMA2: half the issue slots;
// this kernel is issue bound).  `nm` holds -mu and `nri` holds -1/(2 var), so that x + nm = x - mu and
// fma(d*d, nri, a) = fma(-d*d, ri, a) exactly.  Even and odd dimensions accumulate separately and are
// added at the end: <= 6 roundings per term and <= G + 1 per summation path, i.e. <= 1.4 * 2^-20 * R, inside
// the 3 * 2^-20 * R that prep_row budgets; the value is only ever compared with other values of this function.
template <int G>
__device__ __forceinline__ float prefilter_score_smem(const float (&x)[G], const float *__restrict__ nm,
                                                      const float *__restrict__ nri, float hb)
{
    if constexpr (G % 4 == 0) {
        float2 acc = make_float2(0.f, 0.f);
        const float2 hb2 = make_float2(hb, hb);
#pragma unroll
        for (int k4 = 0; k4 < G / 4; ++k4) {
            const float4 mm = reinterpret_cast<const float4 *>(nm)[k4];
            const float4 rr = reinterpret_cast<const float4 *>(nri)[k4];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 xv = make_float2(x[4 * k4 + 2 * j], x[4 * k4 + 2 * j + 1]);
                const float2 mv = j ? make_float2(mm.z, mm.w) : make_float2(mm.x, mm.y);
                const float2 rv = j ? make_float2(rr.z, rr.w) : make_float2(rr.x, rr.y);
                const float2 d = __fadd2_rn(xv, mv);
                acc = __ffma2_rn(__fmul2_rn(d, d), rv, acc);
                acc = __ffma2_rn(__fmul2_rn(hb2, xv), xv, acc);
            }
        }
        return acc.x + acc.y;
    } else {
        float a2 = 0.f;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float d = x[k] + nm[k];
            a2 = fmaf(d * d, nri[k], a2);
            a2 = fmaf(hb * x[k], x[k], a2);
        }
        return a2;
    }
}

constexpr int kMaxChunks = 64;   // surviving chunks one row may carry into the rescoring (else: exact scan)

// ---- pruned exact scan of a row the search could not certify ----------------------------------------------------
// Such rows are the sharp ones (a variance of ~1e-5 or less in some dimension: the reference's extension-test
// latents, std = |randn|): their fp16 / split-precision windows are wider than their score gaps.  But a sharp
// dimension k* prunes by itself: t_k*(x) = -(x - mu)^2 / (2 v) + beta x^2 / 2 + const falls off so fast that only
// codewords with x_k* within a few sigma of mu_k* can win.  The codebook is kept sorted per dimension
// (gq_codebook::sort_val / sort_idx); one warp walks outward from mu_k* in that order, 16 candidates per side
// and step, scoring each in the reference's exact op order (lane = candidate; lowest index wins ties), and stops a
// side as soon as an upper bound on every unvisited candidate's score there -- the k* term at the next distance
// plus, for every other dimension, the maximum of its term over the codebook's range of that coordinate, evaluated
// in double, with slack for the fp32 rounding of the exact evaluation -- falls strictly below the best exact score
// seen.  Rigorous, so the result is the reference argmax; rows that do not prune within kPrunedBudget candidates (or
// hold non-finite values) return false and take the sliced full scan.
constexpr int kPrunedBudget = 8192;

template <int G>
__device__ __noinline__ bool pruned_exact_row(int64_t r, const float *__restrict__ mu, const float *__restrict__ std,
                                 const float *__restrict__ lstd, const float *__restrict__ cb, const float *__restrict__ nlpb,
                                 const float *__restrict__ meta, const float *__restrict__ sort_val,
                                 const uint32_t *__restrict__ sort_idx, int64_t N, float beta, int lane,
                                 int64_t *__restrict__ idx_out, float *__restrict__ zhat_out)
{
    float m[G], den[G], ls[G];
    int ks = 0;
    float smin = CUDART_INF_F;
    double crest = 0.0, cabs = 0.0, c_ks = 0.0, inv2v_ks = 0.0;
    {
        float sd[G];
        load_row<G>(mu + r * G, m);
        load_row<G>(std + r * G, sd);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            den[k] = __fmul_rn(2.0f, __fmul_rn(sd[k], sd[k]));
            ls[k] = lstd ? __ldg(lstd + r * G + k) : logf(sd[k]);
            if (sd[k] < smin) { smin = sd[k]; ks = k; }
        }
        const double c0 = (double)GQ_C0F;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const double v = (double)sd[k] * (double)sd[k];
            const double cst = -(double)ls[k] - c0 + (double)beta * c0;
            if (k == ks) {
                inv2v_ks = 0.5 / v;
                c_ks = 0.5 * (double)beta * (double)meta[2 + k] + cst;       // beta x^2 / 2 <= beta max x_k^2 / 2
                cabs += fabs(c_ks);
            } else {
                // max over |x| <= X of  a x^2 + b x - mu^2 / (2 v),  a = beta/2 - 1/(2v),  b = mu / v
                const double a = 0.5 * (double)beta - 0.5 / v, bb = (double)m[k] / v, X = (double)meta[2 + G + k];
                double fmx;
                if (a < 0.0) {
                    double xs = -bb / (2.0 * a);
                    xs = xs > X ? X : (xs < -X ? -X : xs);
                    fmx = a * xs * xs + bb * xs;
                } else {
                    fmx = a * X * X + fabs(bb) * X;
                }
                const double term = fmx - 0.5 * (double)m[k] * (double)m[k] / v + cst;
                crest += term;
                cabs += fabs(fmx) + 0.5 * (double)m[k] * (double)m[k] / v + fabs(cst);
            }
        }
    }
    // everything an unvisited candidate can score beyond its k* distance term, plus slack for the fp32 evaluation
    const double cbound = crest + c_ks + 1e-5 * cabs + 1e-3;
    const float *sv = sort_val + (int64_t)ks * N;
    const uint32_t *si = sort_idx + (int64_t)ks * N;
    const float mk = m[ks];
    int64_t lo = 0, hi = N;          // lower_bound(sv, mk)
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(sv + mid) < mk) lo = mid + 1; else hi = mid;
    }
    int64_t left = lo - 1, right = lo;   // next unvisited position on either side
    bool open_l = left >= 0, open_r = right < N;
    float best = -CUDART_INF_F;
    int best_idx = 0x7fffffff;
    int visited = 0;
    if (!(mk == mk) || !(cbound == cbound)) return false;      // NaN: the sliced scan decides
    while (open_l || open_r) {
        const bool is_l = lane < 16;
        const int64_t pos = is_l ? left - lane : right + (lane - 16);
        if ((is_l ? open_l : open_r) && pos >= 0 && pos < N) {
            const int64_t n = (int64_t)__ldg(si + pos);
            float x[G], nb[G];
            load_row<G>(cb + n * G, x);
            load_row<G>(nlpb + n * G, nb);
            const float sc = score_exact<G>(x, nb, m, den, ls);
            if (sc > best || (sc == best && (int)n < best_idx)) { best = sc; best_idx = (int)n; }
        }
        if (open_l) left -= 16;
        if (open_r) right += 16;
        visited += 32;
        float gb = best;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gb = fmaxf(gb, __shfl_xor_sync(0xffffffffu, gb, off));
        if (open_l) {
            if (left < 0) open_l = false;
            else {
                const double d = (double)mk - (double)__ldg(sv + left);
                if (-(d * d) * inv2v_ks * (1.0 - 1.0 / 262144.0) + cbound < (double)gb) open_l = false;
            }
        }
        if (open_r) {
            if (right >= N) open_r = false;
            else {
                const double d = (double)__ldg(sv + right) - (double)mk;
                if (-(d * d) * inv2v_ks * (1.0 - 1.0 / 262144.0) + cbound < (double)gb) open_r = false;
            }
        }
        if (visited >= kPrunedBudget && (open_l || open_r)) return false;
        // Is this row prunable at all?  With the best score seen so far, every codeword whose k* coordinate lies within
        // d* of mu_k* -- d*^2 / (2 v) = cbound - best -- would still have to be visited.  For a row WITHOUT a sharp
        // dimension (a trained-like row that merely missed its certificate) that is most of the codebook: one warp
        // would walk its whole budget for nothing (measured: 615 us for a single such row in a 65 536-row batch, 2 x the
        // whole step) before the sliced scan, which uses every SM, gets the row anyway.  Asked twice (the best score
        // improves as the walk goes on); a wrong "no" only costs speed, the sliced scan is exact too.
        if ((visited == 32 || visited == 1024) && (open_l || open_r)) {
            const double gap = cbound - (double)gb;
            if (!(gap <= 0.0)) {
                const double dneed = sqrt(gap / (inv2v_ks * (1.0 - 1.0 / 262144.0)));
                const float x_lo = (float)((double)mk - dneed), x_hi = (float)((double)mk + dneed);
                int64_t a = 0, b = N;
                while (a < b) { const int64_t mid = (a + b) >> 1; if (__ldg(sv + mid) < x_lo) a = mid + 1; else b = mid; }
                const int64_t first = a;
                b = N;
                while (a < b) { const int64_t mid = (a + b) >> 1; if (__ldg(sv + mid) <= x_hi) a = mid + 1; else b = mid; }
                if (!(dneed == dneed) || (a - first) - (int64_t)visited > 4 * (int64_t)kPrunedBudget) return false;
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
        if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    }
    if (best_idx == 0x7fffffff || !(best > -CUDART_INF_F)) return false;
    if (lane == 0) idx_out[r] = (int64_t)best_idx;
    if (zhat_out && lane < G) zhat_out[r * G + lane] = __ldg(cb + (int64_t)best_idx * G + lane);
    return true;
}

// Rows whose fp32 prefilter left more than one contender (a few percent): one warp per listed row
// re-reads the segment summaries and scores every codeword of the surviving chunks (lane = codeword)
// in the reference's exact fp32 op order; first-max tie-break.  Runs after rescore_kernel, before the
// exact-scan fallback (to which it can still defer a row with non-finite scores).
// One row whose fp32 prefilter left more than one contender: score every codeword of its surviving chunks (lane =
// codeword) in the reference's exact fp32 op order, first-max tie-break, and write the result.  `my_word`: lane 0 holds
// the number of listed chunks (<= 7, else the summaries / full record lists are walked), lanes 1..7 the chunk ids.
// Returns false when no finite score was found (the caller sends the row on).
template <int G>
__device__ __forceinline__ bool settle_row(int64_t r, bool full, uint32_t my_word, int lane,
                                           const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
                                           const float *__restrict__ cb, const float *__restrict__ nlpb, const float *__restrict__ wrow,
                                           const SegHdr *__restrict__ seg_hdr, const Cand *__restrict__ cand,
                                           const uint32_t *__restrict__ rec_cnt, int64_t N, int tiles, int steps_per_cta, int nseg,
                                           int unit_rows, float win_rel, int64_t *__restrict__ idx_out, float *__restrict__ zhat_out)
{
    const uint32_t n_listed = __shfl_sync(0xffffffffu, my_word, 0);
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
    float best = -CUDART_INF_F;
    int best_idx = 0x7fffffff;
    auto score_chunk = [&](uint32_t chunk) {
        const int64_t n = (int64_t)chunk * kChunk + lane;
        if (lane < kChunk && n < N) {
            float x[G], nb[G];
            load_row<G>(cb + n * G, x);
            load_row<G>(nlpb + n * G, nb);
            const float sc = score_exact<G>(x, nb, m, den, ls);
            if (sc > best || (sc == best && (int)n < best_idx)) { best = sc; best_idx = (int)n; }
        }
    };
    if (!full && n_listed <= 7u) {
        for (uint32_t c = 0; c < n_listed; ++c) score_chunk(__shfl_sync(0xffffffffu, my_word, (int)c + 1));
    } else {
        // Walk the summaries (or, for rows flagged `full`, the complete record lists) with the warp in
        // parallel: lane s takes segment s for the running maxima, lane i takes record i of a list
        // (kCap = 32 = one record per lane, one coalesced load per segment); the survivors' chunk ids
        // are then broadcast one by one.  A single lane walking the lists cost 15 us per launch.
        const uint32_t s0 = (uint32_t)(r / unit_rows) * (uint32_t)tiles;
        const int npieces = kHalves * ((int)((s0 + (uint32_t)tiles - 1u) / (uint32_t)steps_per_cta) - (int)(s0 / (uint32_t)steps_per_cta) + 1);
        float gmax = -CUDART_INF_F;
        for (int s = lane; s < npieces; s += 32) gmax = fmaxf(gmax, seg_hdr[r * nseg + s].runmax);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
        const float thr = win_thr(gmax, wrow[r], win_rel);
        static_assert(kCap == 32, "one record per lane");
        // eight segments per trip, their (independent) loads issued before any is consumed: with many
        // segments per row (small M: a unit is spread over up to 148 CTAs) this walk is the whole kernel
        constexpr int kU = 8;
        for (int sb = 0; sb < npieces; sb += kU) {
            Cand c[kU];
            uint32_t cnt[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int s = sb + u;
                c[u] = Cand{-CUDART_INF_F, 0u};
                cnt[u] = 0u;
                if (s < npieces) {
                    const SegHdr *h = &seg_hdr[r * nseg + s];
                    const Cand *list = full ? cand + (r * nseg + s) * (int64_t)kCap : h->surv;
                    cnt[u] = full ? rec_cnt[r * nseg + s] : h->cnt;   // <= kCap resp. <= kSurv by construction
                    if (lane < (full ? kCap : kSurv)) c[u] = list[lane];   // in bounds; masked by cnt below
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                unsigned mask = __ballot_sync(0xffffffffu, (uint32_t)lane < cnt[u] && c[u].score >= thr);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    score_chunk(__shfl_sync(0xffffffffu, c[u].idx, src));
                }
            }
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, off);
        const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
        if (ob > best || (ob == best && oi < best_idx)) { best = ob; best_idx = oi; }
    }
    if (best_idx == 0x7fffffff || !(best > -CUDART_INF_F)) return false;
    if (lane == 0) idx_out[r] = (int64_t)best_idx;
    if (zhat_out && lane < G) zhat_out[r * G + lane] = __ldg(cb + (int64_t)best_idx * G + lane);
    return true;
}

template <int G>
__global__ void __launch_bounds__(128)
settle_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
              const float *__restrict__ cb, const float *__restrict__ nlpb, const float *__restrict__ wrow,
              const SegHdr *__restrict__ seg_hdr, const Cand *__restrict__ cand, const uint32_t *__restrict__ rec_cnt,
              const uint32_t *__restrict__ rows, const uint32_t *__restrict__ row_chunks, const uint32_t *__restrict__ nrows, int64_t N, int tiles,
              int steps_per_cta, int nseg, int unit_rows, int64_t *__restrict__ idx_out, float *__restrict__ zhat_out,
              uint32_t *__restrict__ fb_rows, uint32_t *__restrict__ fb_count, unsigned long long *__restrict__ fb_keys,
              float win_rel, const uint32_t *__restrict__ pr_rows, const uint32_t *__restrict__ pr_count,
              const float *__restrict__ meta, const float *__restrict__ sort_val, const uint32_t *__restrict__ sort_idx, float beta)
{
    const int lane = threadIdx.x & 31;
    pdl_wait();
    // Second job of this kernel: the rows the rescoring kernel could not certify (pr_rows, complete before this
    // kernel starts) take the pruned exact scan; what that cannot decide goes on to the sliced full scan (fb_rows).
    const uint32_t npr = *pr_count;
    const uint32_t total = *nrows;      // (both counters are fetched together: one round trip at the head of the kernel)
    {
        for (uint32_t it = blockIdx.x * 4 + (threadIdx.x >> 5); it < npr; it += gridDim.x * 4) {
            const int64_t r = (int64_t)pr_rows[it];
            const bool done = pruned_exact_row<G>(r, mu, std, lstd, cb, nlpb, meta, sort_val, sort_idx, N, beta, lane, idx_out, zhat_out);
            if (!done && lane == 0) { fb_keys[r] = 0ull; idx_out[r] = -1; fb_rows[atomicAdd(fb_count, 1u)] = (uint32_t)r; }
        }
    }
    for (uint32_t it = blockIdx.x * 4 + (threadIdx.x >> 5); it < total; it += gridDim.x * 4) {
        const uint32_t entry = rows[it];
        const bool full = (entry & 0x80000000u) != 0u;   // a segment kept more than kSurv survivors: walk its full list
        const int64_t r = entry & 0x7fffffffu;
        // rescore_kernel attached the row's surviving chunks (lane i holds word i of the 8-word entry):
        // one load instead of a walk over the segment summaries
        const uint32_t my_word = lane < 8 ? __ldg(row_chunks + (int64_t)it * 8 + lane) : 0u;
        const bool done = settle_row<G>(r, full, my_word, lane, mu, std, lstd, cb, nlpb, wrow, seg_hdr, cand, rec_cnt, N, tiles,
                                        steps_per_cta, nseg, unit_rows, win_rel, idx_out, zhat_out);
        // (a row on the sliced scan's list carries idx = -1 until that kernel has decided it: see FinalLayout)
        if (!done && lane == 0) { fb_keys[r] = 0ull; idx_out[r] = -1; fb_rows[atomicAdd(fb_count, 1u)] = (uint32_t)r; }
    }
}

// What a rescoring kernel needs to FINISH a row itself (exact-order settlement, pruned exact scan) instead of
// handing it to settle_kernel: both are rare paths (under 1 % of the rows on trained latents), called out of line so
// that their register appetite (they spill under the 64-register cap) stays out of the main loop.
struct TailArgs {
    const float *nlpb;
    const Cand *cand;
    int tiles, steps_per_cta, unit_rows;
    const float *meta;
    const float *sort_val;
    const uint32_t *sort_idx;
    uint32_t *fb2_rows;      // rows left for the sliced full scan
    uint32_t *fb2_count;
};

template <int G>
__device__ __noinline__ bool settle_row_call(int64_t r, bool full, uint32_t my_word, int lane, const float *mu, const float *std,
                                             const float *lstd, const float *cb, const float *wrow, const SegHdr *seg_hdr,
                                             const uint32_t *rec_cnt, int64_t N, int nseg, float win_rel, int64_t *idx_out,
                                             float *zhat_out, const TailArgs &ta)
{
    return settle_row<G>(r, full, my_word, lane, mu, std, lstd, cb, ta.nlpb, wrow, seg_hdr, ta.cand, rec_cnt, N, ta.tiles,
                         ta.steps_per_cta, nseg, ta.unit_rows, win_rel, idx_out, zhat_out);
}

struct RowIn { float w, margin, pm, ps; uint4 q; uint32_t recs; };

template <int G>
__device__ __forceinline__ RowIn fetch_row(int64_t r, int lane, const float *__restrict__ wrow,
                                           const float *__restrict__ margin_row, const SegHdr *__restrict__ seg_hdr,
                                           const float *__restrict__ mu, const float *__restrict__ std,
                                           const uint32_t *__restrict__ rec_cnt, int nseg)
{
    RowIn in;
    const int seg_l = lane >> 2, quarter = lane & 3;
    in.w = __ldg(wrow + r);
    in.margin = __ldg(margin_row + r);
    // a summary counts only if its segment recorded something in THIS call (the prep kernel clears the
    // counters, not the summaries): both loads are issued together and the select (row_summary) happens at the
    // USE, one row later -- selecting here would park the warp on the prefetch it has just issued
    in.recs = 0u;
    in.q = make_uint4(0u, 0u, 0u, 0u);
    if (seg_l < nseg) {
        in.recs = __ldg(rec_cnt + r * nseg + seg_l);
        in.q = __ldg(reinterpret_cast<const uint4 *>(seg_hdr + r * nseg + seg_l) + quarter);
    }
    in.pm = in.ps = 0.f;
    if (lane < G) {
        in.pm = __ldg(mu + r * G + lane);
        in.ps = __ldg(std + r * G + lane);
    }
    return in;
}

// One warp per row.  All a row needs arrives in ONE round of independent loads (window, margin, the
// 64-byte summaries of its <= 8 segments -- lane i takes 16-byte quarter i%4 of segment i/4; untouched
// segments read as empty -- and mu/std/log std on lanes < g), issued one row ahead so that the round
// trip overlaps the previous row's math.  The surviving chunks' codewords (lane = codeword) are the
// second and last round.  Plans with more than 8 segments per row (small M: a unit is spread over many
// CTAs) run rescore_many_kernel instead.
template <int G>
#ifndef GQ_RESCORE_BLOCKS
#define GQ_RESCORE_BLOCKS 4   // resident 256-thread blocks per SM the register budget is sized for (64 registers); 5 and 6 measured
#endif
__global__ void __launch_bounds__(kRescoreWarps * 32, (G <= 16) ? GQ_RESCORE_BLOCKS : 1)
rescore_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
               const float *__restrict__ cb, const float *__restrict__ wrow, const float *__restrict__ margin_row,
               const SegHdr *__restrict__ seg_hdr, int64_t M, int64_t N, int nseg, float beta,
               int64_t *__restrict__ idx_out, float *__restrict__ zhat_out, uint32_t *__restrict__ fb_rows,
               uint32_t *__restrict__ fb_count, uint32_t *__restrict__ settle_rows, uint32_t *__restrict__ settle_chunks,
               uint32_t *__restrict__ settle_count, const uint32_t *__restrict__ rec_cnt,
               unsigned long long *__restrict__ fb_keys, uint32_t *__restrict__ stats, float win_rel, const TailArgs ta)
{
    __shared__ __align__(16) float s_par[kRescoreWarps][2][G < 4 ? 4 : G];   // -mu, -1/den of the warp's row
    __shared__ uint32_t s_chunks[kRescoreWarps][kMaxChunks];
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * kRescoreWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kRescoreWarps;
    const float hb = 0.5f * beta;
    const int quarter = lane & 3;
    const bool head = quarter == 0;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t st_rows = 0, st_pre = 0, st_exact = 0, st_fb = 0, st_ovf = 0;

    {
        RowIn cur{}, nxt{};
        if (warp_global < M) cur = fetch_row<G>(warp_global, lane, wrow, margin_row, seg_hdr, mu, std, rec_cnt, nseg);
        for (int64_t r = warp_global; r < M; r += nwarps) {
            const bool has_next = r + nwarps < M;
            if (has_next)   // in flight while this row is processed
                nxt = fetch_row<G>(r + nwarps, lane, wrow, margin_row, seg_hdr, mu, std, rec_cnt, nseg);

            int verdict = 0;   // 0 = decided here, 1 = settle (summaries), 2 = settle (full lists), 3 = exact scan
            int best_idx = 0;
            int n_chunks_row = 0;
            const uint4 q = cur.recs != 0u ? cur.q : make_uint4(0xff800000u, 0u, 0xff800000u, 0u);
            float gmax = head ? __uint_as_float(q.x) : -CUDART_INF_F;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
            const bool overflow = __any_sync(0xffffffffu, head && q.y == kSegOverflow);
            if (!(cur.w >= 0.f) || !(gmax > -CUDART_INF_F)) verdict = 3;
            else if (overflow) {
                bool lost = false;
                if (lane < nseg) lost = rec_cnt[r * nseg + lane] > (uint32_t)kCap;
                verdict = __any_sync(0xffffffffu, lost) ? 3 : 2;
            } else {
                const float thr = win_thr(gmax, cur.w, win_rel);
                // surviving chunk ids -> shared memory
                // quarter 0 holds survivor 0 in (z,w); quarter k>0 holds survivors 2k-1 in (x,y) and 2k in (z,w)
                const uint32_t cnt = __shfl_sync(0xffffffffu, q.y, lane & ~3);
                const uint32_t slot_a = head ? 0u : (uint32_t)(2 * quarter - 1);
                const float sc_a = __uint_as_float(head ? q.z : q.x);
                const uint32_t id_a = head ? q.w : q.y;
                const bool ok_a = slot_a < cnt && sc_a >= thr;
                const bool ok_b = !head && (uint32_t)(2 * quarter) < cnt && __uint_as_float(q.z) >= thr;
                const unsigned ma = __ballot_sync(0xffffffffu, ok_a);
                const unsigned mb = __ballot_sync(0xffffffffu, ok_b);
                const int na = __popc(ma);
                const int nchunks = na + __popc(mb);
                n_chunks_row = nchunks;
                if (ok_a) s_chunks[wib][__popc(ma & lt)] = id_a;               // <= 8*7 = 56 < kMaxChunks
                if (ok_b) s_chunks[wib][na + __popc(mb & lt)] = q.w;
                if (lane < G) {
                    const float dn = __fmul_rn(2.0f, __fmul_rn(cur.ps, cur.ps));
                    s_par[wib][0][lane] = -cur.pm;          // negated: see prefilter_score_smem
                    s_par[wib][1][lane] = -(1.0f / dn);
                }
                __syncwarp();
                if (nchunks == 0) verdict = 3;
                else {
                    // fp32 FMA prefilter over the surviving chunks: per-lane best and runner-up
                    float a_best = -CUDART_INF_F, a_second = -CUDART_INF_F;
                    int n_best = 0;
                    st_pre += (uint32_t)nchunks;
                    constexpr int kPer = 32 / kChunk;   // chunks per warp pass (lane = codeword)
                    for (int c = 0; c < nchunks; c += kPer) {
                        const int ci = c + lane / kChunk;
                        const int64_t n0 = ci < nchunks ? (int64_t)s_chunks[wib][ci] * kChunk + lane % kChunk : N;
                        if (n0 < N) {
                            float x0[G];
                            load_row<G>(cb + n0 * G, x0);
                            const float a2 = prefilter_score_smem<G>(x0, &s_par[wib][0][0], &s_par[wib][1][0], hb);
                            if (a2 > a_best) { a_second = a_best; a_best = a2; n_best = (int)n0; }
                            else if (a2 > a_second) a_second = a2;
                        }
                    }
                    float g2 = a_best;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) g2 = fmaxf(g2, __shfl_xor_sync(0xffffffffu, g2, off));
                    const float lim = g2 - cur.margin;
                    const unsigned firsts = __ballot_sync(0xffffffffu, a_best >= lim);
                    const unsigned seconds = __ballot_sync(0xffffffffu, a_second >= lim);
                    if (!(g2 > -CUDART_INF_F) || !(cur.margin >= 0.f) || firsts == 0u) verdict = 3;   // NaN / -inf rows
                    else if (__popc(firsts) == 1 && seconds == 0u) best_idx = __shfl_sync(0xffffffffu, n_best, __ffs(firsts) - 1);
                    else verdict = 1;   // several codewords within the prefilter margin: exact op order decides
                }
                __syncwarp();
            }
            if (verdict == 0) {
                if (lane == 0) idx_out[r] = (int64_t)best_idx;
                if (zhat_out && lane < G) zhat_out[r * G + lane] = __ldg(cb + (int64_t)best_idx * G + lane);
            } else if (kRescoreFinishes) {
                // finish the row here (no settle_kernel launch): exact-order settlement of the listed chunks, resp. the
                // pruned exact scan for a row the search could not certify; leftovers go to the sliced full scan
                bool done;
                if (verdict == 3) {
                    done = pruned_exact_row<G>(r, mu, std, lstd, cb, ta.nlpb, ta.meta, ta.sort_val, ta.sort_idx, N, beta, lane, idx_out, zhat_out);
                } else {
                    const bool listed = verdict == 1 && n_chunks_row <= 7;
                    uint32_t v = listed ? (uint32_t)n_chunks_row : 0xffffffffu;
                    if (lane > 0) v = (listed && lane < 8 && lane - 1 < n_chunks_row) ? s_chunks[wib][lane - 1] : 0u;
                    done = settle_row_call<G>(r, verdict == 2, v, lane, mu, std, lstd, cb, wrow, seg_hdr, rec_cnt, N, nseg, win_rel,
                                              idx_out, zhat_out, ta);
                }
                if (!done && lane == 0) { fb_keys[r] = 0ull; idx_out[r] = -1; ta.fb2_rows[atomicAdd(ta.fb2_count, 1u)] = (uint32_t)r; }
                __syncwarp();
            } else if (verdict == 3) {
                if (lane == 0) { fb_keys[r] = 0ull; fb_rows[atomicAdd(fb_count, 1u)] = (uint32_t)r; }
            } else {
                // hand the row to settle_kernel together with the chunk list just built (lanes 1..7), so
                // that it need not walk the segment summaries again (a chain of dependent loads)
                uint32_t slot = 0;
                if (lane == 0) {
                    slot = atomicAdd(settle_count, 1u);
                    settle_rows[slot] = (uint32_t)r | (verdict == 2 ? 0x80000000u : 0u);
                }
                slot = __shfl_sync(0xffffffffu, slot, 0);
                if (lane < 8) {
                    const bool listed = verdict == 1 && n_chunks_row <= 7;
                    uint32_t v = listed ? (uint32_t)n_chunks_row : 0xffffffffu;
                    if (lane > 0) v = (listed && lane - 1 < n_chunks_row) ? s_chunks[wib][lane - 1] : 0u;
                    settle_chunks[(int64_t)slot * 8 + lane] = v;
                }
            }
            st_rows += 1;
            st_exact += verdict == 1 ? 1u : 0u;
            st_fb += verdict == 3 ? 1u : 0u;
            st_ovf += overflow ? 1u : 0u;
            if (has_next) cur = nxt;
        }
    }
    if (stats && lane == 0) {
        atomicAdd(&stats[0], st_rows);
        atomicAdd(&stats[1], st_pre * (uint32_t)kChunk);
        atomicAdd(&stats[2], st_fb);
        atomicAdd(&stats[3], st_ovf);
        atomicAdd(&stats[4], st_exact);
    }
}

// The same decision for plans with more than 8 segments per row (small M, e.g. one 256-px image = 1024 rows:
// its 4 units are spread over all SMs, 76 short segments per row).  One warp per row, lane = segment: a
// first pass reduces the running maxima (8 bytes per segment), a second pass visits only the segments
// whose maximum lies inside the row's window (one or two) and collects their surviving chunk ids; the
// prefilter and the hand-over to settle_kernel are those of rescore_kernel.  Before this kernel every
// row of such a plan walked all its summaries in settle_kernel in the exact op order (23 us of a 59 us
// single-image call).
template <int G>
__global__ void __launch_bounds__(kRescoreWarps * 32)
rescore_many_kernel(const float *__restrict__ mu, const float *__restrict__ std, const float *__restrict__ lstd,
                    const float *__restrict__ cb, const float *__restrict__ wrow, const float *__restrict__ margin_row,
                    const SegHdr *__restrict__ seg_hdr, int64_t M, int64_t N, int nseg, float beta,
                    int64_t *__restrict__ idx_out, float *__restrict__ zhat_out, uint32_t *__restrict__ fb_rows,
                    uint32_t *__restrict__ fb_count, uint32_t *__restrict__ settle_rows, uint32_t *__restrict__ settle_chunks,
                    uint32_t *__restrict__ settle_count, const uint32_t *__restrict__ rec_cnt,
                    unsigned long long *__restrict__ fb_keys, uint32_t *__restrict__ stats, float win_rel,
                    const float *__restrict__ nlpb, const Cand *__restrict__ cand, int tiles, int steps_per_cta, int unit_rows,
                    const float *__restrict__ meta, const float *__restrict__ sort_val, const uint32_t *__restrict__ sort_idx)
{
    __shared__ __align__(16) float s_par[kRescoreWarps][2][G < 4 ? 4 : G];   // -mu, -1/den of the warp's row
    __shared__ uint32_t s_chunks[kRescoreWarps][kMaxChunks];
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * kRescoreWarps + wib;
    const int64_t nwarps = (int64_t)gridDim.x * kRescoreWarps;
    const float hb = 0.5f * beta;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t st_rows = 0, st_pre = 0, st_exact = 0, st_fb = 0, st_ovf = 0;

    for (int64_t r = warp_global; r < M; r += nwarps) {
        const float w = __ldg(wrow + r);
        const float margin = __ldg(margin_row + r);
        float pm = 0.f, ps = 1.f;
        if (lane < G) {
            pm = __ldg(mu + r * G + lane);
            ps = __ldg(std + r * G + lane);
        }
        // pass 1: global running maximum over the segments that recorded something in THIS call.  Up to 128
        // segments (every plan at N <= 65536) stay in registers, four per lane, with the counter and the summary
        // head of all four loaded together (eight independent loads instead of a counter -> summary chain per trip).
        constexpr int kSlots = 4;
        const bool in_regs = nseg <= 32 * kSlots;
        float rm[kSlots];
        uint32_t cn[kSlots];
        float gmax = -CUDART_INF_F;
        bool overflow = false, lost = false;
        if (in_regs) {
            uint32_t rc[kSlots];
            uint2 hh[kSlots];
#pragma unroll
            for (int u = 0; u < kSlots; ++u) {
                const int sg = lane + 32 * u;
                rc[u] = 0u;
                hh[u] = make_uint2(0xff800000u, 0u);
                if (sg < nseg) {
                    rc[u] = __ldg(rec_cnt + r * nseg + sg);
                    hh[u] = __ldg(reinterpret_cast<const uint2 *>(seg_hdr + r * nseg + sg));
                }
            }
#pragma unroll
            for (int u = 0; u < kSlots; ++u) {
                const bool rec = rc[u] != 0u;
                rm[u] = rec ? __uint_as_float(hh[u].x) : -CUDART_INF_F;
                cn[u] = rec ? hh[u].y : 0u;
                gmax = fmaxf(gmax, rm[u]);
                overflow |= rec && hh[u].y == kSegOverflow;
                lost |= rc[u] > (uint32_t)kCap;
            }
        } else {
            for (int sg = lane; sg < nseg; sg += 32) {
                const uint32_t recs = __ldg(rec_cnt + r * nseg + sg);
                if (recs != 0u) {
                    const uint2 h = __ldg(reinterpret_cast<const uint2 *>(seg_hdr + r * nseg + sg));
                    gmax = fmaxf(gmax, __uint_as_float(h.x));
                    overflow |= h.y == kSegOverflow;
                    lost |= recs > (uint32_t)kCap;
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
        overflow = __any_sync(0xffffffffu, overflow);
        lost = __any_sync(0xffffffffu, lost);

        int verdict = 0;   // 0 = decided here, 1 = settle (listed chunks / summaries), 2 = settle (full lists), 3 = exact scan
        int best_idx = 0;
        int nchunks = 0;
        if (!(w >= 0.f) || !(gmax > -CUDART_INF_F)) verdict = 3;
        else if (overflow) verdict = lost ? 3 : 2;
        else {
            const float thr = win_thr(gmax, w, win_rel);
            // pass 2: only segments whose own maximum is inside the window can hold a surviving record
            auto visit = [&](int sb, bool live, uint32_t cnt) {
                unsigned mask = __ballot_sync(0xffffffffu, live);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t c_n = __shfl_sync(0xffffffffu, cnt, src);   // <= kSurv (no overflow on this path)
                    const SegHdr *h = seg_hdr + r * nseg + sb + src;
                    Cand c{-CUDART_INF_F, 0u};
                    if ((uint32_t)lane < c_n) c = h->surv[lane];
                    const bool ok = (uint32_t)lane < c_n && c.score >= thr;
                    const unsigned m2 = __ballot_sync(0xffffffffu, ok);
                    const int pos = nchunks + __popc(m2 & lt);
                    if (ok && pos < kMaxChunks) s_chunks[wib][pos] = c.idx;
                    nchunks += __popc(m2);
                }
            };
            if (in_regs) {
#pragma unroll
                for (int u = 0; u < kSlots; ++u)
                    if (32 * u < nseg) visit(32 * u, rm[u] >= thr, cn[u]);
            } else {
                for (int sb = 0; sb < nseg; sb += 32) {
                    const int sg = sb + lane;
                    bool live = false;
                    uint32_t cnt = 0u;
                    if (sg < nseg && __ldg(rec_cnt + r * nseg + sg) != 0u) {
                        const uint2 h = __ldg(reinterpret_cast<const uint2 *>(seg_hdr + r * nseg + sg));
                        live = __uint_as_float(h.x) >= thr;
                        cnt = h.y;
                    }
                    visit(sb, live, cnt);
                }
            }
            if (lane < G) {
                const float dn = __fmul_rn(2.0f, __fmul_rn(ps, ps));
                s_par[wib][0][lane] = -pm;          // negated: see prefilter_score_smem
                s_par[wib][1][lane] = -(1.0f / dn);
            }
            __syncwarp();
            if (nchunks > kMaxChunks) verdict = lost ? 3 : 2;   // settle_kernel walks the complete record lists
            else if (nchunks == 0) verdict = 3;
            else {
                float a_best = -CUDART_INF_F, a_second = -CUDART_INF_F;
                int n_best = 0;
                st_pre += (uint32_t)nchunks;
                constexpr int kPer = 32 / kChunk;   // chunks per warp pass (lane = codeword)
                for (int c = 0; c < nchunks; c += kPer) {
                    const int ci = c + lane / kChunk;
                    const int64_t n0 = ci < nchunks ? (int64_t)s_chunks[wib][ci] * kChunk + lane % kChunk : N;
                    if (n0 < N) {
                        float x0[G];
                        load_row<G>(cb + n0 * G, x0);
                        const float a2 = prefilter_score_smem<G>(x0, &s_par[wib][0][0], &s_par[wib][1][0], hb);
                        if (a2 > a_best) { a_second = a_best; a_best = a2; n_best = (int)n0; }
                        else if (a2 > a_second) a_second = a2;
                    }
                }
                float g2 = a_best;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) g2 = fmaxf(g2, __shfl_xor_sync(0xffffffffu, g2, off));
                const float lim = g2 - margin;
                const unsigned firsts = __ballot_sync(0xffffffffu, a_best >= lim);
                const unsigned seconds = __ballot_sync(0xffffffffu, a_second >= lim);
                if (!(g2 > -CUDART_INF_F) || !(margin >= 0.f) || firsts == 0u) verdict = 3;   // NaN / -inf rows
                else if (__popc(firsts) == 1 && seconds == 0u) best_idx = __shfl_sync(0xffffffffu, n_best, __ffs(firsts) - 1);
                else verdict = 1;   // several codewords within the prefilter margin: exact op order decides
            }
        }
        // Plans with this many segments are small calls (one image), where every dependent launch is a fixed cost:
        // the row is finished HERE -- ambiguous rows are settled in the exact op order, uncertified rows take the pruned
        // exact scan -- and only what neither decides goes on to the sliced full scan.  No settle_kernel launch.
        if (verdict == 0) {
            if (lane == 0) idx_out[r] = (int64_t)best_idx;
            if (zhat_out && lane < G) zhat_out[r * G + lane] = __ldg(cb + (int64_t)best_idx * G + lane);
        } else {
            bool done;
            if (verdict == 3) {
                done = pruned_exact_row<G>(r, mu, std, lstd, cb, nlpb, meta, sort_val, sort_idx, N, beta, lane, idx_out, zhat_out);
            } else {
                const bool listed = verdict == 1 && nchunks <= 7;
                uint32_t v = listed ? (uint32_t)nchunks : 0xffffffffu;
                if (lane > 0) v = (listed && lane < 8 && lane - 1 < nchunks) ? s_chunks[wib][lane - 1] : 0u;
                done = settle_row<G>(r, verdict == 2, v, lane, mu, std, lstd, cb, nlpb, wrow, seg_hdr, cand, rec_cnt, N, tiles,
                                     steps_per_cta, nseg, unit_rows, win_rel, idx_out, zhat_out);
            }
            if (!done && lane == 0) { fb_keys[r] = 0ull; idx_out[r] = -1; fb_rows[atomicAdd(fb_count, 1u)] = (uint32_t)r; }
        }
        __syncwarp();   // s_chunks / s_par are rewritten for the next row
        st_rows += 1;
        st_exact += verdict == 1 ? 1u : 0u;
        st_fb += verdict == 3 ? 1u : 0u;
        st_ovf += overflow ? 1u : 0u;
    }
    if (stats && lane == 0) {
        atomicAdd(&stats[0], st_rows);
        atomicAdd(&stats[1], st_pre * (uint32_t)kChunk);
        atomicAdd(&stats[2], st_fb);
        atomicAdd(&stats[3], st_ovf);
        atomicAdd(&stats[4], st_exact);
    }
}

// ---- host side ------------------------------------------------------------------------------------
template <int G, int KPAD, bool SPLIT>
static int run_tc(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                  int64_t *idx, float *zhat, const TcPlan &p, const TcWorkspace &w, uint32_t *stats,
                  float w_max, cudaStream_t stream, bool do_prep = true, const FinalLayout *fin = nullptr)
{
    using K = SearchKernel<G, KPAD, SPLIT, false>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(K::fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmem);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
        attr_set = true;
    }
    // (the device-side lists / arrival counters are zeroed by the prep or front-end kernel: a memset node
    // between two kernels would break the programmatic launch chain)
    int rc = GQ_OK;
    if (do_prep) {
        rows_prep_kernel<G, KPAD, SPLIT><<<(unsigned)((p.M_pad + 127) / 128), 128, 0, stream>>>(
            mu, std, lstd, cb->meta, w.a_img, w.wrow, w.margin, w.seg, w.rec_cnt, p.nseg, M, p.M_pad, cb->beta, w_max,
            w.counters, w.counter_bytes / 4, cb->wide_dev);
        rc = check_launch("rows_prep_kernel");
        if (rc) return rc;
        count_launch(1);
    }

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // an event between two kernels orders them fully, i.e. switches the programmatic launch overlap off
    // at that boundary: sample every n-th launch so that the measurement barely perturbs what it measures
    const bool g_prof = g_prof_period > 0 && (g_prof_calls++ % (unsigned)g_prof_period) == 0;
    if (g_prof) {
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }
    constexpr bool ACC16 = gq_acc16(G, SPLIT);
    constexpr float win_rel = ACC16 ? kWinRel16 : 0.f;
    launch_chained(K::fn(), dim3(p.ctas), dim3(K::kBlock), K::kSmem, stream,
                   w.a_img, (SPLIT && gq_has_robust_image(G)) ? cb->bimg_robust : cb->bimg, w.wrow, w.seg, w.cand, w.rec_cnt, cb->N,
                   p.tiles, p.total_steps, p.steps_per_cta, p.nseg, nullptr);
    if (g_prof) {
        cudaEventRecord(ev1, stream);
        g_prof_events.emplace_back(ev0, ev1);
    }
    rc = check_launch("quantize_tc_kernel");
    if (rc) return rc;

    {
        int64_t need = (M + kRescoreWarps - 1) / kRescoreWarps;
        int blocks = (int)imin(need, (int64_t)device_sm_count() * 8);
        if (p.nseg > 8) {
            // small calls: rescore_many_kernel finishes its rows itself (settlement and pruned scan inline);
            // its leftovers go straight to the sliced scan's list
            launch_chained(rescore_many_kernel<G>, dim3(blocks), dim3(kRescoreWarps * 32), 0, stream,
                           mu, std, lstd, cb->cb, w.wrow, w.margin, w.seg, M, cb->N, p.nseg, cb->beta, idx, zhat, w.fb_rows2,
                           w.counters + 2, w.st_rows, w.st_chunks, w.counters + 1, w.rec_cnt, w.fb_keys, stats, win_rel,
                           cb->nlpb, w.cand, p.tiles, p.steps_per_cta, p.unit_rows, cb->meta, cb->sort_val, cb->sort_idx);
            rc = check_launch("rescore_many_kernel");
            if (rc) return rc;
            count_launch(1);
        } else {
            const TailArgs ta{cb->nlpb, w.cand, p.tiles, p.steps_per_cta, p.unit_rows, cb->meta, cb->sort_val, cb->sort_idx, w.fb_rows2, w.counters + 2};
            launch_chained(rescore_kernel<G>, dim3(blocks), dim3(kRescoreWarps * 32), 0, stream,
                           mu, std, lstd, cb->cb, w.wrow, w.margin, w.seg, M, cb->N, p.nseg, cb->beta, idx, zhat, w.fb_rows,
                           w.counters, w.st_rows, w.st_chunks, w.counters + 1, w.rec_cnt, w.fb_keys, stats, win_rel, ta);
            rc = check_launch("rescore_kernel");
            if (rc) return rc;
            if (kRescoreFinishes) { count_launch(1); goto tail_done; }
            int sblocks = (int)imin((M + 3) / 4, (int64_t)device_sm_count() * 8);
            launch_chained(settle_kernel<G>, dim3(sblocks), dim3(128), 0, stream, mu, std, lstd, cb->cb, cb->nlpb, w.wrow, w.seg,
                           w.cand, w.rec_cnt, w.st_rows, w.st_chunks, w.counters + 1, cb->N, p.tiles, p.steps_per_cta, p.nseg, p.unit_rows, idx, zhat,
                           w.fb_rows2, w.counters + 2, w.fb_keys, win_rel, w.fb_rows, w.counters, cb->meta, cb->sort_val, cb->sort_idx, cb->beta);
            rc = check_launch("settle_kernel");
            if (rc) return rc;
            count_launch(2);
        }
    }
tail_done:
    count_launch(1);   // the search kernel

    // exact scan over the rows the search could not certify (device-side count, no host sync),
    // sliced over the codebook so that a handful of rows still spreads over the whole GPU
    return launch_exact_sliced(cb, mu, std, lstd, M, w.fb_rows2, w.counters + 2, w.fb_keys, w.counters + 64, idx, zhat, stream, fin);
}

// Debug/diagnostic path (tests only): runs row prep + the search with the DUMP epilogue and
// returns the raw fp16 tensor-core scores [M_pad, n_pad] plus the scaled windows [M_pad].
template <int G, int KPAD, bool SPLIT>
static int run_tc_dump(const gq_codebook *cb, const float *mu, const float *std, const float *lstd, int64_t M,
                       const TcPlan &p, const TcWorkspace &w, float *dump, float *wrow_out, float w_max,
                       cudaStream_t stream)
{
    using K = SearchKernel<G, KPAD, SPLIT, true>;
    cudaError_t e = cudaFuncSetAttribute(K::fn(), cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    rows_prep_kernel<G, KPAD, SPLIT><<<(unsigned)((p.M_pad + 127) / 128), 128, 0, stream>>>(
        mu, std, lstd, cb->meta, w.a_img, w.wrow, w.margin, w.seg, w.rec_cnt, p.nseg, M, p.M_pad, cb->beta, w_max,
        w.counters, w.counter_bytes / 4, cb->wide_dev);
    int rc = check_launch("rows_prep_kernel");
    if (rc) return rc;
    K::fn()<<<p.ctas, K::kBlock, K::kSmem, stream>>>(
        w.a_img, (SPLIT && gq_has_robust_image(G)) ? cb->bimg_robust : cb->bimg, w.wrow, w.seg, w.cand, w.rec_cnt, cb->N, p.tiles, p.total_steps, p.steps_per_cta, p.nseg,
        dump);
    rc = check_launch("quantize_tc_kernel<dump>");
    if (rc) return rc;
    e = cudaMemcpyAsync(wrow_out, w.wrow, sizeof(float) * p.M_pad, cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) { set_error("memcpy: %s", cudaGetErrorString(e)); return GQ_ERR_CUDA; }
    count_launch(2);
    return GQ_OK;
}


// ---- fused regularizer front-end: raw latent -> rows, operand image, windows, zhat_noquant ----------
// Replaces the torch prologue of the eval branch (gaussian.py:62-80,121-123: float, rearrange, chunk,
// clamp, exp, randn*std+mu, the reshape/permute into rows, and std.log()) plus rows_prep_kernel.
// Thread <-> (token, sub-codebook j), addressing as in gq_kl.cu.  std = exp(0.5*clamp(logvar)) and
// log(std) use the same libdevice expf/logf torch's CUDA kernels call; zhat_noquant = mu + eps*std is
// formed with one multiply and one add (no FMA), as torch's two kernels do.
template <int G, int KPAD, bool SPLIT>
__global__ void __launch_bounds__(128)
frontend_kernel(const float *__restrict__ z, const float *__restrict__ eps, const float *__restrict__ meta,
                float *__restrict__ mu_r, float *__restrict__ std_r,
                float *__restrict__ zhat_nq, __half *__restrict__ a_img, float *__restrict__ wrow,
                float *__restrict__ margin, SegHdr *__restrict__ seg_hdr, uint32_t *__restrict__ rec_cnt, int nseg,
                int64_t b, int64_t l, int C, int layout, float lv_lo, float lv_hi, int64_t M, int64_t M_pad,
                float beta, float w_max, uint32_t *__restrict__ counters, int64_t counter_words, uint32_t *__restrict__ wide_dev)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < counter_words) counters[i] = 0u;   // device-side lists / arrival counters of this call (counter_words <= M_pad)
    if (i >= M_pad) return;
    const int ncb = C / G;
    float m[G], s[G], ls[G];
    int64_t r = i;
    const bool valid = i < M;
    if (valid) {
        // eps / zhat_noquant share zhat's layout: [b,C,h,w] or [b,l,C] (base eb, channel stride es)
        int64_t zb, zs, lvo, eb, es;
        if (layout == GQ_LAYOUT_BCHW) {
            const int64_t t = i % l;
            const int j = (int)((i / l) % ncb);
            const int64_t bi = i / (l * ncb);
            zb = (bi * 2 * C + j) * l + t;
            zs = (int64_t)ncb * l;
            lvo = (int64_t)C * l;
            r = (bi * l + t) * ncb + j;
            eb = (bi * C + j) * l + t;
            es = (int64_t)ncb * l;
        } else {
            const int j = (int)(i % ncb);
            const int64_t tok = i / ncb;
            zb = tok * 2 * C + j;
            zs = ncb;
            lvo = C;
            eb = tok * C + j;
            es = ncb;
        }
#pragma unroll
        for (int k = 0; k < G; ++k) {
            m[k] = __ldg(z + zb + k * zs);
            const float lv = clamp_torch(__ldg(z + zb + lvo + k * zs), lv_lo, lv_hi);
            s[k] = expf(__fmul_rn(0.5f, lv));
            ls[k] = logf(s[k]);
            if (zhat_nq) zhat_nq[eb + k * es] = __fadd_rn(m[k], __fmul_rn(__ldg(eps + eb + k * es), s[k]));
        }
        if constexpr (G % 4 == 0) {
#pragma unroll
            for (int k = 0; k < G / 4; ++k) {
                reinterpret_cast<float4 *>(mu_r + r * G)[k] = make_float4(m[4 * k], m[4 * k + 1], m[4 * k + 2], m[4 * k + 3]);
                reinterpret_cast<float4 *>(std_r + r * G)[k] = make_float4(s[4 * k], s[4 * k + 1], s[4 * k + 2], s[4 * k + 3]);
            }
        } else {
#pragma unroll
            for (int k = 0; k < G; ++k) { mu_r[r * G + k] = m[k]; std_r[r * G + k] = s[k]; }
        }
    } else {
#pragma unroll
        for (int k = 0; k < G; ++k) { m[k] = 0.f; s[k] = 1.f; ls[k] = 0.f; }
    }
    prep_row<G, KPAD, SPLIT>(m, s, ls, meta, beta, w_max, valid, a_img + r * KPAD, wrow + r, margin + r,
                      seg_hdr + r * nseg, rec_cnt + r * nseg, nseg, wide_dev);
}

// The same front-end for the plain fp16 image of group 16 with FOUR lanes per row (each lane takes four of the
// sixteen dimensions): the thread-per-row kernel above walks ~2 K dependent instructions per row on 64 K threads
// (33 % of the HBM rate at batch 64, and 13 us for ONE image, where it occupies eight SMs); here the per-row
// sums and maxima of the window are folded over the four lanes with two shuffles each.  The window is a set of
// upper bounds, so the different summation order only moves it within its 1 % slack; everything that defines
// results -- mu, std = exp(0.5 clamp(logvar)), log(std), zhat_noquant and the fp16 image values -- is computed per
// element exactly as above.
template <int KPAD>
__global__ void __launch_bounds__(256)
frontend16x4_kernel(const float *__restrict__ z, const float *__restrict__ eps, const float *__restrict__ meta,
                    float *__restrict__ mu_r, float *__restrict__ std_r, float *__restrict__ zhat_nq,
                    __half *__restrict__ a_img, float *__restrict__ wrow, float *__restrict__ margin_out,
                    uint32_t *__restrict__ rec_cnt, int nseg, int64_t b, int64_t l, int C, int layout, float lv_lo,
                    float lv_hi, int64_t M, int64_t M_pad, float beta, float w_max, uint32_t *__restrict__ counters,
                    int64_t counter_words, uint32_t *__restrict__ wide_dev)
{
    constexpr int G = 16, DL = 4;
    static_assert(KPAD == 32, "plain image of group 16");
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < counter_words) counters[gid] = 0u;   // device-side lists / arrival counters of this call (counter_words <= M_pad)
    const int64_t i = gid >> 2;
    const int sub = (int)(gid & 3);
    if (i >= M_pad) return;                        // whole quads leave together
    const int ncb = C / G;
    const bool valid = i < M;
    int64_t r = i;
    float a[DL], bq[DL];
    float amax = 0.f, bnorm2 = 0.f, R = 0.f, p_naive = 0.f, maxA = 0.f, sumA = 0.f, eA = 0.f;
    if (valid) {
        int64_t zb, zs, lvo, eb, es;
        if (layout == GQ_LAYOUT_BCHW) {
            const int64_t t = i % l;
            const int j = (int)((i / l) % ncb);
            const int64_t bi = i / (l * ncb);
            zb = (bi * 2 * C + j) * l + t;
            zs = (int64_t)ncb * l;
            lvo = (int64_t)C * l;
            r = (bi * l + t) * ncb + j;
            eb = (bi * C + j) * l + t;
            es = (int64_t)ncb * l;
        } else {
            const int j = (int)(i % ncb);
            const int64_t tok = i / ncb;
            zb = tok * 2 * C + j;
            zs = ncb;
            lvo = C;
            eb = tok * C + j;
            es = ncb;
        }
        float m[DL], sd[DL];
#pragma unroll
        for (int kk = 0; kk < DL; ++kk) {
            const int k = sub * DL + kk;
            m[kk] = __ldg(z + zb + k * zs);
            const float lv = clamp_torch(__ldg(z + zb + lvo + k * zs), lv_lo, lv_hi);
            sd[kk] = expf(__fmul_rn(0.5f, lv));
            const float ls = logf(sd[kk]);
            if (zhat_nq) zhat_nq[eb + k * es] = __fadd_rn(m[kk], __fmul_rn(__ldg(eps + eb + k * es), sd[kk]));
            const float v = sd[kk] * sd[kk];
            const float rr = 0.5f / v;                  // 1/(2v)
            a[kk] = 0.5f * beta - rr;
            bq[kk] = m[kk] / v;
            const float x2m = meta[2 + k], xm = meta[2 + G + k];
            amax = fmaxf(amax, fabsf(a[kk]));
            bnorm2 = fmaf(bq[kk], bq[kk], bnorm2);
            maxA = fmaxf(maxA, fmaxf(fabsf(a[kk]), fabsf(bq[kk])));
            p_naive += fabsf(a[kk]) * x2m + fabsf(bq[kk]) * xm;
            sumA += fabsf(a[kk]) + fabsf(bq[kk]);
            eA += (0.5f * beta + rr) * x2m + fabsf(bq[kk]) * xm;
            const float q = xm + fabsf(m[kk]);
            R += q * q * rr + fabsf(ls) + 2.0f * fmaxf(1.0f, beta) + beta * x2m;
        }
        reinterpret_cast<float4 *>(mu_r + r * G)[sub] = make_float4(m[0], m[1], m[2], m[3]);
        reinterpret_cast<float4 *>(std_r + r * G)[sub] = make_float4(sd[0], sd[1], sd[2], sd[3]);
        for (int sgi = sub; sgi < nseg; sgi += 4) rec_cnt[r * nseg + sgi] = 0u;   // (see prep_row: only the counters are cleared)
    } else {
#pragma unroll
        for (int kk = 0; kk < DL; ++kk) { a[kk] = 0.f; bq[kk] = 0.f; }
    }
    // fold the four lanes of the row (xor 1, 2 stay inside the quad; the whole warp is here: quads leave together)
#pragma unroll
    for (int off = 1; off <= 2; off <<= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, off));
        maxA = fmaxf(maxA, __shfl_xor_sync(0xffffffffu, maxA, off));
        bnorm2 += __shfl_xor_sync(0xffffffffu, bnorm2, off);
        R += __shfl_xor_sync(0xffffffffu, R, off);
        p_naive += __shfl_xor_sync(0xffffffffu, p_naive, off);
        sumA += __shfl_xor_sync(0xffffffffu, sumA, off);
        eA += __shfl_xor_sync(0xffffffffu, eA, off);
    }
    bool bad = !valid;
    if (!(maxA > 0.f) || !isfinite(maxA) || !isfinite(R) || !isfinite(bnorm2)) bad = true;
    float scale = 1.f, w_scaled = -1.f, margin = 0.f;
    bool wide = false;
    if (!bad) {       // the window of prep_row<16, 32, false> (fp16 accumulators), term by term
        int ex, exp_p;
        frexpf(maxA, &ex);
        const float u2 = sqrtf(bnorm2) * 1.000001f * meta[1];
        const float p_norm = amax * meta[0] + u2;
        const float P = fminf(p_naive, p_norm);
        frexpf(P * 1.002f, &exp_p);
        scale = gq_acc16(G, false) ? ldexpf(1.0f, imin(15 - exp_p, 15 - ex)) : ldexpf(1.0f, 8 - ex);
        if (!isfinite(P) || !(scale > 0.f) || !isfinite(scale)) bad = true;
        const float sumB = meta[2 + 2 * kMaxG];
        const float e_round = 1.01f * ldexpf(P, -10);
        const float e_floor = ldexpf(sumB / scale + sumA, -25);
        const float e_acc = gq_acc16(G, false) ? 1.004f * ldexpf(fminf(u2, amax * meta[0]), -11) + ldexpf(1.0f / scale, -24) : 0.f;
        const float e_tc = e_round + e_floor + ldexpf(eA, -22) + e_acc;
        const float W = 2.02f * (e_tc + ldexpf(R, -20));
        margin = 2.02f * 4.0f * ldexpf(R, -20);
        if (!(W <= w_max) || !isfinite(W * scale)) bad = true;
        else w_scaled = W * scale * 1.001f;
        wide = !(W <= w_max);      // plain image: W is the plain window (GQ_MODE_AUTO feedback)
    }
    {
        const unsigned mw = __ballot_sync(0xffffffffu, wide && sub == 0 && valid);
        if (mw && (threadIdx.x & 31) == (unsigned)(__ffs(mw) - 1)) atomicAdd(wide_dev, (uint32_t)__popc(mw));
    }
    // fp16 image row [a(16) | b(16)]: this lane's four a's and four b's, 8 bytes each
    __half2 ha[2], hb[2];
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
        ha[h2] = __floats2half2_rn(bad ? 0.f : a[2 * h2] * scale, bad ? 0.f : a[2 * h2 + 1] * scale);
        hb[h2] = __floats2half2_rn(bad ? 0.f : bq[2 * h2] * scale, bad ? 0.f : bq[2 * h2 + 1] * scale);
    }
    __half *arow = a_img + r * KPAD;
    *reinterpret_cast<uint2 *>(arow + sub * DL) = *reinterpret_cast<uint2 *>(ha);
    *reinterpret_cast<uint2 *>(arow + G + sub * DL) = *reinterpret_cast<uint2 *>(hb);
    if (sub == 0) {
        wrow[r] = w_scaled;
        margin_out[r] = margin;
    }
}

// Row indices -> the module's output tensors (gaussian.py:153-158): indices [b,l,ncb] / [b,ncb,h,w]
// and zhat [b,l,C] / [b,C,h,w] with channel k*ncb + j = codebook[idx][k].
template <int G>
__global__ void __launch_bounds__(256)
finalize_kernel(const int64_t *__restrict__ idx_rows, const float *__restrict__ cb, int64_t *__restrict__ idx_out,
                float *__restrict__ zhat_out, int64_t b, int64_t l, int ncb, int layout,
                unsigned long long *__restrict__ usage_hist)
{
    pdl_wait();
    const int64_t total = b * l * ncb;
    const int C = ncb * G;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t bi, t;
        int j;
        if (layout == GQ_LAYOUT_BCHW) { t = i % l; j = (int)((i / l) % ncb); bi = i / (l * ncb); }
        else { j = (int)(i % ncb); t = (i / ncb) % l; bi = i / (l * ncb); }
        const int64_t r = (bi * l + t) * ncb + j;
        const int64_t id = idx_rows[r];
        idx_out[i] = id;
        if (usage_hist) atomicAdd(usage_hist + id, 1ull);   // codebook usage (eval.py:152-154), fused
        float v[G];
        load_row<G>(cb + id * G, v);
        if (layout == GQ_LAYOUT_BCHW) {
            float *o = zhat_out + (bi * C + j) * l + t;
#pragma unroll
            for (int k = 0; k < G; ++k) o[(int64_t)k * ncb * l] = v[k];
        } else {
            float *o = zhat_out + (bi * l + t) * C + j;
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

struct LatentWorkspace {
    float *mu_r, *std_r, *lstd_r;
    int64_t *idx_rows;
    void *tc;
    int64_t tc_bytes;
    int64_t bytes;
};

static LatentWorkspace carve_latent(const gq_codebook *cb, int64_t M, void *base)
{
    LatentWorkspace w;
    char *ptr = reinterpret_cast<char *>(base);
    int64_t off = 0;
    const int64_t row_bytes = align_up(M * cb->g * 4, 256);
    w.mu_r = reinterpret_cast<float *>(ptr + off); off += row_bytes;
    w.std_r = reinterpret_cast<float *>(ptr + off); off += row_bytes;
    w.lstd_r = reinterpret_cast<float *>(ptr + off); off += row_bytes;
    w.idx_rows = reinterpret_cast<int64_t *>(ptr + off); off += align_up(M * 8, 256);
    w.tc = ptr + off;
    w.tc_bytes = tc_workspace_bytes(cb, M);
    off += w.tc_bytes;
    w.bytes = off;
    return w;
}

int64_t latent_workspace_bytes(const gq_codebook *cb, int64_t M)
{
    return carve_latent(cb, M > 0 ? M : 1, nullptr).bytes;
}

#ifndef GQ_TAIL_MERGE
#define GQ_TAIL_MERGE 1   // 0: separate finalize_kernel launch behind the exact-scan fallback (A/B)
#endif
static float g_w_max = 3.0f;

// Which operand image a call uses (groups 8 and 16; gq_mode in gq_b200.h).  AUTO reads the feedback the
// last completed call on this codebook left in pinned host memory -- no synchronisation, so the decision
// may lag by a call or two, which is fine for a stationary latent distribution.
static bool use_robust_image(const gq_codebook *cb)
{
    if (!cb->kpad_robust) return false;
    if (cb->mode == GQ_MODE_ROBUST) return true;
    if (cb->mode == GQ_MODE_FAST || !cb->fb_host) return false;
    const uint32_t wide = cb->fb_host[0], rows = cb->fb_host[1];
    return rows > 0 && (uint64_t)wide * 16u > (uint64_t)rows;
}

template <int G, int KPAD, bool SPLIT>
static int run_latent(const gq_codebook *cb, const float *z, const float *eps, int64_t b, int64_t l, int C,
                      int layout, float lv_lo, float lv_hi, int64_t *idx_out, float *zhat_out, float *zhat_nq,
                      const LatentWorkspace &lw, uint32_t *stats, cudaStream_t stream)
{
    const int ncb = C / G;
    const int64_t M = b * l * ncb;
    TcPlan p = make_plan(cb, M);
    TcWorkspace w = carve(cb, p, lw.tc);
    if constexpr (G == 16 && !SPLIT && GQ_FRONTEND_X4) {
        frontend16x4_kernel<KPAD><<<(unsigned)((p.M_pad * 4 + 255) / 256), 256, 0, stream>>>(
            z, eps, cb->meta, lw.mu_r, lw.std_r, zhat_nq, w.a_img, w.wrow, w.margin, w.rec_cnt, p.nseg, b, l, C,
            layout, lv_lo, lv_hi, M, p.M_pad, cb->beta, g_w_max, w.counters, w.counter_bytes / 4, cb->wide_dev);
    } else {
        frontend_kernel<G, KPAD, SPLIT><<<(unsigned)((p.M_pad + 127) / 128), 128, 0, stream>>>(
            z, eps, cb->meta, lw.mu_r, lw.std_r, zhat_nq, w.a_img, w.wrow, w.margin, w.seg, w.rec_cnt, p.nseg, b, l, C,
            layout, lv_lo, lv_hi, M, p.M_pad, cb->beta, g_w_max, w.counters, w.counter_bytes / 4, cb->wide_dev);
    }
    int rc = check_launch("frontend_kernel");
    if (rc) return rc;
    count_launch(1);
    if (layout == GQ_LAYOUT_BLC && ncb == 1) {
        // rows are tokens and a row's codeword is the token's channel vector: the rescoring kernels write
        // the final tensors themselves, no permute / gather pass is needed
        rc = run_tc<G, KPAD, SPLIT>(cb, lw.mu_r, lw.std_r, nullptr, M, idx_out, zhat_out, p, w, stats, g_w_max, stream,
                                    /*do_prep=*/false);
        if (rc == GQ_OK && cb->usage_hist) rc = launch_index_histogram(idx_out, M, cb->N, cb->usage_hist, nullptr, stream);
        return rc;
    }
    // (log std is not materialised: the settle / exact kernels take logf(std) with the libdevice function
    // the front-end and torch use, so the values are the same bits)
    if (GQ_TAIL_MERGE) {
        // the last kernel of the chain (the exact-scan fallback, empty for most calls) also writes the final tensors:
        // one launch less per call than a separate finalize pass
        const FinalLayout fin{idx_out, zhat_out, b, l, ncb, layout, cb->usage_hist};
        return run_tc<G, KPAD, SPLIT>(cb, lw.mu_r, lw.std_r, nullptr, M, lw.idx_rows, nullptr, p, w, stats, g_w_max, stream,
                                      /*do_prep=*/false, &fin);
    }
    rc = run_tc<G, KPAD, SPLIT>(cb, lw.mu_r, lw.std_r, nullptr, M, lw.idx_rows, nullptr, p, w, stats, g_w_max, stream,
                         /*do_prep=*/false);
    if (rc) return rc;
    int blocks = (int)imin((M + 255) / 256, (int64_t)device_sm_count() * 8);
    launch_chained(finalize_kernel<G>, dim3(blocks), dim3(256), 0, stream, lw.idx_rows, cb->cb, idx_out, zhat_out, b, l, ncb, layout,
                   cb->usage_hist);
    count_launch(1);
    return check_launch("finalize_kernel");
}

int launch_quantize_latent(const gq_codebook *cb, const float *z, const float *eps, int64_t b, int64_t l, int C,
                           int layout, float lv_lo, float lv_hi, int64_t *idx_out, float *zhat_out, float *zhat_nq,
                           void *ws, int64_t ws_bytes, uint32_t *stats, cudaStream_t stream)
{
    if (C % cb->g) { set_error("channels %d not divisible by group %d", C, cb->g); return GQ_ERR_INVALID; }
    const int64_t M = b * l * (C / cb->g);
    if (M == 0) return GQ_OK;
    if (M >= ((int64_t)1 << 31) || (M / kBlockRows + 1) * (cb->n_pad / kTileN) >= ((int64_t)1 << 31) ||
        (M + kUnitRows) * make_plan(cb, M).nseg >= ((int64_t)1 << 32)) {
        set_error("problem too large: rows x codebook tiles must stay below 2^31");
        return GQ_ERR_INVALID;
    }
    if (device_cc_major() != 10) {
        set_error("gq_quantize_latent needs an sm_100 device (tcgen05/TMEM); found cc major %d", device_cc_major());
        return GQ_ERR_UNSUPPORTED;
    }
    LatentWorkspace lw = carve_latent(cb, M, ws);
    if (ws == nullptr || ws_bytes < lw.bytes) {
        set_error("workspace too small: need %lld bytes, got %lld", (long long)lw.bytes, (long long)ws_bytes);
        return GQ_ERR_WORKSPACE;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) { set_error("workspace must be 256-byte aligned"); return GQ_ERR_INVALID; }
#define GQ_LAT(GG, KK, SS) return run_latent<GG, KK, SS>(cb, z, eps, b, l, C, layout, lv_lo, lv_hi, idx_out, zhat_out, zhat_nq, lw, stats, stream)
    const bool robust = use_robust_image(cb);
    switch (cb->g) {
    case 1: GQ_LAT(1, 16, true);
    case 2: GQ_LAT(2, 16, true);
    case 4: GQ_LAT(4, 32, true);
    case 8: if (robust) GQ_LAT(8, 48, true); else GQ_LAT(8, 16, false);
    case 16: if (robust) GQ_LAT(16, 96, true); else GQ_LAT(16, 32, false);
    case 32: GQ_LAT(32, 64, false);
    default: set_error("group %d unsupported", cb->g); return GQ_ERR_INVALID;
    }
#undef GQ_LAT
}

void set_tc_window_max(float w) { g_w_max = w; }

int launch_quantize_tc(const gq_codebook *cb, const float *mu, const float *std, const float *lstd,
                       int64_t M, int64_t *idx, float *zhat, void *ws, int64_t ws_bytes, uint32_t *stats,
                       cudaStream_t stream)
{
    if (M == 0) return GQ_OK;
    if (M >= ((int64_t)1 << 31) || (M / kBlockRows + 1) * (cb->n_pad / kTileN) >= ((int64_t)1 << 31) ||
        (M + kUnitRows) * make_plan(cb, M).nseg >= ((int64_t)1 << 32)) {
        set_error("problem too large: rows x codebook tiles must stay below 2^31");
        return GQ_ERR_INVALID;
    }
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
    const bool robust = use_robust_image(cb);
    switch (cb->g) {
    case 1: return run_tc<1, 16, true>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 2: return run_tc<2, 16, true>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 4: return run_tc<4, 32, true>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 8:
        if (robust) return run_tc<8, 48, true>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
        return run_tc<8, 16, false>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 16:
        if (robust) return run_tc<16, 96, true>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
        return run_tc<16, 32, false>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
    case 32: return run_tc<32, 64, false>(cb, mu, std, lstd, M, idx, zhat, p, w, stats, g_w_max, stream);
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
    case 4: return run_tc_dump<4, 32, true>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
    case 8: return run_tc_dump<8, 16, false>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
    case 16:
        if (cb->mode == GQ_MODE_ROBUST) return run_tc_dump<16, 96, true>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
        return run_tc_dump<16, 32, false>(cb, mu, std, lstd, M, p, w, dump, wrow_out, g_w_max, stream);
    default: set_error("debug dump supports g in {4,8,16}"); return GQ_ERR_INVALID;
    }
}

// Host-only: the stream-K decomposition for (N, M) on `sms` SMs, for CPU-side tests.
// out[0..5] = {units, tiles, ctas, steps_per_cta, nseg, M_pad}
void debug_plan(int64_t N, int64_t M, int sms, int acc16, int64_t *out)
{
    TcPlan p = make_plan_sms(N, M, sms, acc16 != 0);
    out[0] = p.units; out[1] = p.tiles; out[2] = p.ctas; out[3] = p.steps_per_cta; out[4] = p.nseg; out[5] = p.M_pad;
    out[6] = p.unit_rows; out[7] = p.tile_cols;
}

}  // namespace gq

#ifdef GQ_TC_TRACE
extern "C" int gq_debug_tc_trace(unsigned int *out)
{
    cudaMemcpyFromSymbol(out, gq::g_tc_trace, sizeof(unsigned int) * 32 * gq::kTraceN * 4);
    return gq::kTraceN;
}
#endif
