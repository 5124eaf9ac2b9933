// Probe: what does tcgen05.mma kind::f16 do with an fp16 accumulator (idesc D format = f16) on sm_100a?
//  1. where do the columns land when read with tcgen05.ld ... .pack::16b
//  2. how is the accumulator rounded (nearest / truncation; once per instruction or once per product)
//  3. is the error relative to the result or to the largest addend (internal alignment)
// One CTA: A[128,32] x B[128,32]^T, two K=16 MMAs (the second accumulates), fp16 operands, fp16 D.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f16acc_probe f16acc_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../vq-vae-from-gaussian-vae_b200/csrc/gq_ptx.cuh"
using namespace gq::ptx;

constexpr int K = 32;
__host__ __device__ inline int img_off(int r, int k) { return (r / 8) * (K / 8) * 128 + (k / 8) * 128 + (r % 8) * 16 + (k % 8) * 2; }

// nmma: 1 = only the first K=16 slice, 2 = both; dfmt: 0 = f16 accumulator, 1 = f32
__global__ void __launch_bounds__(128, 1) probe(const uint8_t *a_img, const uint8_t *b_img, uint32_t *out, int nmma, int dfmt)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 128 * K * 2 / 16; i += 128) {
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(a_img)[i];
        reinterpret_cast<uint4 *>(smem + 128 * K * 2)[i] = reinterpret_cast<const uint4 *>(b_img)[i];
    }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&slot, 128); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = ((uint32_t)dfmt << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = sa + 128 * K * 2;
        for (int j = 0; j < nmma; ++j) {
            const uint64_t ad = umma_desc_kmajor_noswz(sa + j * 256, 128, (K / 8) * 128);
            const uint64_t bd = umma_desc_kmajor_noswz(sb + j * 256, 128, (K / 8) * 128);
            umma_f16_ss(tmem, ad, bd, idesc, j > 0 ? 1u : 0u);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
    const int row = warp * 32 + lane;
    if (dfmt == 0) {
        for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\ttcgen05.wait::ld.sync.aligned;"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr + h * 64) : "memory");
            for (int i = 0; i < 32; ++i) out[row * 128 + h * 32 + i] = r[i];        // packed: 64 words per row used
        }
        // also the raw 32-bit cells of the first 32 columns (what sits in the upper halves?)
        float v[32];
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait_32(v);
        for (int i = 0; i < 32; ++i) out[row * 128 + 64 + i] = __float_as_uint(v[i]);
    } else {
        for (int h = 0; h < 4; ++h) {
            float v[32];
            tmem_ld_32x32(taddr + h * 32, v);
            tmem_ld_wait_32(v);
            for (int i = 0; i < 32; ++i) out[row * 128 + h * 32 + i] = __float_as_uint(v[i]);
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 128); }
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
static uint16_t f2h_rn(double v) { __half x = __double2half(v); uint16_t h; memcpy(&h, &x, 2); return h; }   // RN from double
static uint16_t f2h_rz(double v)
{
    uint16_t h = f2h_rn(v);
    double back = h2f(h);
    if (fabs(back) > fabs(v)) h -= 1;   // step toward zero (same sign, magnitude bits)
    return h;
}

struct Case { std::vector<float> A, B; };   // fp16-representable values, [128][32]

static void run(const Case &c, int nmma, int dfmt, std::vector<uint32_t> &out)
{
    std::vector<uint8_t> ai(128 * K * 2), bi(128 * K * 2);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < K; ++k) {
            uint16_t ha = f2h_rn(c.A[r * K + k]), hb = f2h_rn(c.B[r * K + k]);
            memcpy(&ai[img_off(r, k)], &ha, 2);
            memcpy(&bi[img_off(r, k)], &hb, 2);
        }
    uint8_t *da, *db; uint32_t *dout;
    cudaMalloc(&da, ai.size()); cudaMalloc(&db, bi.size()); cudaMalloc(&dout, 128 * 128 * 4);
    cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice);
    cudaMemset(dout, 0xff, 128 * 128 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * K * 2);
    probe<<<1, 128, 2 * 128 * K * 2>>>(da, db, dout, nmma, dfmt);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    out.resize(128 * 128);
    cudaMemcpy(out.data(), dout, 128 * 128 * 4, cudaMemcpyDeviceToHost);
    cudaFree(da); cudaFree(db); cudaFree(dout);
}

int main()
{
    srand(1);
    auto rnd = []() { return (double)rand() / RAND_MAX * 2.0 - 1.0; };
    auto q = [](double v) { return (float)h2f(f2h_rn(v)); };
    std::vector<uint32_t> out;
    // ---- 1. layout: D[r,n] = n + r/128 style markers ----
    {
        Case c; c.A.assign(128 * K, 0.f); c.B.assign(128 * K, 0.f);
        for (int r = 0; r < 128; ++r) { c.A[r * K] = 1.f; c.B[r * K] = (float)r; }   // D[r,n] = n
        run(c, 2, 0, out);
        printf("layout row 5 packed words 0..7 (lo,hi):");
        for (int i = 0; i < 8; ++i) printf(" (%g,%g)", h2f(out[5 * 128 + i] & 0xffff), h2f(out[5 * 128 + i] >> 16));
        printf("\nlayout row 5 packed words 32..35 (second x32 load at +64 cols):");
        for (int i = 32; i < 36; ++i) printf(" (%g,%g)", h2f(out[5 * 128 + i] & 0xffff), h2f(out[5 * 128 + i] >> 16));
        printf("\nraw 32-bit cells of columns 0..5: ");
        for (int i = 0; i < 6; ++i) printf(" %08x", out[5 * 128 + 64 + i]);
        int bad = 0;
        for (int r = 0; r < 128; ++r)
            for (int i = 0; i < 64; ++i) {
                const uint32_t w = out[r * 128 + i];
                if (h2f(w & 0xffff) != (float)(2 * i) || h2f(w >> 16) != (float)(2 * i + 1)) ++bad;
            }
        printf("\nlayout: word i = (col 2i, col 2i+1) mismatches: %d of %d\n", bad, 128 * 64);
    }
    // ---- 2/3. rounding hypotheses on random data ----
    for (int scen = 0; scen < 4; ++scen) {
        Case c; c.A.resize(128 * K); c.B.resize(128 * K);
        for (int i = 0; i < 128 * K; ++i) {
            double a = rnd(), b = rnd();
            if (scen == 1) { a *= 30; b *= 8; }                       // larger magnitudes (sums ~ 500)
            if (scen == 2 && (i % K) < 2) { a = (i % K) ? -a * 200 : a * 200; b = 5.0 + 0.001 * rnd(); }  // two big cancelling-ish addends in slice 0
            if (scen == 3) { if ((i % K) >= 16) a *= 100; }          // second slice much bigger than the first
            c.A[i] = q(a); c.B[i] = q(b);
        }
        for (int nmma = 1; nmma <= 2; ++nmma) {
            run(c, nmma, 0, out);
            std::vector<uint32_t> out32;
            run(c, nmma, 1, out32);
            long n = 0, m_rn = 0, m_rz = 0, m_rn2 = 0, m_rz2 = 0, m_f32rn = 0, m_f32rz = 0;
            double worst_rel_res = 0, worst_rel_addend = 0;
            for (int r = 0; r < 128; ++r)
                for (int col = 0; col < 128; ++col) {
                    const uint32_t w = out[r * 128 + col / 2];
                    const uint16_t got = (col & 1) ? (uint16_t)(w >> 16) : (uint16_t)(w & 0xffff);
                    double p1 = 0, p2 = 0, big = 0;
                    for (int k = 0; k < 16 * nmma; ++k) {
                        const double pr = (double)c.A[r * K + k] * (double)c.B[col * K + k];
                        if (k < 16) p1 += pr; else p2 += pr;
                        big = fmax(big, fabs(pr));
                    }
                    const double ex = p1 + p2;
                    float f32v; memcpy(&f32v, &out32[r * 128 + col], 4);
                    ++n;
                    m_rn += got == f2h_rn(ex);
                    m_rz += got == f2h_rz(ex);
                    m_rn2 += got == f2h_rn((double)h2f(f2h_rn(p1)) + p2);
                    m_rz2 += got == f2h_rz((double)h2f(f2h_rz(p1)) + p2);
                    m_f32rn += got == f2h_rn((double)f32v);
                    m_f32rz += got == f2h_rz((double)f32v);
                    const double err = fabs((double)h2f(got) - ex);
                    if (fabs(ex) > 0) worst_rel_res = fmax(worst_rel_res, err / fmax(fabs(ex), 1e-30));
                    worst_rel_addend = fmax(worst_rel_addend, err / fmax(big, 1e-30));
                }
            printf("scenario %d, %d MMA(s): n=%ld  ==RN(exact) %ld  ==RZ(exact) %ld  ==RN(RN(p1)+p2) %ld  ==RZ(RZ(p1)+p2) %ld  ==RN(f32 result) %ld  ==RZ(f32 result) %ld | max err/|result| = %.3g (2^%.1f), max err/max|addend| = %.3g (2^%.1f)\n",
                   scen, nmma, n, m_rn, m_rz, m_rn2, m_rz2, m_f32rn, m_f32rz, worst_rel_res, log2(worst_rel_res), worst_rel_addend, log2(worst_rel_addend));
        }
    }
    return 0;
}
