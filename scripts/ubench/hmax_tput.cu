// Micro-benchmark: packed fp16 row-maximum on sm_100a.  How fast can one SM sub-partition reduce the
// 64 scores a thread holds as 32 packed f16x2 registers (tcgen05.ld ... .pack::16b of fp16 accumulators)
// compared with the 64 fp32 registers of the fp32-accumulator epilogue (FMNMX3 tree)?
//   mode 0  FMNMX3 tree over 64 fp32 registers (the round-1 epilogue)
//   mode 1  max.f16x2 chain over 32 packed registers (ptxas fuses pairs into 3-input VHMNMX)
//   mode 2  2-input max.f16x2 tree, fusion blocked by volatile asm per level (HMNMX2)
//   mode 3  __vimax3_s16x2 chain over 32 packed registers (VIMNMX3.S16x2)
// One CTA per SM, W warps per scheduler.  Prints cycles per 64-score reduction per warp and scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmax_tput hmax_tput.cu && ./hmax_tput
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint32_t hmax2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmax2v(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm("add.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <int MODE>
__global__ void __launch_bounds__(640, 1) k(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, long long *cyc, int rounds)
{
    constexpr int NR = MODE == 0 ? 64 : 32;
    uint32_t v[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) v[i] = in[(threadIdx.x * NR + i) & 4095];
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        uint32_t m;
        if (MODE == 0) {
            float a[22];
#pragma unroll
            for (int i = 0; i < 21; ++i) a[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
            a[21] = __uint_as_float(v[63]);
            float b[8];
#pragma unroll
            for (int i = 0; i < 7; ++i) b[i] = max3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
            b[7] = a[21];
            float c0 = max3(b[0], b[1], b[2]), c1 = max3(b[3], b[4], b[5]), c2 = fmaxf(b[6], b[7]);
            m = __float_as_uint(max3(c0, c1, c2));
#pragma unroll
            for (int i = 0; i < 64; i += 3) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(m));
        } else if (MODE == 1) {
            // leaves: 10 x (3 regs) + 1 x (2 regs); then 3-input levels
            uint32_t a[11];
#pragma unroll
            for (int i = 0; i < 10; ++i) a[i] = hmax2(hmax2(v[3 * i], v[3 * i + 1]), v[3 * i + 2]);
            a[10] = hmax2(v[30], v[31]);
            uint32_t b0 = hmax2(hmax2(a[0], a[1]), a[2]), b1 = hmax2(hmax2(a[3], a[4]), a[5]), b2 = hmax2(hmax2(a[6], a[7]), a[8]);
            uint32_t b3 = hmax2(a[9], a[10]);
            m = hmax2(hmax2(hmax2(b0, b1), b2), b3);
#pragma unroll
            for (int i = 0; i < 32; i += 3) v[i] = hadd2(v[i], m);
        } else if (MODE == 2) {
            uint32_t a[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = hmax2v(v[2 * i], v[2 * i + 1]);
#pragma unroll
            for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                for (int i = 0; i < w; ++i) a[i] = hmax2v(a[i], a[i + w]);
            m = a[0];
#pragma unroll
            for (int i = 0; i < 32; i += 2) v[i] = hadd2(v[i], m);
        } else {
            uint32_t a[11];
#pragma unroll
            for (int i = 0; i < 10; ++i) a[i] = __vimax3_s16x2(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
            a[10] = __vmaxs2(v[30], v[31]);
            uint32_t b0 = __vimax3_s16x2(a[0], a[1], a[2]), b1 = __vimax3_s16x2(a[3], a[4], a[5]), b2 = __vimax3_s16x2(a[6], a[7], a[8]);
            uint32_t b3 = __vmaxs2(a[9], a[10]);
            m = __vmaxs2(__vimax3_s16x2(b0, b1, b2), b3);
#pragma unroll
            for (int i = 0; i < 32; i += 3) v[i] = hadd2(v[i], m);
        }
        acc += m;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, uint32_t *in, uint32_t *out, long long *cyc)
{
    for (int wps = 1; wps <= 5; ++wps) {
        const int threads = 128 * wps;
        const int rounds = 2000;
        k<MODE><<<148, threads>>>(in, out, cyc, rounds);
        cudaDeviceSynchronize();
        k<MODE><<<148, threads>>>(in, out, cyc, rounds);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
        printf("%-36s warps/scheduler %d: %7.1f cycles per 64-score reduction per warp, %7.1f per scheduler\n", name, wps,
               (double)c / rounds, (double)c / rounds / wps);
    }
}

int main()
{
    uint32_t *in, *out;
    long long *cyc;
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    uint32_t h[4096];
    for (int i = 0; i < 4096; ++i) {
        __half2 p = __floats2half2_rn((float)((i * 2654435761u) >> 20) * 1e-3f, (float)((i * 40503u) & 1023) * 1e-2f);
        h[i] = *reinterpret_cast<uint32_t *>(&p);
    }
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("FMNMX3 tree, 64 fp32 regs", in, out, cyc);
    run<1>("max.f16x2 chain (VHMNMX), 32 packed", in, out, cyc);
    run<2>("max.f16x2 2-input tree, 32 packed", in, out, cyc);
    run<3>("vimax3_s16x2 (VIMNMX3), 32 packed", in, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
