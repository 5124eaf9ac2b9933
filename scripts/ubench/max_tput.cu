// Micro-benchmark: how fast can one SM sub-partition reduce 64 fp32 registers to a chunk maximum?
// Variants: 3-input max tree (FMNMX3), 2-input tree (FMNMX), and a split that sends part of the
// compare work to the FMA pipe.  One CTA per SM, W warps per scheduler.  Prints cycles per
// 64-value reduction per warp and per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o max_tput max_tput.cu && ./max_tput
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float max3(float a, float b, float c)
{
    float d;
    asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int MODE>
__device__ __forceinline__ float reduce64(const float (&v)[64], float thr)
{
    if (MODE == 0) {          // FMNMX3 tree
        float a[22];
#pragma unroll
        for (int i = 0; i < 21; ++i) a[i] = max3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        a[21] = v[63];
        float b[8];
#pragma unroll
        for (int i = 0; i < 7; ++i) b[i] = max3(a[3 * i], a[3 * i + 1], a[3 * i + 2]);
        b[7] = a[21];
        float c0 = max3(b[0], b[1], b[2]), c1 = max3(b[3], b[4], b[5]), c2 = fmaxf(b[6], b[7]);
        return max3(c0, c1, c2);
    } else if (MODE == 1) {   // FMNMX 2-input tree
        // 2-input tree, volatile so that nothing is fused or hoisted
        float a[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) asm volatile("max.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(v[2 * i]), "f"(v[2 * i + 1]));
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) asm volatile("max.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(a[i]), "f"(a[i + w]));
        return a[0];
    } else {                  // MODE 2: FMA-pipe count of values above the threshold:
        // s = sum_i sat((v_i - thr) * 2^100)  -> 0 iff no value exceeds thr.  FADD + FFMA.SAT per value
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            s0 += __saturatef((v[i] - thr) * 1.2676506e30f);
            s1 += __saturatef((v[i + 1] - thr) * 1.2676506e30f);
        }
        return s0 + s1;
    }
}

template <int MODE>
__global__ void __launch_bounds__(640, 1) k(const float *__restrict__ in, float *__restrict__ out, long long *cyc, int rounds)
{
    float v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = in[(threadIdx.x * 64 + i) & 4095];
    float acc = 0.f;
    float thr = in[5];
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        const float m = reduce64<MODE>(v, thr);
        acc += m;
        // perturb two registers so that the tree cannot be hoisted out of the loop
        // one changing input per leaf operation (FADDs on the FMA pipe), else the tree is loop invariant
#pragma unroll
        for (int i = 0; i < 64; i += (MODE == 1 ? 2 : 3)) v[i] += m;
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, float *in, float *out, long long *cyc)
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
        printf("%-28s warps/scheduler %d: %7.1f cycles per 64-value reduction per warp, %7.1f per scheduler\n", name, wps,
               (double)c / rounds, (double)c / rounds / wps);
    }
}

int main()
{
    float *in, *out;
    long long *cyc;
    cudaMalloc(&in, 4096 * 4);
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 8);
    float h[4096];
    for (int i = 0; i < 4096; ++i) h[i] = (float)((i * 2654435761u) >> 8) * 1e-6f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("FMNMX3 tree", in, out, cyc);
    run<1>("FMNMX 2-input tree", in, out, cyc);
    run<2>("FADD+FFMA.SAT count (fma)", in, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
