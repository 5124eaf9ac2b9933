// Microbenchmark: tcgen05.ld of fp16 accumulators with .pack::16b versus plain 32-bit loads on sm_100a.
// Every warp reads the same 64 TMEM columns of its lane quarter: either one 32x32b.x64 (64 registers)
// or one 32x32b.x32.pack::16b (64 columns -> 32 registers).  Prints cycles per load and columns/cycle/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_pack tmem_ld_pack.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vq-vae-from-gaussian-vae_b200/csrc/gq_ptx.cuh"
using namespace gq::ptx;

template <int PACK, int BATCH>
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned long long *cycles, float *sink)
{
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 7) * 64);
    uint32_t acc = 0;
    __syncthreads();
    const unsigned long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
            if (PACK) {
                uint32_t r[32];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(base) : "memory");
                acc += r[b & 31];
            } else {
                uint32_t r[64];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(base) : "memory");
                acc += r[b & 63];
            }
        }
        tmem_ld_wait();
    }
    const unsigned long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) sink[0] = 1.f;
    __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}

template <int PACK, int BATCH>
void run(int W, unsigned long long *cyc, float *sink)
{
    const int iters = 2000;
    k<PACK, BATCH><<<148, W * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double cols = (double)W * 32 * 64 * BATCH * iters;
    printf("%s warps=%2d batch=%d  %s  cycles/load/warp=%.1f  lane-columns/cycle/SM=%.1f\n", PACK ? "x32.pack::16b" : "x64.b32      ", W, BATCH,
           cudaGetErrorString(e), avg / iters / BATCH, cols / avg);
}

int main()
{
    unsigned long long *cyc; float *sink;
    cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
    for (int W : {1, 4, 8, 16}) {
        run<0, 1>(W, cyc, sink); run<1, 1>(W, cyc, sink);
        run<0, 4>(W, cyc, sink); run<1, 4>(W, cyc, sink);
    }
    return 0;
}
