// Micro-benchmark: issue rate of tcgen05.mma kind::f16 (fp16 x fp16, K = 16 per instruction) on sm_100a when
// nothing else runs: one CTA per SM, one thread issues `iters` MMAs back to back and commits once.
//   SS: A and B from shared memory (M = 128, N = 128 or 256)     TS: A from tensor memory, B from shared memory
// Prints cycles per instruction and the implied fraction of the 8192 flop/cycle/SM tensor peak.  If SS at N = 128
// needs more than 64 cycles, the operand fetch from shared memory (A 4 KB + B 4 KB per instruction), not the
// tensor pipe, paces a kernel built from such instructions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../vq-vae-from-gaussian-vae_b200/csrc/gq_ptx.cuh"
using namespace gq::ptx;

// mode 0: SS N=128, 1: SS N=256, 2: TS N=128, 3: TS N=256; acc16: fp16 accumulators; nacc: accumulators used in rotation
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFFu);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout & 7u) << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) k(int mode, int acc16, int nacc, int iters, unsigned long long *cycles, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (65536 + 65536) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;   // fp16 ones
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    const int n = (mode & 1) ? 256 : 128;
    if (mode >= 2) {   // A operand (128 rows x K=16 fp16 = 8 columns) into tensor memory, columns 480..
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) r[i] = 0x3c003c00u;
        tmem_st_32x8(tmem + ((uint32_t)(warp * 32) << 16) + 480u, r);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = ((uint32_t)(acc16 ? 0 : 1) << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = sa + 65536;
        const uint64_t ad = mk_desc(sa, lbo, sbo, layout);
        const uint64_t bd = mk_desc(sb, lbo, sbo, layout);
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (uint32_t)((i % nacc) * n) % 480u;
            if (mode >= 2) umma_f16_ts(d, tmem + 480u, bd, idesc, 1u);
            else umma_f16_ss(d, ad, bd, idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const unsigned long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}


// ---- cta_group::2: one instruction spans a CTA pair (M = 256: 128 rows per CTA; each CTA holds HALF of the N columns of B) ----
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k2(int n, int ts, int iters, unsigned long long *cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_rank();
    for (int i = threadIdx.x; i < (65536 + 65536) / 4; i += 128) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tmem = slot;
    if (ts) {
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) r[i] = 0x3c003c00u;
        tmem_st_32x8(tmem + ((uint32_t)(warp * 32) << 16) + 480u, r);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    if (rank == 0 && warp == 0 && elect_one()) {
        const uint32_t idesc = (0u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
        const uint32_t sa = smem_u32(smem), sb = sa + 65536;
        const uint64_t ad = mk_desc(sa, 128, 256, 0);
        const uint64_t bd = mk_desc(sb, 128, 256, 0);
        const unsigned long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t d = tmem + (uint32_t)((i & 1) * (n <= 240 ? n : 0));
            if (ts)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                             ::"r"(d), "r"(tmem + 480u), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
        mbar_wait(&bar, 0);
        const unsigned long long t1 = clock64();
        cycles[blockIdx.x / 2] = t1 - t0;
    } else if (threadIdx.x == 0 && rank == 1) {
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    tc_fence_before(); cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

// ---- the same instruction stream with sixteen other warps reading tensor memory (tcgen05.ld ... pack::16b) all the time:
// does the TS form keep its rate when the epilogue's loads share tensor memory with its operand fetch? ----
__global__ void __launch_bounds__(544, 1) k3(int ts, int iters, int ld_on, unsigned long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (65536 + 65536) / 4; i += 544) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (warp >= 1 && warp <= 4) {
        uint32_t r[8];
        for (int i = 0; i < 8; ++i) r[i] = 0x3c003c00u;
        tmem_st_32x8(tmem + ((uint32_t)((warp - 1) * 32) << 16) + 480u, r);
        tmem_st_wait();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = (0u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint32_t sa = smem_u32(smem), sb = sa + 65536;
            const uint64_t ad = mk_desc(sa, 128, 256, 0), bd = mk_desc(sb, 128, 256, 0);
            const unsigned long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t d = tmem + (uint32_t)((i & 1) * 128);
                if (ts) umma_f16_ts(d, tmem + 480u, bd, idesc, 1u);
                else umma_f16_ss(d, ad, bd, idesc, 1u);
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const unsigned long long t1 = clock64();
            cycles[blockIdx.x] = t1 - t0;
            stop = 1;
        }
        __syncwarp();
    } else if (ld_on) {
        // 16 reader warps: lane quarter (warp - 1) % 4, 64 packed columns at (warp - 1) / 4 * 64 -- the epilogue's pattern
        const uint32_t taddr = tmem + ((uint32_t)(((warp - 1) & 3) * 32) << 16) + (uint32_t)(((warp - 1) >> 2) * 64);
        uint32_t acc = 0;
        while (!stop) {
            uint32_t v[32];
            tmem_ld_32x64p_sync(taddr, v);
            acc += v[(warp + acc) & 31];
        }
        if (acc == 0x1234567u) sink[0] = acc;
    }
    __syncthreads();
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main()
{
    unsigned long long *cyc;
    cudaMalloc(&cyc, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 65536);
    const int iters = 4000;
    const char *names[] = {"SS M128 N128", "SS M128 N256", "TS M128 N128", "TS M128 N256"};
    struct L { const char *name; uint32_t lbo, sbo, layout; };
    const L layouts[] = {{"interleave lbo128 sbo256 (K=16 image)", 128, 256, 0}, {"interleave lbo128 sbo512 (the K=32 image of the search kernel)", 128, 512, 0},
                         {"interleave lbo2048 sbo128 (core matrices of one K step 2 KB apart)", 2048, 128, 0},
                         {"swizzle32B sbo256", 16, 256, 6}, {"swizzle64B sbo512", 16, 512, 4}, {"swizzle128B sbo1024", 16, 1024, 2}};
    for (const L &l : layouts)
        for (int mode = 0; mode < 4; ++mode) {
            k<<<148, 128, 65536 + 65536>>>(mode, 1, 2, iters, cyc, l.lbo, l.sbo, l.layout);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            const int n = (mode & 1) ? 256 : 128;
            const double per = avg / iters, ideal = 128.0 * n * 16 * 2 / 8192.0;
            printf("%-70s %s  %s  cycles/instruction %.1f (ideal %.0f)  tensor fraction %.2f\n", l.name, names[mode], cudaGetErrorString(e), per, ideal, ideal / per);
        }
    cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 65536);
    for (int ts = 0; ts < 2; ++ts)
        for (int n : {128, 256}) {
            k2<<<148, 128, 65536 + 65536>>>(n, ts, iters, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[74];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 74; ++i) avg += h[i]; avg /= 74;
            const double per = avg / iters, ideal = 128.0 * n * 16 * 2 / 8192.0;   // per CTA: 128 rows x n columns
            printf("cta_group::2 %s M256 (128 per CTA) N%d  %s  cycles/instruction %.1f (ideal %.0f per CTA)  tensor fraction %.2f\n", ts ? "TS" : "SS", n,
                   cudaGetErrorString(e), per, ideal, ideal / per);
        }
    cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 65536);
    uint32_t *sink; cudaMalloc(&sink, 4);
    for (int ts = 0; ts < 2; ++ts)
        for (int ld_on = 0; ld_on < 2; ++ld_on) {
            k3<<<148, 544, 65536 + 65536>>>(ts, iters, ld_on, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            unsigned long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            printf("%s M128 N128 with 16 warps %s tensor memory  %s  cycles/instruction %.1f\n", ts ? "TS" : "SS", ld_on ? "reading (tcgen05.ld x32.pack::16b loop)" : "idle on",
                   cudaGetErrorString(e), avg / iters);
        }
    return 0;
}
