// gq_tokens.cu -- what happens to the indices after the quantizer:
//
//   * codebook-usage histogram and entropy, the statistic behind the reference's headline
//     "almost 100 % codebook usage": eval.py:127 (all_hist = zeros(65536)), eval.py:152-154
//     (hist = torch.histogram(indices.float(), range(0, 65536)); all_hist += hist) and cal_ent,
//     eval.py:137-141 (usage = 1 - mean(hist == 0); p = hist / sum(hist);
//     ent = -sum(p * log2(p + 1e-5)));
//   * a packed index stream: the reference carries indices as int64 everywhere
//     (gaussian.py:138,159) although 16 bits hold an index into a 2^16 codebook; pack / unpack
//     and a dequant gather (gaussian.py:162-178) that reads the packed form directly.  Opt-in:
//     int64 stays the default wire format.
//
// All of it is integer / byte work bounded by HBM bandwidth (8 B read per index for the
// histogram, 8 B read + 2 B written per index for the pack).
#include "gq_common.cuh"
#include "gq_internal.h"
#include "gq_ptx.cuh"

#include <math_constants.h>

namespace gq {

// ---- histogram --------------------------------------------------------------------------------
// hist[idx] += 1, accumulating like `all_hist += hist`.  Indices are close to uniform over the
// codebook (that is the point of the method), so plain L2 atomics see almost no contention; a
// shared-memory privatised copy would not even fit (65536 x 8 B).  Out-of-range indices are
// skipped and flagged (torch.histogram silently drops them too).
__global__ void __launch_bounds__(256)
index_histogram_kernel(const int64_t *__restrict__ idx, int64_t n, int64_t N, unsigned long long *__restrict__ hist,
                       int32_t *__restrict__ err_flag, bool vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (!vec) {   // index pointer not 16-byte aligned: one index per thread and trip
        for (; i < n; i += stride) {
            const int64_t v = __ldg(idx + i);
            if (v >= 0 && v < N) atomicAdd(hist + v, 1ull); else if (err_flag) *err_flag = 1;
        }
        return;
    }
    // two 16-byte loads in flight per thread
    for (; 2 * i + 1 < n; i += stride) {
        const longlong2 v = __ldg(reinterpret_cast<const longlong2 *>(idx) + i);
        if (v.x >= 0 && v.x < N) atomicAdd(hist + v.x, 1ull); else if (err_flag) *err_flag = 1;
        if (v.y >= 0 && v.y < N) atomicAdd(hist + v.y, 1ull); else if (err_flag) *err_flag = 1;
    }
    if (2 * i < n && 2 * i + 1 >= n) {   // odd tail element
        const int64_t v = __ldg(idx + 2 * i);
        if (v >= 0 && v < N) atomicAdd(hist + v, 1ull); else if (err_flag) *err_flag = 1;
    }
}

int launch_index_histogram(const int64_t *idx, int64_t n, int64_t N, unsigned long long *hist, int32_t *err_flag,
                           cudaStream_t stream)
{
    if (n == 0) return GQ_OK;
    const int64_t pairs = (n + 1) / 2;
    const int blocks = (int)imin((pairs + 255) / 256, (int64_t)148 * 8);
    const bool vec = (reinterpret_cast<uintptr_t>(idx) & 15u) == 0;
    index_histogram_kernel<<<blocks, 256, 0, stream>>>(idx, n, N, hist, err_flag, vec);
    count_launch();
    return check_launch("index_histogram_kernel");
}

// ---- usage and entropy (eval.py:137-141) --------------------------------------------------------
// One block; pass 1: exact integer total and number of empty bins, pass 2: -sum p log2(p + 1e-5) with
// p = float(hist) / float(total) as the reference's fp32 tensor ops compute it.  Fixed reduction
// order (deterministic).  out[0] = usage = 1 - empty / N, out[1] = entropy in bits.
constexpr int kUsageThreads = 1024;

__global__ void __launch_bounds__(kUsageThreads)
usage_entropy_kernel(const unsigned long long *__restrict__ hist, int64_t N, float *__restrict__ out)
{
    __shared__ unsigned long long s_tot[kUsageThreads / 32];
    __shared__ unsigned int s_zero[kUsageThreads / 32];
    __shared__ float s_ent[kUsageThreads / 32];
    __shared__ float s_total_f;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long tot = 0;
    unsigned int zeros = 0;
    for (int64_t i = threadIdx.x; i < N; i += kUsageThreads) {
        const unsigned long long h = hist[i];
        tot += h;
        zeros += h == 0ull ? 1u : 0u;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, off);
        zeros += __shfl_xor_sync(0xffffffffu, zeros, off);
    }
    if (lane == 0) { s_tot[wid] = tot; s_zero[wid] = zeros; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        unsigned int z = 0;
        for (int w = 0; w < kUsageThreads / 32; ++w) { t += s_tot[w]; z += s_zero[w]; }
        s_total_f = (float)t;
        // usage = 1 - sum(hist == 0) / N, both in fp32 like the reference
        out[0] = 1.0f - (float)z / (float)N;
    }
    __syncthreads();
    const float total = s_total_f;
    float ent = 0.f;
    for (int64_t i = threadIdx.x; i < N; i += kUsageThreads) {
        const float p = (float)hist[i] / total;
        ent += p * log2f(p + 1e-5f);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ent += __shfl_xor_sync(0xffffffffu, ent, off);
    if (lane == 0) s_ent[wid] = ent;
    __syncthreads();
    if (threadIdx.x == 0) {
        float e = 0.f;
        for (int w = 0; w < kUsageThreads / 32; ++w) e += s_ent[w];
        out[1] = -e;
    }
}

int launch_usage_entropy(const unsigned long long *hist, int64_t N, float *out, cudaStream_t stream)
{
    usage_entropy_kernel<<<1, kUsageThreads, 0, stream>>>(hist, N, out);
    count_launch();
    return check_launch("usage_entropy_kernel");
}

// ---- packed index stream ---------------------------------------------------------------------
// Width: 2 bytes when N <= 65536, else 4.  Eight indices per thread for the 16-bit form (4 x 16-byte
// loads, one 16-byte store), four for the 32-bit form.
template <typename P>
__global__ void __launch_bounds__(256)
pack_indices_kernel(const int64_t *__restrict__ idx, int64_t n, int64_t N, P *__restrict__ packed,
                    int32_t *__restrict__ err_flag, bool vec)
{
    constexpr int V = 16 / (int)sizeof(P);   // indices per 16-byte store
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t groups = vec ? n / V : 0;   // unaligned pointers: everything goes through the scalar loop
    for (int64_t gI = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gI < groups; gI += stride) {
        int64_t v[V];
#pragma unroll
        for (int k = 0; k < V / 2; ++k) {
            const longlong2 t = __ldg(reinterpret_cast<const longlong2 *>(idx + gI * V) + k);
            v[2 * k] = t.x;
            v[2 * k + 1] = t.y;
        }
        __align__(16) P o[V];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            bad |= v[k] < 0 || v[k] >= N;
            o[k] = (P)v[k];
        }
        if (bad && err_flag) *err_flag = 1;
        *reinterpret_cast<uint4 *>(packed + gI * V) = *reinterpret_cast<const uint4 *>(o);
    }
    // tail (fewer than V indices), or the whole stream when the pointers are not 16-byte aligned
    for (int64_t i = groups * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t v = idx[i];
        if ((v < 0 || v >= N) && err_flag) *err_flag = 1;
        packed[i] = (P)v;
    }
}

// Unpack: one 16-byte store (two int64) per thread and trip, so a warp writes 512 contiguous bytes.
template <typename P>
__global__ void __launch_bounds__(256)
unpack_indices_kernel(const P *__restrict__ packed, int64_t n, int64_t *__restrict__ idx, bool vec)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t pairs = vec ? n / 2 : 0;   // unaligned pointers: everything goes through the scalar loop
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
        __align__(8) P two[2];
        if constexpr (sizeof(P) == 2) *reinterpret_cast<uint32_t *>(two) = __ldg(reinterpret_cast<const uint32_t *>(packed) + i);
        else *reinterpret_cast<uint2 *>(two) = __ldg(reinterpret_cast<const uint2 *>(packed) + i);
        reinterpret_cast<longlong2 *>(idx)[i] = make_longlong2((long long)two[0], (long long)two[1]);
    }
    for (int64_t i = pairs * 2 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) idx[i] = (int64_t)packed[i];
}

// ---- index exchange over NVLink peer memory ------------------------------------------------------------
// A sharded caller that wants the whole batch's indices on every rank (eval.py:167-201 does it with an all_gather)
// needs, per rank, "pack my shard" followed by "deliver it to everyone".  One kernel does both: each thread converts
// 16 output bytes worth of int64 indices and stores them at the shard's offset into EVERY rank's gather buffer --
// through peer-mapped pointers (NVLink P2P stores), or with ONE multimem.st to an NVSwitch multicast address, which
// the switch replicates to all ranks.  The caller orders ranks around it (torch symmetric memory's barrier).
__device__ __forceinline__ void multimem_st_16(void *mc_addr, const uint4 &v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename P>
__global__ void __launch_bounds__(256)
peer_gather_kernel(const int64_t *__restrict__ idx, int64_t n, int64_t N, void *const *__restrict__ peer_bufs, void *mc_ptr,
                   int world, int64_t elem_offset, int32_t *__restrict__ err_flag, bool vec, int rank, int64_t flag_off, uint32_t seq)
{
    constexpr int V = 16 / (int)sizeof(P);   // indices per 16-byte store
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t groups = vec ? n / V : 0;
    for (int64_t gI = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; gI < groups; gI += stride) {
        int64_t v[V];
#pragma unroll
        for (int k = 0; k < V / 2; ++k) {
            const longlong2 t = __ldg(reinterpret_cast<const longlong2 *>(idx + gI * V) + k);
            v[2 * k] = t.x;
            v[2 * k + 1] = t.y;
        }
        __align__(16) P o[V];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            bad |= v[k] < 0 || v[k] >= N;
            o[k] = (P)v[k];
        }
        if (bad && err_flag) *err_flag = 1;
        const uint4 q = *reinterpret_cast<const uint4 *>(o);
        const int64_t at = elem_offset + gI * V;
        if (mc_ptr) multimem_st_16(reinterpret_cast<P *>(mc_ptr) + at, q);
        else
            for (int p = 0; p < world; ++p) *reinterpret_cast<uint4 *>(reinterpret_cast<P *>(peer_bufs[p]) + at) = q;
    }
    // tail (fewer than V indices), or the whole shard when something is not 16-byte aligned: element-wise peer stores
    for (int64_t i = groups * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t v = idx[i];
        if ((v < 0 || v >= N) && err_flag) *err_flag = 1;
        for (int p = 0; p < world; ++p) reinterpret_cast<P *>(peer_bufs[p])[elem_offset + i] = (P)v;
    }
    if (seq == 0u) return;   // the caller synchronises the ranks itself
    // ---- cross-rank barrier inside the kernel: the LAST block of this grid announces "my shard is in your buffer" to
    // every rank and waits for the same announcement from every rank; whatever follows on the stream may read the
    // gathered tensor.  Flag block (in the rank's own peer-mapped buffer): 64 slots, then this grid's arrival counter.
    __shared__ bool s_last;
    __threadfence_system();
    __syncthreads();
    uint32_t *arrivals = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(peer_bufs[rank]) + flag_off) + 64;
    if (threadIdx.x == 0) s_last = atomicAdd(arrivals, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < world && !ptx::peer_signal_and_wait(peer_bufs, flag_off, world, rank, (int)threadIdx.x, seq) && err_flag)
        *err_flag = 2;
    __syncthreads();
    if (threadIdx.x == 0) *arrivals = 0u;   // for the next launch (stream-ordered)
}

int launch_peer_gather(const int64_t *idx, int64_t n, int64_t N, int width, void *const *peer_bufs, void *mc_ptr, int world,
                       int64_t elem_offset, int32_t *err_flag, cudaStream_t stream, int rank, int64_t flag_off, uint32_t seq)
{
    if (n == 0 && seq == 0u) return GQ_OK;
    const bool vec = (reinterpret_cast<uintptr_t>(idx) & 15u) == 0 && ((elem_offset * width) & 15) == 0;
    const int per = 16 / width;
    const int blocks = (int)imax(1, imin(((vec ? n / per : n) + 255) / 256, (int64_t)148 * 4));
    if (width == 2)
        peer_gather_kernel<uint16_t><<<blocks, 256, 0, stream>>>(idx, n, N, peer_bufs, mc_ptr, world, elem_offset, err_flag, vec, rank, flag_off, seq);
    else if (width == 4)
        peer_gather_kernel<uint32_t><<<blocks, 256, 0, stream>>>(idx, n, N, peer_bufs, mc_ptr, world, elem_offset, err_flag, vec, rank, flag_off, seq);
    else
        peer_gather_kernel<int64_t><<<blocks, 256, 0, stream>>>(idx, n, N, peer_bufs, mc_ptr, world, elem_offset, err_flag, vec, rank, flag_off, seq);
    count_launch();
    return check_launch("peer_gather_kernel");
}

int index_width(int64_t N) { return N <= 65536 ? 2 : 4; }

int launch_pack_indices(const int64_t *idx, int64_t n, int64_t N, void *packed, int32_t *err_flag, cudaStream_t stream)
{
    if (n == 0) return GQ_OK;
    const bool vec = ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(packed)) & 15u) == 0;
    if (index_width(N) == 2) {
        const int blocks = (int)imax(1, imin(((vec ? n / 8 : n) + 255) / 256, (int64_t)148 * 8));
        pack_indices_kernel<uint16_t><<<blocks, 256, 0, stream>>>(idx, n, N, reinterpret_cast<uint16_t *>(packed), err_flag, vec);
    } else {
        const int blocks = (int)imax(1, imin(((vec ? n / 4 : n) + 255) / 256, (int64_t)148 * 8));
        pack_indices_kernel<uint32_t><<<blocks, 256, 0, stream>>>(idx, n, N, reinterpret_cast<uint32_t *>(packed), err_flag, vec);
    }
    count_launch();
    return check_launch("pack_indices_kernel");
}

int launch_unpack_indices(const void *packed, int64_t n, int64_t N, int64_t *idx, cudaStream_t stream)
{
    if (n == 0) return GQ_OK;
    const bool vec = ((reinterpret_cast<uintptr_t>(idx) | reinterpret_cast<uintptr_t>(packed)) & 15u) == 0;
    if (index_width(N) == 2) {
        const int blocks = (int)imax(1, imin(((vec ? n / 2 : n) + 255) / 256, (int64_t)148 * 16));
        unpack_indices_kernel<uint16_t><<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint16_t *>(packed), n, idx, vec);
    } else {
        const int blocks = (int)imax(1, imin(((vec ? n / 2 : n) + 255) / 256, (int64_t)148 * 16));
        unpack_indices_kernel<uint32_t><<<blocks, 256, 0, stream>>>(reinterpret_cast<const uint32_t *>(packed), n, idx, vec);
    }
    count_launch();
    return check_launch("unpack_indices_kernel");
}

}  // namespace gq
