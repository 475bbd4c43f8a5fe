// scan.cuh -- in-place exclusive prefix sum over int32 (two launches, no host sync).
//   pass A: every block reduces its chunk -> block_sums[b] (and optionally the chunk max)
//   pass B: every block sums block_sums[0..b) itself, then scans its chunk with that offset
// Chunk = kScanThreads * kScanItems elements; the counts it scans are bin populations or row lengths.
#pragma once
#include "common.cuh"

namespace abm {

constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;      // k_scan_apply moves them as two int4
constexpr int kScanChunk = kScanThreads * kScanItems;

inline int scan_blocks(int64_t n) { return ceil_div(n, kScanChunk); }

__device__ __forceinline__ int block_reduce_sum(int v, int *smem /*>=32 ints*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane_id() == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    int nw = blockDim.x >> 5;
    int t = (threadIdx.x < nw) ? smem[threadIdx.x] : 0;
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) smem[0] = t;
    }
    __syncthreads();
    int r = smem[0];
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const int *__restrict__ data, int64_t n,
                                                              int *__restrict__ block_sums,
                                                              int *__restrict__ max_out) {
    __shared__ int smem[32];
    int64_t base = (int64_t)blockIdx.x * kScanChunk;
    int sum = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        int v = (i < n) ? data[i] : 0;
        sum += v;
        mx = max(mx, v);
    }
    int total = block_reduce_sum(sum, smem);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
    if (max_out) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane_id() == 0 && mx > 0) atomicMax(max_out, mx);
    }
}

// exclusive scan in place; if total_out != nullptr the grand total is written there by the last block, and if
// copy_out != nullptr the scanned values are written there as well (same alignment as data).  Optional riders
// that save separate launches: zero_out (n ints, same alignment as data) is cleared, max_out is raised to the
// largest input value.
static __global__ void __launch_bounds__(kScanThreads) k_scan_apply(int *__restrict__ data, int64_t n,
                                                             const int *__restrict__ block_sums,
                                                             int *__restrict__ total_out,
                                                             int *__restrict__ copy_out,
                                                             int *__restrict__ zero_out = nullptr,
                                                             int *__restrict__ max_out = nullptr) {
    __shared__ int smem[32];
    __shared__ int warp_tot[kScanThreads / 32];
    pdl_launch_dependents();
    pdl_wait();
    // each thread owns kScanItems CONSECUTIVE elements, moved as two 16-byte vectors when the chunk is full
    // (loaded first: their latency overlaps the reduction of the preceding block sums below)
    int64_t first = (int64_t)blockIdx.x * kScanChunk + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int tsum = 0;
    const bool full = first + kScanItems <= n && ((reinterpret_cast<uintptr_t>(data + first) & 15) == 0);
    // offset of this chunk = sum of the preceding block sums
    int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += kScanThreads) part += block_sums[b];
    if (full) {
        int4 a = *reinterpret_cast<const int4 *>(data + first);
        int4 b = *reinterpret_cast<const int4 *>(data + first + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            int64_t i = first + k;
            v[k] = (i < n) ? data[i] : 0;
        }
    }
    int offset = block_reduce_sum(part, smem);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += v[k];
    int incl = warp_inclusive_scan(tsum);
    if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    if (threadIdx.x < 32) {
        int w = (threadIdx.x < kScanThreads / 32) ? warp_tot[threadIdx.x] : 0;
        int wi = warp_inclusive_scan(w);
        if (threadIdx.x < kScanThreads / 32) warp_tot[threadIdx.x] = wi - w;   // exclusive
    }
    __syncthreads();
    int run = offset + warp_tot[threadIdx.x >> 5] + (incl - tsum);
    int o[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        o[k] = run;
        run += v[k];
    }
    if (full) {
        *reinterpret_cast<int4 *>(data + first) = make_int4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<int4 *>(data + first + 4) = make_int4(o[4], o[5], o[6], o[7]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k)
            if (first + k < n) data[first + k] = o[k];
    }
    if (copy_out) {
        if (full && ((reinterpret_cast<uintptr_t>(copy_out + first) & 15) == 0)) {
            *reinterpret_cast<int4 *>(copy_out + first) = make_int4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<int4 *>(copy_out + first + 4) = make_int4(o[4], o[5], o[6], o[7]);
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k)
                if (first + k < n) copy_out[first + k] = o[k];
        }
    }
    if (zero_out) {
        if (full && ((reinterpret_cast<uintptr_t>(zero_out + first) & 15) == 0)) {
            *reinterpret_cast<int4 *>(zero_out + first) = make_int4(0, 0, 0, 0);
            *reinterpret_cast<int4 *>(zero_out + first + 4) = make_int4(0, 0, 0, 0);
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; ++k)
                if (first + k < n) zero_out[first + k] = 0;
        }
    }
    if (max_out) {
        int mx = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) mx = max(mx, v[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane_id() == 0 && mx > 0) atomicMax(max_out, mx);
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == kScanThreads - 1) *total_out = run;
}

// data[0..n) -> exclusive prefix in place.  block_sums: scan_blocks(n) ints of scratch.
inline int exclusive_scan_inplace(int *data, int64_t n, int *block_sums, int *max_out, int *total_out,
                                  cudaStream_t st, int *copy_out = nullptr) {
    int nb = scan_blocks(n);
    if (nb == 0) return 0;
    k_scan_reduce<<<nb, kScanThreads, 0, st>>>(data, n, block_sums, max_out);
    ABM_CHECK_LAUNCH("k_scan_reduce");
    k_scan_apply<<<nb, kScanThreads, 0, st>>>(data, n, block_sums, total_out, copy_out);
    ABM_CHECK_LAUNCH("k_scan_apply");
    return 0;
}

}  // namespace abm
