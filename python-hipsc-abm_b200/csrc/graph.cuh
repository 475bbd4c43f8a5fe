// graph.cuh -- the device bond/neighbour graph ("row CSR"): cell i owns cols[start(i) .. start(i)+len(i)).
// Two placements share the view: packed rows with an explicit row_start (what the search emits through
// warp-aggregated reservation) and fixed-stride rows (row_start == nullptr, start(i) = i*stride; what the
// fused mechanics substep emits).  It replaces the igraph store of reference backend.py:13-28.
#pragma once
#include "common.cuh"

namespace abm {

struct GraphView {
    const int *row_start;   // may be nullptr -> fixed stride
    long long stride;
    const int *row_len;
    const int *cols;
    __device__ __forceinline__ long long start(int i) const {
        return row_start ? (long long)row_start[i] : (long long)i * stride;
    }
    __device__ __forceinline__ bool contains(int i, int j) const {
        long long s = start(i);
        int len = row_len[i];
        for (int k = 0; k < len; ++k)
            if (cols[s + k] == j) return true;
        return false;
    }
    __host__ __device__ bool valid() const { return row_len != nullptr; }
};

// Stencil iteration: for the cell in bin (bx,by,bz) visit, for every (dz,dy), the contiguous run of
// sorted slots covering bins bx-1..bx+1 (x is the fastest bin axis, so the three bins are adjacent).
struct Stencil {
    int dims[3];
    // b / dims[k] as a multiply-high: q = umulhi(b, magic) with magic = floor(2^32/d)+1 is exact while b*d < 2^32
    // (set() checks that for the largest bin id; magic == 0 falls back to the division)
    unsigned magic0, magic1;
    __host__ void set(const int d[3]) {
        for (int k = 0; k < 3; ++k) dims[k] = d[k];
        unsigned long long nbins = (unsigned long long)d[0] * d[1] * d[2];
        unsigned long long rows = (unsigned long long)d[1] * d[2];
        magic0 = (d[0] > 1 && nbins * d[0] < (1ull << 32)) ? (unsigned)((1ull << 32) / d[0]) + 1u : 0u;
        magic1 = (d[1] > 1 && rows * d[1] < (1ull << 32)) ? (unsigned)((1ull << 32) / d[1]) + 1u : 0u;
    }
    __device__ __forceinline__ void decode(int b, int &bx, int &by, int &bz) const {
        int t = magic0 ? (int)__umulhi((unsigned)b, magic0) : b / dims[0];
        bx = b - t * dims[0];
        bz = magic1 ? (int)__umulhi((unsigned)t, magic1) : t / dims[1];
        by = t - bz * dims[1];
    }
};

}  // namespace abm
