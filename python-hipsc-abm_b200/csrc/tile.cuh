// tile.cuh -- warp-level staging of the bin-stencil candidates in shared memory.
//
// A warp owns 32 consecutive sorted slots.  Because the bins are ordered x fastest, the candidates of all its
// cells for one stencil row (dy,dz) are ONE short contiguous slot run [L,H): the warp loads that run once with
// coalesced 32-byte reads, converts it to float4 {x-ox, y-oy, z-oz, slot} relative to the warp's first cell and
// parks it in shared memory.  Every lane then walks its own sub-run out of shared memory (one LDS.128 per
// candidate, ~30 cycles) with an fp32 distance test instead of chasing L1/L2 gathers (~300+ cycles each) in fp64.
//
// The fp32 test is only a FILTER: a candidate passes when its fp32 squared distance is below the threshold plus an
// error band, and everything that passes is re-evaluated from the fp64 positions with the exact predicate.  The
// filter can let through a few candidates too many (band ~1e-3 um^2: a few per million), never one too few, so
// the result is the exact predicate's for every pair.
//
// Error bound of the fp32 squared distance.  Coordinates relative to the origin have magnitude <= m; each is
// rounded to fp32 once (error <= m 2^-24), the difference once more, so |d_k(fp32) - d_k| <= m 2^-23 =: e.  For a
// pair near a threshold T (sum |d_k| <= sqrt(3 T)):  |s2(fp32) - s2| <= 2 sqrt(3T) e + 3 e^2 + 4 T 2^-24.
// tile_band() returns twice that.
#pragma once
#include "graph.cuh"

namespace abm {

constexpr int kTileCap = 64;          // staged candidates per stencil row and warp (longer runs: exact path)
constexpr int kTileRows = 3;          // the three dy rows of one z layer are staged together
constexpr unsigned kFullWarp = 0xffffffffu;

struct TileOrigin {
    double x, y, z;
};

__device__ __forceinline__ float tile_band(float px, float py, float pz, float sqrt_threshold) {
    float m = fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz)) + 2.0f * sqrt_threshold;
    float e = m * 0x1p-23f;
    float t = sqrt_threshold * sqrt_threshold;
    return 2.0f * (3.5f * sqrt_threshold * e + 3.0f * e * e + t * 0x1p-22f) + 1e-5f;
}

// The same bound with the threshold-dependent factors evaluated once on the host: band = c0 + m (c1 + c2 m),
// m = max |relative coordinate of the lane's own cell| (the + 2 sqrt(T) of tile_band is folded into the constants
// with (m + k)^2 <= 2 m^2 + 2 k^2).
struct BandConst {
    float c0, c1, c2;
    __host__ static BandConst make(double sqrt_threshold) {
        const double k = 2.0 * sqrt_threshold, u = 0x1p-23, t = sqrt_threshold * sqrt_threshold;
        BandConst b;
        b.c1 = (float)(2.0 * 3.5 * sqrt_threshold * u * 1.0001);
        b.c2 = (float)(2.0 * 3.0 * u * u * 2.0 * 1.0001);
        b.c0 = (float)((2.0 * (3.5 * sqrt_threshold * u * k + 3.0 * u * u * 2.0 * k * k + t * 0x1p-22) + 1e-5) * 1.0001);
        return b;
    }
    __device__ __forceinline__ float band(float px, float py, float pz) const {
        float m = fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz));
        return fmaf(m, fmaf(m, c2, c1), c0);
    }
};

// One z layer of a lane's stencil: per dy row, where the lane's run sits in the tile and in the sorted arrays.
struct LayerRuns {
    int first[kTileRows];   // tile index of the lane's first candidate of row r
    int lo[kTileRows];      // global slot of the lane's first candidate of row r
    int n[kTileRows];       // candidates of the lane in row r
    unsigned staged;        // warp-uniform: bit r set if row r sits in the tile
};

// Stage the runs of one z layer.  Every lane of the warp must call it (lanes without a cell pass live=false).
// tile: kTileRows*kTileCap float4 of this warp; entry = {x-ox, y-oy, z-oz, slot as int bits}.
// ids / tile_id (optional): also stage sorted_id of every candidate (tile_id: kTileRows*kTileCap ints of the warp).
__device__ __forceinline__ LayerRuns stage_layer(bool live, int bx, int by, int z, const Stencil &st,
                                                 const int *__restrict__ bin_start,
                                                 const double4 *__restrict__ spos, const TileOrigin &o,
                                                 float4 *tile, const int *__restrict__ ids = nullptr,
                                                 int *tile_id = nullptr) {
    LayerRuns R;
    const int lane = threadIdx.x & 31;
    const bool zok = live && z >= 0 && z < st.dims[2];
    const int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
    int L[kTileRows], W[kTileRows];
    R.staged = 0;
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) {
        int y = by + r - 1;
        bool ok = zok && y >= 0 && y < st.dims[1];
        int lo = 0, hi = 0;
        if (ok) {
            int rowb = (z * st.dims[1] + y) * st.dims[0];
            lo = bin_start[rowb + x0];
            hi = bin_start[rowb + x1 + 1];
        }
        R.lo[r] = lo;
        R.n[r] = hi - lo;
        L[r] = __reduce_min_sync(kFullWarp, hi > lo ? lo : 0x7fffffff);
        int H = __reduce_max_sync(kFullWarp, hi > lo ? hi : 0);
        W[r] = H - L[r];                                    // <= 0: no lane has a candidate in this row
        if (W[r] > 0 && W[r] <= kTileCap) R.staged |= 1u << r;
        R.first[r] = lo - L[r] + r * kTileCap;
    }
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) {
        if (R.staged & (1u << r)) {
#pragma unroll
            for (int c0 = 0; c0 < kTileCap; c0 += 32) {
                int c = c0 + lane;
                if (c < W[r]) {
                    int t = L[r] + c;
                    double4 q = spos[t];
                    tile[r * kTileCap + c] = make_float4((float)(q.x - o.x), (float)(q.y - o.y), (float)(q.z - o.z),
                                                         __int_as_float(t));
                    if (ids) tile_id[r * kTileCap + c] = ids[t];
                }
            }
        }
    }
    __syncwarp();
    return R;
}

}  // namespace abm
