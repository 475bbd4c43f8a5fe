// search.cu -- fixed-radius neighbour search over the counting-sorted cells, emitting a row CSR.
// Replaces get_neighbors_cpu/gpu + edges[if_edge] + graph.add_edges/simplify
// (reference backend.py:64-160, simulation.py:242-288).
//
// One thread per sorted slot filters the 3x3 (dy,dz) runs of the bin stencil out of a shared-memory tile staged
// by its warp; a hit is kept in a small per-thread column of shared memory.  Row space is reserved with ONE atomicAdd per warp (inclusive shuffle scan of the
// 32 row lengths), so rows are packed, and each thread then writes its own row ascending.
#include "graph.cuh"
#include "scan.cuh"
#include "tile.cuh"

namespace abm {

constexpr int kSearchThreads = 128;
constexpr int kRowCap = 32;   // rows longer than this are re-walked instead of buffered

template <bool kWrite>
__device__ __forceinline__ int walk_stencil(int s, int bx, int by, int bz, const Stencil &st,
                                            const int *__restrict__ bin_start, const int *__restrict__ sorted_id,
                                            const double4 *__restrict__ spos, double4 p, const Radius &rad,
                                            int *list, int list_cap, int *__restrict__ out, int cnt) {
    for (int dz = -1; dz <= 1; ++dz) {
        int z = bz + dz;
        if (z < 0 || z >= st.dims[2]) continue;
        for (int dy = -1; dy <= 1; ++dy) {
            int y = by + dy;
            if (y < 0 || y >= st.dims[1]) continue;
            int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
            int row = (z * st.dims[1] + y) * st.dims[0];
            int lo = bin_start[row + x0], hi = bin_start[row + x1 + 1];
            for (int t = lo; t < hi; ++t) {
                if (t == s) continue;
                double4 q = spos[t];
                double s2 = sumsq_nofma(q.x - p.x, q.y - p.y, q.z - p.z);
                if (rad.within(s2)) {
                    if (kWrite) out[cnt] = sorted_id[t];
                    else if (cnt < list_cap) list[cnt] = sorted_id[t];
                    ++cnt;
                }
            }
        }
    }
    return cnt;
}

// entries of the old row that the stencil did not find (dist > distance): kept by the union
template <bool kWrite>
__device__ __forceinline__ int walk_old_row(int i, double4 p, const GraphView &old, const double *__restrict__ loc,
                                            const Radius &rad, int *list, int list_cap, int *__restrict__ out,
                                            int cnt) {
    long long os = old.start(i);
    int olen = old.row_len[i];
    for (int k = 0; k < olen; ++k) {
        int j = old.cols[os + k];
        double s2 = sumsq_nofma(loc[3 * (long long)j + 0] - p.x, loc[3 * (long long)j + 1] - p.y,
                                loc[3 * (long long)j + 2] - p.z);
        if (!rad.within(s2)) {
            if (kWrite) out[cnt] = j;
            else if (cnt < list_cap) list[cnt] = j;
            ++cnt;
        }
    }
    return cnt;
}

// Candidates are classified in fp32 out of the warp's shared-memory tile (tile.cuh), which also carries their ids;
// only a candidate inside the fp32 error band of the radius is confirmed with the exact fp64 predicate.  Partner
// ids are collected in a per-thread column of shared memory.  A cell with more than kRowCap partners, or whose
// candidate run does not fit the tile, falls back to the exact walk from global memory.
__global__ void __launch_bounds__(kSearchThreads)
k_search(const int *__restrict__ bin_start, const int *__restrict__ sorted_id, const double4 *__restrict__ spos,
         const int *__restrict__ sorted_bin, int n, Stencil st, Radius rad, const double *__restrict__ loc,
         GraphView old, int *__restrict__ row_start, int *__restrict__ row_len, int *__restrict__ cols,
         long long capacity, int *__restrict__ cursor, const int *__restrict__ n_dev) {
    __shared__ int column[kRowCap][kSearchThreads];
    __shared__ float4 tiles[kSearchThreads / 32][kTileRows * kTileCap];
    __shared__ int tile_ids[kSearchThreads / 32][kTileRows * kTileCap];
    const int tid = threadIdx.x;
    const int s = blockIdx.x * kSearchThreads + tid;
    n = device_count(n, n_dev);
    if (s - (tid & 31) >= n) return;                           // the whole warp is past the last cell
    const bool active = s < n;
    float4 *tile = tiles[tid >> 5];
    int *tile_id = tile_ids[tid >> 5];
    int i = 0, bx = 0, by = 0, bz = 0;
    double4 p = make_double4(0, 0, 0, 0);
    if (active) {
        i = sorted_id[s];
        p = spos[s];
        st.decode(sorted_bin[s], bx, by, bz);
    }
    // fp32 classification over the staged stencil: below the error band a candidate is a neighbour, above it it
    // is not, inside it (a few per million) the exact fp64 predicate decides.  The column holds partner ids.
    TileOrigin org;
    org.x = __shfl_sync(kFullWarp, p.x, 0);
    org.y = __shfl_sync(kFullWarp, p.y, 0);
    org.z = __shfl_sync(kFullWarp, p.z, 0);
    const float pfx = (float)(p.x - org.x), pfy = (float)(p.y - org.y), pfz = (float)(p.z - org.z);
    const float band = tile_band(pfx, pfy, pfz, (float)rad.d);
    const float t_in = (float)rad.lo - band, t_out = (float)rad.hi + band;
    int cnt = 0;
    bool slow = false;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = bz + dz;
        if (__all_sync(kFullWarp, !active || z < 0 || z >= st.dims[2])) continue;
        const LayerRuns R = stage_layer(active, bx, by, z, st, bin_start, spos, org, tile, sorted_id, tile_id);
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
            if (!((R.staged >> r) & 1u)) {                     // warp-uniform
                slow |= R.n[r] > 0;
                continue;
            }
            const float4 *run = tile + R.first[r];
            const int *run_id = tile_id + R.first[r];
            const int nr = R.n[r];
            for (int u = 0; u < nr; ++u) {
                const float4 q = run[u];
                const float dx = pfx - q.x, dy = pfy - q.y, dzf = pfz - q.z;
                const float s2f = fmaf(dzf, dzf, fmaf(dy, dy, dx * dx));
                if (s2f > t_out) continue;
                const int t = __float_as_int(q.w);
                if (t == s) continue;
                if (s2f >= t_in) {                             // inside the band: exact predicate
                    const double4 e = spos[t];
                    if (!rad.within(sumsq_nofma(e.x - p.x, e.y - p.y, e.z - p.z))) continue;
                }
                if (cnt < kRowCap) column[cnt][tid] = run_id[u];
                ++cnt;
            }
        }
        __syncwarp();                                         // the tile is overwritten by the next layer
    }
    slow |= cnt > kRowCap;
    if (active && !slow && old.valid()) {                     // union with the old graph (clear=False)
        const long long os = old.start(i);
        const int olen = old.row_len[i];
        for (int k = 0; k < olen; ++k) {
            const int j = old.cols[os + k];
            const double s2 = sumsq_nofma(loc[3 * (long long)j + 0] - p.x, loc[3 * (long long)j + 1] - p.y,
                                          loc[3 * (long long)j + 2] - p.z);
            if (!rad.within(s2)) {
                if (cnt < kRowCap) column[cnt][tid] = j;
                ++cnt;
            }
        }
        slow |= cnt > kRowCap;
    }
    if (!active) cnt = 0;
    if (active && slow) {                                     // count only; the row is written straight to cols below
        cnt = walk_stencil<false>(s, bx, by, bz, st, bin_start, sorted_id, spos, p, rad, nullptr, 0, nullptr, 0);
        if (old.valid()) cnt = walk_old_row<false>(i, p, old, loc, rad, nullptr, 0, nullptr, cnt);
    }
    // warp-aggregated reservation of row space
    int incl = warp_inclusive_scan(cnt);
    int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane_id() == 31 && total > 0) base = atomicAdd(cursor, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    if (!active) return;
    long long start = (long long)base + incl - cnt;
    row_start[i] = (int)start;
    row_len[i] = cnt;
    if (cnt == 0 || start + cnt > capacity) return;   // overflow: the host sees cursor > capacity and retries
    int *out = cols + start;
    if (!slow) {
        for (int a = 1; a < cnt; ++a) {                       // ascending ids
            int v = column[a][tid], b = a - 1;
            while (b >= 0 && column[b][tid] > v) {
                column[b + 1][tid] = column[b][tid];
                --b;
            }
            column[b + 1][tid] = v;
        }
        for (int k = 0; k < cnt; ++k) out[k] = column[k][tid];
    } else {
        int c = walk_stencil<true>(s, bx, by, bz, st, bin_start, sorted_id, spos, p, rad, nullptr, 0, out, 0);
        if (old.valid()) walk_old_row<true>(i, p, old, loc, rad, nullptr, 0, out, c);
        insertion_sort(out, cnt);
    }
}

// ---- (i<j) edge list of a row CSR ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_upper(GraphView g, int n, int *__restrict__ counts) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i > n) return;
    int c = 0;
    if (i < n) {
        long long s = g.start(i);
        int len = g.row_len[i];
        for (int k = 0; k < len; ++k) c += (g.cols[s + k] > i);
    }
    counts[i] = c;   // counts[n] = 0 so the scan leaves the total there
}

__global__ void __launch_bounds__(256) k_fill_upper(GraphView g, int n, const int *__restrict__ offsets,
                                                    long long *__restrict__ edges, long long capacity) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    long long s = g.start(i);
    int len = g.row_len[i];
    long long o = offsets[i];
    for (int k = 0; k < len; ++k) {
        int j = g.cols[s + k];
        if (j > i) {
            if (o < capacity) {
                edges[2 * o + 0] = i;
                edges[2 * o + 1] = j;
            }
            ++o;
        }
    }
}

}  // namespace abm

using namespace abm;

ABM_EXPORT int abm_neighbor_search(const int32_t *bin_start, const int32_t *sorted_id, const double *sorted_pos,
                                   const int32_t *sorted_bin, int64_t n, const int32_t dims_host[3], double distance,
                                   const double *locations, const int32_t *old_row_start, int64_t old_row_stride,
                                   const int32_t *old_row_len, const int32_t *old_cols, int32_t *row_start,
                                   int32_t *row_len, int32_t *cols, int64_t cols_capacity, int32_t *cursor,
                                   const int32_t *n_device, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    ABM_REQUIRE(cursor && dims_host, "null pointer");
    ABM_REQUIRE(distance > 0, "distance must be positive");
    cudaStream_t st = as_stream(stream);
    int rc = check_cuda(cudaMemsetAsync(cursor, 0, 4, st), "memset cursor");
    if (rc || n == 0) return rc;
    ABM_REQUIRE(bin_start && sorted_id && sorted_pos && sorted_bin && row_start && row_len, "null pointer");
    ABM_REQUIRE(cols || cols_capacity == 0, "null cols with non-zero capacity");
    GraphView old{old_row_start, (long long)old_row_stride, old_row_len, old_cols};
    ABM_REQUIRE(!old.valid() || (locations && old_cols), "union needs locations and old_cols");
    Stencil stc;
    stc.set(dims_host);
    k_search<<<ceil_div(n, kSearchThreads), kSearchThreads, 0, st>>>(
        bin_start, sorted_id, reinterpret_cast<const double4 *>(sorted_pos), sorted_bin, (int)n, stc,
        Radius(distance), locations, old, row_start, row_len, cols, (long long)cols_capacity, cursor, n_device);
    ABM_CHECK_LAUNCH("k_search");
    return 0;
}

ABM_EXPORT size_t abm_graph_edge_list_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return ((size_t)(n + 1) * 4 + 255) / 256 * 256 + (size_t)scan_blocks(n + 1) * 4 + 4;
}

ABM_EXPORT int abm_graph_edge_list(const int32_t *row_start, int64_t row_stride, const int32_t *row_len,
                                   const int32_t *cols, int64_t n, int32_t *scan_ws, size_t scan_ws_bytes,
                                   int64_t *edges, int64_t edges_capacity, int32_t *n_edges, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    ABM_REQUIRE(row_len && scan_ws && n_edges, "null pointer");
    ABM_REQUIRE(scan_ws_bytes >= abm_graph_edge_list_workspace_bytes(n), "scan workspace too small");
    cudaStream_t st = as_stream(stream);
    GraphView g{row_start, (long long)row_stride, row_len, cols};
    int *counts = scan_ws;
    int *block_sums = (int *)((char *)scan_ws + ((size_t)(n + 1) * 4 + 255) / 256 * 256);
    k_count_upper<<<ceil_div(n + 1, 256), 256, 0, st>>>(g, (int)n, counts);
    ABM_CHECK_LAUNCH("k_count_upper");
    int rc = exclusive_scan_inplace(counts, n + 1, block_sums, nullptr, nullptr, st);
    if (rc) return rc;
    rc = check_cuda(cudaMemcpyAsync(n_edges, counts + n, 4, cudaMemcpyDeviceToDevice, st), "copy edge total");
    if (rc) return rc;
    if (n > 0 && edges_capacity > 0) {
        ABM_REQUIRE(edges, "null edges");
        k_fill_upper<<<ceil_div(n, 256), 256, 0, st>>>(g, (int)n, counts, (long long *)edges,
                                                       (long long)edges_capacity);
        ABM_CHECK_LAUNCH("k_fill_upper");
    }
    return 0;
}
