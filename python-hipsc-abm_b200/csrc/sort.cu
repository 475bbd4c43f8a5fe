// sort.cu -- spatial binning as a counting sort (histogram -> scan -> scatter).
// Replaces Simulation.assign_bins / assign_bins_jit (reference simulation.py:190-223, backend.py:31-50):
// instead of dense fixed-capacity buckets (218 MB at 500k cells) the cells are permuted into bin order
// and a (nbins+1) offset table indexes them.
//
// Two placements:
//   fast (default)      : every cell claims the next free slot of its bin with one atomicAdd on a copy of the
//                         offset table.  The order inside a bin is whatever the atomics produced; every
//                         consumer is order-independent (rows are ordered by cell id, force sums are integer
//                         sums), so results stay run-to-run identical.
//   ABM_SORT_STABLE     : ascending cell id inside a bin -- the reference's serial fill order, needed only to
//                         hand out the reference-shaped dense buckets of assign_bins().
// ABM_SORT_PREBINNED skips the histogram: bin_start already holds the bin populations and the scratch buffer
// the bin of every cell, written by the fused mechanics kernel of the previous substep.
#include "sortws.cuh"

#ifndef ABM_JKR_EXP
#define ABM_JKR_EXP 0      // build-time experiment switches of csrc/jkr.cu (bit 16 needs the sort's help)
#endif

namespace abm {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_bin_count(const double *__restrict__ loc, int n, BinGeom g,
                                                        int *__restrict__ counts, int *__restrict__ cell_bin,
                                                        const int *__restrict__ n_dev) {
    int i = blockIdx.x * kThreads + threadIdx.x;
    n = device_count(n, n_dev);
    if (i >= n) return;
    int b = cell_bin_of(loc[3 * (int64_t)i + 0], loc[3 * (int64_t)i + 1], loc[3 * (int64_t)i + 2], g);
    cell_bin[i] = b;
    atomicAdd(&counts[b], 1);
}

// fast placement: coalesced reads by cell id, one 32-byte row + two ints scattered per cell
__global__ void __launch_bounds__(kThreads) k_scatter(const int *__restrict__ cell_bin, int *__restrict__ cursor,
                                                      const double *__restrict__ loc, const double *__restrict__ radii,
                                                      int n, int *__restrict__ sorted_id,
                                                      double4 *__restrict__ sorted_pos, int *__restrict__ sorted_bin,
                                                      int *__restrict__ next_sums, int n_sums,
                                                      const int *__restrict__ gid_of, int *__restrict__ slot_gid,
                                                      const int *__restrict__ n_dev) {
    int i = blockIdx.x * kThreads + threadIdx.x;
    for (int k = i; k < n_sums; k += gridDim.x * kThreads) next_sums[k] = 0;
    n = device_count(n, n_dev);
    if (i >= n) return;
    int b = cell_bin[i];
    int dest = atomicAdd(&cursor[b], 1);
    if (gid_of) slot_gid[dest] = gid_of[i];
    double4 p;
    p.x = loc[3 * (int64_t)i + 0];
    p.y = loc[3 * (int64_t)i + 1];
    p.z = loc[3 * (int64_t)i + 2];
    p.w = radii[i];
    sorted_id[dest] = i;
    sorted_pos[dest] = p;
    sorted_bin[dest] = b;
}

// ABM_SORT_PREBINNED placement: the moved cells arrive in the PREVIOUS sort's slot order (written coalesced by the
// fused mechanics kernel).  Cells move a fraction of a bin per substep, so the destination slots of a warp lie
// within a few cache lines of each other: both sides of the permutation are (nearly) coalesced.
__global__ void __launch_bounds__(kThreads) k_rescatter(const int *__restrict__ next_bin, const int *__restrict__ prev_id,
                                                        const double4 *__restrict__ next_pos, int *__restrict__ cursor,
                                                        int n, int *__restrict__ sorted_id,
                                                        double4 *__restrict__ sorted_pos, int *__restrict__ sorted_bin,
                                                        int *__restrict__ next_sums, int n_sums,
                                                        const int *__restrict__ gid_of, int *__restrict__ slot_gid,
                                                        int *__restrict__ prev_slot, const int *__restrict__ n_dev,
                                                        int compact) {
    pdl_launch_dependents();
    int s = blockIdx.x * kThreads + threadIdx.x;
    pdl_wait();
    n = device_count(n, n_dev);
    // consumed by this sort's scan; the next hand-over adds to it
    for (int k = s; k < n_sums; k += gridDim.x * kThreads) next_sums[k] = 0;
    if (s >= n) return;
    int b = next_bin[s];
    int dest = atomicAdd(&cursor[b], 1);
    int id = prev_id[s];
    if (compact)                                               // 16-byte {x, y} records on both sides (kSortCompactXY)
        reinterpret_cast<double2 *>(sorted_pos)[dest] = reinterpret_cast<const double2 *>(next_pos)[s];
    else
        sorted_pos[dest] = next_pos[s];
    sorted_id[dest] = id;
    sorted_bin[dest] = b;
    if (gid_of) slot_gid[dest] = gid_of[id];
    if (prev_slot) prev_slot[dest] = s;
}

__global__ void __launch_bounds__(kThreads) k_place(const int *__restrict__ cell_bin, int *__restrict__ cursor, int n,
                                                    int *__restrict__ tmp_id) {
    int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= n) return;
    tmp_id[atomicAdd(&cursor[cell_bin[i]], 1)] = i;
}

// stable placement: one thread per provisional slot finds its ascending-id position inside the bin segment
__global__ void __launch_bounds__(kThreads) k_gather(const int *__restrict__ tmp_id, const int *__restrict__ cell_bin,
                                                     const int *__restrict__ bin_start,
                                                     const double *__restrict__ loc, const double *__restrict__ radii,
                                                     int n, int *__restrict__ sorted_id,
                                                     double4 *__restrict__ sorted_pos, int *__restrict__ sorted_bin) {
    int s = blockIdx.x * kThreads + threadIdx.x;
    if (s >= n) return;
    int id = tmp_id[s];
    int b = cell_bin[id];
    int b0 = bin_start[b], b1 = bin_start[b + 1];
    int before = 0;
    for (int t = b0; t < b1; ++t) before += (tmp_id[t] < id);
    int dest = b0 + before;
    double4 p;
    p.x = loc[3 * (int64_t)id + 0];
    p.y = loc[3 * (int64_t)id + 1];
    p.z = loc[3 * (int64_t)id + 2];
    p.w = radii[id];
    sorted_id[dest] = id;
    sorted_pos[dest] = p;
    sorted_bin[dest] = b;
}

__global__ void __launch_bounds__(kThreads) k_bins_dense(const double *__restrict__ loc, int n, BinGeom g,
                                                         const int *__restrict__ bin_start,
                                                         const int *__restrict__ sorted_id, long long max_agents,
                                                         long long *__restrict__ bins, long long *__restrict__ bins_help,
                                                         long long *__restrict__ bin_locations) {
    int s = blockIdx.x * kThreads + threadIdx.x;
    if (s >= n) return;
    int id = sorted_id[s];
    double x = loc[3 * (int64_t)id + 0], y = loc[3 * (int64_t)id + 1], z = loc[3 * (int64_t)id + 2];
    // reference bin_locations are not clamped (simulation.py:210)
    long long rx = (long long)exact_floor_div(x, g.distance) + 1;
    long long ry = (long long)exact_floor_div(y, g.distance) + 1;
    long long rz = (long long)exact_floor_div(z, g.distance) + 1;
    bin_locations[3 * (int64_t)id + 0] = rx;
    bin_locations[3 * (int64_t)id + 1] = ry;
    bin_locations[3 * (int64_t)id + 2] = rz;
    int bx = clampi((int)rx, 0, g.dims[0] - 1), by = clampi((int)ry, 0, g.dims[1] - 1),
        bz = clampi((int)rz, 0, g.dims[2] - 1);
    int b = lin_bin(bx, by, bz, g.dims);
    int place = s - bin_start[b];
    int64_t ref_bin = ((int64_t)bx * g.dims[1] + by) * g.dims[2] + bz;   // reference index order x,y,z
    if (place < max_agents) bins[ref_bin * max_agents + place] = id;
    if (place == 0) bins_help[ref_bin] = bin_start[b + 1] - bin_start[b];
}

}  // namespace abm

using namespace abm;

ABM_EXPORT int abm_bin_dims_host(const double size_host[3], double distance, int flags, int32_t dims_host[3]) {
    ABM_REQUIRE(size_host && dims_host, "null pointer");
    ABM_REQUIRE(distance > 0, "distance must be positive");
    BinGeom g;
    int64_t nbins;
    ABM_REQUIRE(make_bin_geom(size_host, nullptr, distance, flags, g, nbins) == 0, "too many bins");
    for (int k = 0; k < 3; ++k) dims_host[k] = g.dims[k];
    return 0;
}

ABM_EXPORT size_t abm_bin_sort_workspace_bytes(int64_t n, int64_t nbins) {
    if (n < 0 || nbins < 0) return 0;
    return sort_ws_bytes(n, nbins);
}

ABM_EXPORT int abm_bin_sort(const double *locations, const double *radii, int64_t n, const double size_host[3],
                            const double *origin_host, double distance, int flags, int32_t *bin_start,
                            int32_t *sorted_id, double *sorted_pos, int32_t *sorted_bin, int32_t *max_count,
                            void *workspace, size_t workspace_bytes, const int32_t *n_device, void *stream) {
    ABM_REQUIRE(max_count, "null pointer");
    return abm::bin_sort_impl(locations, radii, n, size_host, origin_host, distance, flags, bin_start, sorted_id,
                              sorted_pos, sorted_bin, max_count, workspace, workspace_bytes, nullptr, nullptr, stream,
                              n_device);
}

// zero_table: optional second (nbins+1) offset table cleared on the way (the mechanics loop alternates two tables
// and needs the idle one zeroed for the next histogram).  global_id: optional (n) id table; the workspace's
// slot_gid then receives the global id of the cell in every slot (fast placements only).
int abm::bin_sort_impl(const double *locations, const double *radii, int64_t n, const double size_host[3],
                       const double *origin_host, double distance, int flags, int32_t *bin_start,
                       int32_t *sorted_id, double *sorted_pos, int32_t *sorted_bin, int32_t *max_count,
                       void *workspace, size_t workspace_bytes, int32_t *zero_table, const int32_t *global_id,
                       void *stream, const int32_t *n_device) {
    const bool chained = flags & ABM_SORT_CHAINED;
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    ABM_REQUIRE(size_host && bin_start, "null pointer");
    BinGeom g;
    int64_t nbins;
    ABM_REQUIRE(make_bin_geom(size_host, origin_host, distance, flags, g, nbins) == 0,
                "bad geometry (distance <= 0 or too many bins)");
    ABM_REQUIRE(workspace_bytes >= sort_ws_bytes(n, nbins), "workspace too small");
    cudaStream_t st = as_stream(stream);
    const bool prebinned = flags & ABM_SORT_PREBINNED;
    ABM_REQUIRE(prebinned || !(flags & kSortCompactXY), "compact records exist only between the substeps of a loop");
    int rc;
    if (!prebinned) {
        rc = check_cuda(cudaMemsetAsync(bin_start, 0, (size_t)(nbins + 1) * 4, st), "memset bin_start");
        if (rc) return rc;
    }
    if (max_count) {                                           // nullptr: the caller does not want the statistic
        rc = check_cuda(cudaMemsetAsync(max_count, 0, 4, st), "memset max_count");
        if (rc) return rc;
    }
    if (n == 0) {
        if (zero_table) return check_cuda(cudaMemsetAsync(zero_table, 0, (size_t)(nbins + 1) * 4, st), "memset table");
        return 0;
    }
    ABM_REQUIRE(sorted_id && sorted_pos && sorted_bin && workspace, "null pointer");
    ABM_REQUIRE(prebinned || (locations && radii), "null pointer");
    ABM_REQUIRE(!(prebinned && (flags & ABM_SORT_STABLE)), "ABM_SORT_PREBINNED cannot be combined with ABM_SORT_STABLE");
    ABM_REQUIRE(!(n_device && (flags & ABM_SORT_STABLE)), "a device-side cell count cannot be combined with ABM_SORT_STABLE");
    ABM_REQUIRE(((uintptr_t)sorted_pos & 31) == 0, "sorted_pos must be 32-byte aligned");
    SortWs w = sort_ws_carve(workspace, n, nbins);
    int blocks = ceil_div(n, kThreads);
    if (!prebinned) {
        k_bin_count<<<blocks, kThreads, 0, st>>>(locations, (int)n, g, bin_start, w.cell_bin, n_device);
        ABM_CHECK_LAUNCH("k_bin_count");
    }
    const int nb = scan_blocks(nbins + 1);
    if (prebinned) {
        // the mechanics kernel that binned the cells also summed the histogram per scan chunk: one scan launch
        rc = check_cuda(launch_chained(k_scan_apply, nb, kScanThreads, st, chained, bin_start, nbins + 1,
                                       (const int *)w.next_sums, (int *)nullptr, w.cursor, zero_table, max_count),
                        "k_scan_apply");
        if (rc) return rc;
    } else {
        k_scan_reduce<<<nb, kScanThreads, 0, st>>>(bin_start, nbins + 1, w.block_sums, max_count);
        ABM_CHECK_LAUNCH("k_scan_reduce");
        k_scan_apply<<<nb, kScanThreads, 0, st>>>(bin_start, nbins + 1, w.block_sums, nullptr, w.cursor, zero_table,
                                                  nullptr);
        ABM_CHECK_LAUNCH("k_scan_apply");
    }
    double4 *sp = reinterpret_cast<double4 *>(sorted_pos);
    if (prebinned) {
        rc = check_cuda(launch_chained(k_rescatter, blocks, kThreads, st, chained, (const int *)w.cell_bin,
                                       (const int *)w.tmp_id, (const double4 *)w.next_pos, w.cursor, (int)n, sorted_id,
                                       sp, sorted_bin, w.next_sums, nb, global_id, w.slot_gid,
                                       (ABM_JKR_EXP & 16) ? w.prev_slot : (int *)nullptr, n_device,
                                       (flags & kSortCompactXY) ? 1 : 0),
                        "k_rescatter");
        if (rc) return rc;
    } else if (flags & ABM_SORT_STABLE) {
        rc = check_cuda(cudaMemsetAsync(w.next_sums, 0, (size_t)nb * 4, st), "memset next_sums");
        if (rc) return rc;
        k_place<<<blocks, kThreads, 0, st>>>(w.cell_bin, w.cursor, (int)n, w.tmp_id);
        ABM_CHECK_LAUNCH("k_place");
        k_gather<<<blocks, kThreads, 0, st>>>(w.tmp_id, w.cell_bin, bin_start, locations, radii, (int)n, sorted_id, sp,
                                              sorted_bin);
        ABM_CHECK_LAUNCH("k_gather");
    } else {
        k_scatter<<<blocks, kThreads, 0, st>>>(w.cell_bin, w.cursor, locations, radii, (int)n, sorted_id, sp,
                                               sorted_bin, w.next_sums, nb, global_id, w.slot_gid, n_device);
        ABM_CHECK_LAUNCH("k_scatter");
    }
    return 0;
}

ABM_EXPORT int abm_bins_dense(const double *locations, int64_t n, const double size_host[3], double distance,
                              const int32_t *bin_start, const int32_t *sorted_id, int64_t max_agents, int64_t *bins,
                              int64_t *bins_help, int64_t *bin_locations, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff && max_agents >= 0, "bad sizes");
    if (n == 0) return 0;
    ABM_REQUIRE(locations && size_host && bin_start && sorted_id && bins && bins_help && bin_locations,
                "null pointer");
    BinGeom g;
    int64_t nbins;
    ABM_REQUIRE(make_bin_geom(size_host, nullptr, distance, ABM_SORT_REFERENCE, g, nbins) == 0, "bad geometry");
    k_bins_dense<<<ceil_div(n, kThreads), kThreads, 0, as_stream(stream)>>>(
        locations, (int)n, g, bin_start, sorted_id, (long long)max_agents, (long long *)bins,
        (long long *)bins_help, (long long *)bin_locations);
    ABM_CHECK_LAUNCH("k_bins_dense");
    return 0;
}
