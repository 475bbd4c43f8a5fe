// jkr32.cu -- the fused mechanics loop in SINGLE precision on bin-relative coordinates (north star: "per-step forces
// and positions within rel 1e-5 in fp32").
//
// Why bin-relative.  With absolute coordinates one fp32 ulp at x = 7000 um is 5e-4 um, i.e. 1e-3 of a typical 0.5 um
// JKR overlap: forces would be off by 1e-3 relative, two orders above the 1e-5 gate.  Here a position is the pair
//     (bin coordinate b, offset o in fp32 relative to the bin's lower corner),   x = origin + (b-1)*edge + o,
// o in [0, edge): one ulp is 1e-6 um whatever the size of the colony.  The separation of two cells in adjacent bins is
// db*edge + (o_i - o_j) with db in {-1,0,1}: overlaps, the contact model, the force and the displacement are all
// evaluated in fp32 with ~1e-6 relative error.  Only the bookkeeping that must be exact stays in fp64: the clamp to
// the box and the re-binning of a moved cell (a dozen fp64 operations per cell and substep, not per pair).
//
// What is NOT promised in this mode: bit-identical bond sets.  A pair whose distance lies within ~1e-6 um of the
// contact threshold (or whose d lies within ~1e-6 of -0.360562) can fall on the other side than in fp64; the tests
// bound the number of such pairs and check that each one sits at a threshold.  The r = 15 neighbour search is
// unaffected (it always runs on fp64 positions with the exact predicate).
//
// With equal radii (the shipped model) the record's fourth slot carries the cell's linear bin instead of the radius, so
// that a partner's position is ONE 16-byte gather (no second gather of sorted_bin).
//
// Same launch structure as the fp64 loop (jkr.cu): per substep one scan, one placement, one fused kernel; the moved
// cells are handed to the next sort in slot order.  Records are 16-byte float4 {ox, oy, oz, radius} instead of
// 32-byte double4: the working set of a substep shrinks from ~124 MB to ~90 MB at 500k cells.
#include <string.h>

#include "graph.cuh"
#include "sortws.cuh"
#include "tile.cuh"

namespace abm {

constexpr int k32Threads = 128;
constexpr int k32MaxHits = 16;
constexpr int k32MinBlocks = 8;

struct Lattice32 {            // geometry of the mechanics lattice, both precisions
    double edge, inv_edge, origin[3], size[3];
    float edgef;
    int dims[3];
};

struct Contact32 {            // contact model constants (reference cell_backend.py:79-98) in fp32
    float k23, pi_adh;        // ((pi*adhesion)/e_hat)^(2/3), pi*adhesion
    float uni_rsum, uni_inv_scale, uni_force_fixed;   // equal radii: rsum, 1/(1e6*overlap_), pi*adh*r_hat*2^76
    float friction_uni, stokes;
};

__device__ __forceinline__ void offset_of(double x, double origin, double edge, double inv_edge, int dim, int &b,
                                          float &o) {
    // the reference's bin (floor_divide(x, edge) + 1, clamped into the lattice) and the offset from its lower corner
    int q = __double2int_rd((x - origin) * inv_edge) + 1;
    q = clampi(q, 0, dim - 1);
    b = q;
    o = (float)(x - fma((double)(q - 1), edge, origin));
}

// first sort of a loop: bin + offsets of every cell (by cell id), histogram
__global__ void __launch_bounds__(256)
k_bin_count32(const double *__restrict__ loc, const double *__restrict__ radii, int n, Lattice32 L,
              int *__restrict__ counts, int *__restrict__ cell_bin, float4 *__restrict__ tmp_off,
              const int *__restrict__ n_dev, int bin_in_w) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    n = device_count(n, n_dev);
    if (i >= n) return;
    int b[3];
    float o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) offset_of(loc[3 * (long long)i + k], L.origin[k], L.edge, L.inv_edge, L.dims[k], b[k], o[k]);
    const int lin = lin_bin(b[0], b[1], b[2], L.dims);
    cell_bin[i] = lin;
    tmp_off[i] = make_float4(o[0], o[1], o[2], bin_in_w ? __int_as_float(lin) : (float)radii[i]);
    atomicAdd(&counts[lin], 1);
}

__global__ void __launch_bounds__(256)
k_scatter32(const int *__restrict__ cell_bin, const float4 *__restrict__ tmp_off, int *__restrict__ cursor, int n,
            int *__restrict__ sorted_id, float4 *__restrict__ spos, int *__restrict__ sorted_bin,
            int *__restrict__ next_sums, int n_sums, const int *__restrict__ gid_of, int *__restrict__ slot_gid,
            const int *__restrict__ n_dev) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    for (int k = i; k < n_sums; k += gridDim.x * 256) next_sums[k] = 0;
    n = device_count(n, n_dev);
    if (i >= n) return;
    const int b = cell_bin[i];
    const int dest = atomicAdd(&cursor[b], 1);
    if (gid_of) slot_gid[dest] = gid_of[i];
    sorted_id[dest] = i;
    spos[dest] = tmp_off[i];
    sorted_bin[dest] = b;
}

__global__ void __launch_bounds__(256)
k_rescatter32(const int *__restrict__ next_bin, const int *__restrict__ prev_id, const float4 *__restrict__ next_pos,
              int *__restrict__ cursor, int n, int *__restrict__ sorted_id, float4 *__restrict__ spos,
              int *__restrict__ sorted_bin, int *__restrict__ next_sums, int n_sums, const int *__restrict__ gid_of,
              int *__restrict__ slot_gid, const int *__restrict__ n_dev) {
    pdl_launch_dependents();
    const int s = blockIdx.x * 256 + threadIdx.x;
    pdl_wait();
    for (int k = s; k < n_sums; k += gridDim.x * 256) next_sums[k] = 0;
    n = device_count(n, n_dev);
    if (s >= n) return;
    const int b = next_bin[s];
    const int dest = atomicAdd(&cursor[b], 1);
    const int id = prev_id[s];
    spos[dest] = next_pos[s];
    sorted_id[dest] = id;
    sorted_bin[dest] = b;
    if (gid_of) slot_gid[dest] = gid_of[id];
}

struct Next32 {
    int *counts, *cell_bin, *prev_id, *chunk_sums;
    float4 *next_pos;
};

// candidates of one z layer into the warp's tile: {x, y, z relative to the warp's first cell, slot}
__device__ __forceinline__ LayerRuns stage_layer32(bool live, int bx, int by, int z, const Stencil &st,
                                                   const int *__restrict__ bin_start, const float4 *__restrict__ spos,
                                                   const int *__restrict__ sorted_bin, int obx, int oby, int obz,
                                                   float oox, float ooy, float ooz, float edgef, float4 *tile,
                                                   bool bin_in_w) {
    LayerRuns R;
    const int lane = threadIdx.x & 31;
    const bool zok = live && z >= 0 && z < st.dims[2];
    const int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
    int L[kTileRows], W[kTileRows];
    R.staged = 0;
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) {
        const int y = by + r - 1;
        const bool ok = zok && y >= 0 && y < st.dims[1];
        int lo = 0, hi = 0;
        if (ok) {
            const int rowb = (z * st.dims[1] + y) * st.dims[0];
            lo = bin_start[rowb + x0];
            hi = bin_start[rowb + x1 + 1];
        }
        R.lo[r] = lo;
        R.n[r] = hi - lo;
        L[r] = __reduce_min_sync(kFullWarp, hi > lo ? lo : 0x7fffffff);
        const int H = __reduce_max_sync(kFullWarp, hi > lo ? hi : 0);
        W[r] = H - L[r];
        if (W[r] > 0 && W[r] <= kTileCap) R.staged |= 1u << r;
        R.first[r] = lo - L[r] + r * kTileCap;
    }
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) {
        if (R.staged & (1u << r)) {
#pragma unroll
            for (int c0 = 0; c0 < kTileCap; c0 += 32) {
                const int c = c0 + lane;
                if (c < W[r]) {
                    const int t = L[r] + c;
                    const float4 q = spos[t];
                    int qx, qy, qz;
                    st.decode(bin_in_w ? __float_as_int(q.w) : sorted_bin[t], qx, qy, qz);
                    tile[r * kTileCap + c] = make_float4(fmaf((float)(qx - obx), edgef, q.x - oox),
                                                         fmaf((float)(qy - oby), edgef, q.y - ooy),
                                                         fmaf((float)(qz - obz), edgef, q.z - ooz), __int_as_float(t));
                }
            }
        }
    }
    __syncwarp();
    return R;
}

template <bool kUniform>
__global__ void __launch_bounds__(k32Threads, k32MinBlocks)
k_jkr_substep32(const int *__restrict__ bin_start, const int *__restrict__ sorted_id,
                const float4 *__restrict__ spos, const int *__restrict__ sorted_bin, int n, Stencil st, Lattice32 L,
                float dist2, float reach2, BandConst bandc, GraphView old, int *__restrict__ new_len,
                int *__restrict__ new_cols, int stride, int *__restrict__ max_len, const double *__restrict__ mot,
                double *__restrict__ loc, Contact32 c, float dt, const int *__restrict__ gid_of,
                const int *__restrict__ slot_pid, Next32 next, int sort_rows, const int *__restrict__ n_dev,
                const double *__restrict__ move_window) {
    __shared__ int column[k32MaxHits][k32Threads];
    __shared__ float4 tiles[k32Threads / 32][kTileRows * kTileCap];
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int s = blockIdx.x * k32Threads + tid;
    pdl_wait();
    n = device_count(n, n_dev);
    if (s - (tid & 31) >= n) return;
    const bool live = s < n;
    float4 *tile = tiles[tid >> 5];
    int i = 0, bx = 0, by = 0, bz = 0;
    float4 p = make_float4(0.f, 0.f, 0.f, 1.f);
    if (live) {
        i = sorted_id[s];
        p = spos[s];
        st.decode(kUniform ? __float_as_int(p.w) : sorted_bin[s], bx, by, bz);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(mot + 3 * (long long)i));
    }
    const float my_radius = kUniform ? 0.5f * c.uni_rsum : p.w;
    // ---- phase 1: filter over the staged stencil, coordinates relative to the warp's first cell ---------------------
    const int obx = __shfl_sync(kFullWarp, bx, 0), oby = __shfl_sync(kFullWarp, by, 0), obz = __shfl_sync(kFullWarp, bz, 0);
    const float oox = __shfl_sync(kFullWarp, p.x, 0), ooy = __shfl_sync(kFullWarp, p.y, 0),
                ooz = __shfl_sync(kFullWarp, p.z, 0);
    const float pfx = fmaf((float)(bx - obx), L.edgef, p.x - oox), pfy = fmaf((float)(by - oby), L.edgef, p.y - ooy),
                pfz = fmaf((float)(bz - obz), L.edgef, p.z - ooz);
    const float t_out = reach2 + bandc.band(pfx, pfy, pfz);
    int nh = 0;
    bool slow = false;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = bz + dz;
        if (__all_sync(kFullWarp, !live || z < 0 || z >= st.dims[2])) continue;
        const LayerRuns R = stage_layer32(live, bx, by, z, st, bin_start, spos, sorted_bin, obx, oby, obz, oox, ooy,
                                          ooz, L.edgef, tile, kUniform);
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
            if (!((R.staged >> r) & 1u)) {
                slow |= R.n[r] > 0;
                continue;
            }
            const float4 *run = tile + R.first[r];
            const int nr = R.n[r];
            for (int u = 0; u < nr; ++u) {
                const float4 q = run[u];
                const float dx = pfx - q.x, dy = pfy - q.y, dzf = pfz - q.z;
                const float s2f = fmaf(dzf, dzf, fmaf(dy, dy, dx * dx));
                const int t = __float_as_int(q.w);
                if (s2f <= t_out && t != s) {
                    if (nh < k32MaxHits) column[nh][tid] = t;
                    ++nh;
                }
            }
        }
        __syncwarp();
    }
    if (!live) return;
    slow |= nh > k32MaxHits;

    int *row = new_cols + (long long)i * stride;
    long long fxi = 0, fyi = 0, fzi = 0;
    int cnt = 0, direct = 0;
    auto partner_id = [&](int t) {
        if (slot_pid) return slot_pid[t];
        const int j = sorted_id[t];
        return gid_of ? gid_of[j] : j;
    };
    // contact model for the partner at separation v = p - q (fp32), |v|^2 = s2
    auto bond = [&](float vx, float vy, float vz, float s2, float rq, int j) {
        float rsum, inv_scale, force_fixed;
        if (kUniform) {
            rsum = c.uni_rsum;
            inv_scale = c.uni_inv_scale;
            force_fixed = c.uni_force_fixed;
        } else {
            const float r_hat = 1.0f / (1e6f * ((1.0f / my_radius) + (1.0f / rq)));
            inv_scale = 1.0f / (1e6f * (c.k23 * cbrtf(r_hat)));
            force_fixed = c.pi_adh * r_hat * 0x1p76f;
            rsum = my_radius + rq;
        }
        const float inv_mag = s2 > 0.0f ? rsqrtf(s2) : 0.0f;
        const float mag = s2 * inv_mag;
        const float d = (rsum - mag) * inv_scale;
        if (d > -0.360562f) {
            if (s2 > 0.0f) {
                const float f = fmaf(fmaf(fmaf(-0.0204f, d, 0.4942f), d, 1.0801f), d, -1.324f);
                const float w = f * force_fixed * inv_mag;            // force / |v| in 2^-76 N per um
                fxi += __float2ll_rn(w * vx);
                fyi += __float2ll_rn(w * vy);
                fzi += __float2ll_rn(w * vz);
            }
            if (cnt < k32MaxHits) {
                column[cnt++][tid] = j;
            } else {
                if (k32MaxHits + direct < stride) row[k32MaxHits + direct] = j;
                ++direct;
            }
        }
    };
    auto separation = [&](int t, const float4 &q, float &vx, float &vy, float &vz) {
        int qx, qy, qz;
        st.decode(kUniform ? __float_as_int(q.w) : sorted_bin[t], qx, qy, qz);
        vx = fmaf((float)(bx - qx), L.edgef, p.x - q.x);
        vy = fmaf((float)(by - qy), L.edgef, p.y - q.y);
        vz = fmaf((float)(bz - qz), L.edgef, p.z - q.z);
    };
    if (!slow) {
        // ---- phase 2: the parked partners (next partner fetched while the current one is evaluated) -----------------
        int t_next = nh > 0 ? column[0][tid] : s;
        float4 q_next = spos[t_next];
        int b_next = kUniform ? 0 : sorted_bin[t_next];
        for (int k = 0; k < nh; ++k) {
            const int t = t_next;
            const float4 q = q_next;
            const int bq = kUniform ? __float_as_int(q.w) : b_next;
            if (k + 1 < nh) {
                t_next = column[k + 1][tid];
                q_next = spos[t_next];
                if (!kUniform) b_next = sorted_bin[t_next];
            }
            int qx, qy, qz;
            st.decode(bq, qx, qy, qz);
            const float vx = fmaf((float)(bx - qx), L.edgef, p.x - q.x), vy = fmaf((float)(by - qy), L.edgef, p.y - q.y),
                        vz = fmaf((float)(bz - qz), L.edgef, p.z - q.z);
            const float s2 = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
            if (s2 > reach2) continue;
            if (s2 <= dist2) {
                bond(vx, vy, vz, s2, q.w, partner_id(t));
            } else if (old.valid()) {                          // thin shell: only pre-existing bonds persist
                const int j = partner_id(t);
                if (old.contains(i, j)) bond(vx, vy, vz, s2, q.w, j);
            }
        }
    } else {
        // ---- slow path: crowded cell or candidate run longer than the tile ---------------------------------------------
        const int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = bz + dz;
            if (z < 0 || z >= st.dims[2]) continue;
            for (int dy = -1; dy <= 1; ++dy) {
                const int y = by + dy;
                if (y < 0 || y >= st.dims[1]) continue;
                const int rowb = (z * st.dims[1] + y) * st.dims[0];
                const int lo = bin_start[rowb + x0], hi = bin_start[rowb + x1 + 1];
                for (int t = lo; t < hi; ++t) {
                    if (t == s) continue;
                    const float4 q = spos[t];
                    float vx, vy, vz;
                    separation(t, q, vx, vy, vz);
                    const float s2 = fmaf(vz, vz, fmaf(vy, vy, vx * vx));
                    if (s2 > reach2) continue;
                    const int j = partner_id(t);
                    if (s2 <= dist2 || (old.valid() && old.contains(i, j))) bond(vx, vy, vz, s2, q.w, j);
                }
            }
        }
    }
    // ---- phase 3: the bond row ---------------------------------------------------------------------------------------
    const int total = cnt + direct;
    const int kept = min(total, stride);
    if (sort_rows && direct == 0) {
        for (int a = 1; a < cnt; ++a) {
            int v = column[a][tid], b = a - 1;
            while (b >= 0 && column[b][tid] > v) {
                column[b + 1][tid] = column[b][tid];
                --b;
            }
            column[b + 1][tid] = v;
        }
    }
    {
        const int m = min(cnt, stride);
        if ((stride & 3) == 0) {
            for (int k = 0; k < m; k += 4) {
                int4 v;
                v.x = column[k][tid];
                v.y = k + 1 < m ? column[k + 1][tid] : -1;
                v.z = k + 2 < m ? column[k + 2][tid] : -1;
                v.w = k + 3 < m ? column[k + 3][tid] : -1;
                *reinterpret_cast<int4 *>(row + k) = v;
            }
        } else {
            for (int k = 0; k < m; ++k) row[k] = column[k][tid];
        }
        if (sort_rows && direct > 0) insertion_sort(row, kept);
    }
    new_len[i] = kept;
    if (total > stride) atomicMax(max_len, total);
    // ---- phase 4: integration (cell_backend.py:155-170): the displacement in fp32, clamp and re-binning in fp64 ------
    const float friction = kUniform ? c.friction_uni : 6.0f * 3.14159265f * c.stokes * (my_radius / 1e6f);
    const float m3[3] = {(float)mot[3 * (long long)i + 0], (float)mot[3 * (long long)i + 1],
                         (float)mot[3 * (long long)i + 2]};
    const float f3[3] = {(float)fxi * 0x1p-76f, (float)fyi * 0x1p-76f, (float)fzi * 0x1p-76f};
    const int b3[3] = {bx, by, bz};
    const float o3[3] = {p.x, p.y, p.z};
    int nb[3];
    float no[3], moved2 = 0.0f;
    double nx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float velocity = (m3[k] + f3[k]) / friction;
        const float disp = dt * (velocity * 1e6f);
        const double x0 = fma((double)(b3[k] - 1), L.edge, L.origin[k]) + (double)o3[k];
        double x = x0 + (double)disp;
        x = x > L.size[k] ? L.size[k] : (x < 0.0 ? 0.0 : x);
        nx[k] = x;
        moved2 += (float)((x - x0) * (x - x0));
        offset_of(x, L.origin[k], L.edge, L.inv_edge, L.dims[k], nb[k], no[k]);
    }
    record_move(max_len, move_window, nx[0], moved2);
    if (loc) {
        loc[3 * (long long)i + 0] = nx[0];
        loc[3 * (long long)i + 1] = nx[1];
        loc[3 * (long long)i + 2] = nx[2];
    }
    if (next.counts) {
        const int b = lin_bin(nb[0], nb[1], nb[2], L.dims);
        next.next_pos[s] = make_float4(no[0], no[1], no[2], kUniform ? __int_as_float(b) : p.w);
        next.cell_bin[s] = b;
        next.prev_id[s] = i;
        atomicAdd(&next.counts[b], 1);
        const int chunk = b / kScanChunk;
        const unsigned peers = __match_any_sync(__activemask(), chunk);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&next.chunk_sums[chunk], __popc(peers));
    }
}

}  // namespace abm

using namespace abm;

// The substep loop of CellMethods.apply_forces in fp32 (see the header); called by abm_mechanics_loop(ABM_LOOP_FP32).
// Same arguments and buffers as the fp64 loop; sorted_pos is used as (n) float4, the workspace's next_pos region as
// two (n) float4 arrays.
int abm::mechanics_loop32(const double *radii, double *locations, const double *motility_forces, int64_t n,
                          const double extent_host[3], const double *origin_host, const double size_host[3],
                          double distance, double reach, int n_substeps, const double *dts_host, int32_t *bin_start_a,
                          int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos, int32_t *sorted_bin,
                          int32_t *max_count, void *sort_workspace, size_t sort_workspace_bytes,
                          const int32_t *old_row_start, int64_t old_row_stride, const int32_t *old_row_len,
                          const int32_t *old_cols, int32_t *len_a, int32_t *cols_a, int32_t *len_b, int32_t *cols_b,
                          int64_t row_stride, int32_t *max_len, const int32_t *global_id, double uniform_radius,
                          double poisson, double youngs, double adhesion_const, double stokes, int flags,
                          const int32_t *n_device, const double *move_window, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff && n_substeps >= 0 && dts_host, "bad sizes");
    if (n == 0 || n_substeps == 0) return 0;
    const bool carry_in = flags & ABM_LOOP_CARRY_IN, carry_out = flags & ABM_LOOP_CARRY_OUT;
    ABM_REQUIRE(radii && locations && motility_forces && extent_host && size_host && bin_start_a && bin_start_b &&
                    bin_start_a != bin_start_b && sorted_id && sorted_pos && sorted_bin && sort_workspace && len_a &&
                    cols_a && len_b && cols_b && max_len,
                "null pointer");
    ABM_REQUIRE(row_stride > 0 && reach >= distance && distance > 0, "bad arguments");
    BinGeom g;
    int64_t nbins;
    ABM_REQUIRE(make_bin_geom(extent_host, origin_host, reach, 0, g, nbins) == 0, "bad geometry");
    ABM_REQUIRE(sort_workspace_bytes >= sort_ws_bytes(n, nbins), "workspace too small");
    const bool chained = flags & ABM_LOOP_CHAINED;
    cudaStream_t st = as_stream(stream);
    Lattice32 L;
    L.edge = reach;
    L.inv_edge = 1.0 / reach;
    L.edgef = (float)reach;
    for (int k = 0; k < 3; ++k) {
        L.origin[k] = g.origin[k];
        L.size[k] = size_host[k];
        L.dims[k] = g.dims[k];
    }
    Stencil stc;
    stc.set(g.dims);
    JkrConst c64 = make_jkr_const(poisson, youngs, adhesion_const, uniform_radius);
    Contact32 c;
    const double kPi = 3.141592653589793;
    c.k23 = (float)c64.k23;
    c.pi_adh = (float)(kPi * adhesion_const);
    c.stokes = (float)stokes;
    c.uni_rsum = (float)(2.0 * uniform_radius);
    c.uni_inv_scale = (float)c64.uni_inv_scale;
    c.uni_force_fixed = (float)(c64.uni_force_scale * 0x1p76);
    c.friction_uni = (float)(((6.0 * kPi) * stokes) * (uniform_radius / 1e6));
    SortWs w = sort_ws_carve(sort_workspace, n, nbins);
    float4 *next_pos = reinterpret_cast<float4 *>(w.next_pos);
    float4 *tmp_off = next_pos + n;                              // second half of the 32-byte-per-cell region
    float4 *spos = reinterpret_cast<float4 *>(sorted_pos);
    const int blocks = ceil_div(n, 256), nb = scan_blocks(nbins + 1);
    const float dist2 = (float)(distance * distance), reach2 = (float)(reach * reach);
    GraphView cur{old_row_start, (long long)old_row_stride, old_row_len, old_cols};
    auto kernel = uniform_radius > 0 ? k_jkr_substep32<true> : k_jkr_substep32<false>;
    for (int k = 0; k < n_substeps; ++k) {
        int32_t *table = (k & 1) ? bin_start_b : bin_start_a, *other = (k & 1) ? bin_start_a : bin_start_b;
        const bool last = k + 1 == n_substeps;
        const bool prebin = !last || carry_out;              // hand the moved cells to the next sort
        int rc;
        if (k == 0 && !carry_in) {
            rc = check_cuda(cudaMemsetAsync(table, 0, (size_t)(nbins + 1) * 4, st), "memset bin_start");
            if (rc) return rc;
            if (max_count) {
                rc = check_cuda(cudaMemsetAsync(max_count, 0, 4, st), "memset max_count");
                if (rc) return rc;
            }
            k_bin_count32<<<blocks, 256, 0, st>>>(locations, radii, (int)n, L, table, w.cell_bin, tmp_off, n_device,
                                                  uniform_radius > 0 ? 1 : 0);
            ABM_CHECK_LAUNCH("k_bin_count32");
            k_scan_reduce<<<nb, kScanThreads, 0, st>>>(table, nbins + 1, w.block_sums, max_count);
            ABM_CHECK_LAUNCH("k_scan_reduce");
            k_scan_apply<<<nb, kScanThreads, 0, st>>>(table, nbins + 1, w.block_sums, nullptr, w.cursor,
                                                      prebin ? other : nullptr, nullptr);
            ABM_CHECK_LAUNCH("k_scan_apply");
            k_scatter32<<<blocks, 256, 0, st>>>(w.cell_bin, tmp_off, w.cursor, (int)n, sorted_id, spos, sorted_bin,
                                                w.next_sums, nb, global_id, w.slot_gid, n_device);
            ABM_CHECK_LAUNCH("k_scatter32");
        } else {
            rc = check_cuda(launch_chained(k_scan_apply, nb, kScanThreads, st, chained, table, nbins + 1,
                                           (const int *)w.next_sums, (int *)nullptr, w.cursor,
                                           prebin ? other : (int *)nullptr, (int *)nullptr),
                            "k_scan_apply");
            if (rc) return rc;
            rc = check_cuda(launch_chained(k_rescatter32, blocks, 256, st, chained, (const int *)w.cell_bin,
                                           (const int *)w.tmp_id, (const float4 *)next_pos, w.cursor, (int)n,
                                           sorted_id, spos, sorted_bin, w.next_sums, nb, global_id, w.slot_gid,
                                           n_device),
                            "k_rescatter32");
            if (rc) return rc;
        }
        Next32 next;
        memset(&next, 0, sizeof(next));
        if (prebin) {
            next.counts = other;
            next.cell_bin = w.cell_bin;
            next.prev_id = w.tmp_id;
            next.chunk_sums = w.next_sums;
            next.next_pos = next_pos;
        }
        int32_t *nl = (k & 1) ? len_b : len_a, *nc = (k & 1) ? cols_b : cols_a;
        rc = check_cuda(launch_chained(kernel, ceil_div(n, k32Threads), k32Threads, st, chained,
                                       (const int *)table, (const int *)sorted_id, (const float4 *)spos,
                                       (const int *)sorted_bin, (int)n, stc, L, dist2, reach2,
                                       BandConst::make(reach * 1.0000001), cur, nl, nc, (int)row_stride, max_len,
                                       motility_forces, last ? locations : (double *)nullptr, c, (float)dts_host[k],
                                       global_id, global_id ? (const int *)w.slot_gid : (const int *)sorted_id, next,
                                       last ? 1 : 0, n_device, move_window),
                        "k_jkr_substep32");
        if (rc) return rc;
        cur = GraphView{nullptr, (long long)row_stride, nl, nc};
    }
    return 0;
}
