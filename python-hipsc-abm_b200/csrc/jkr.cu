// jkr.cu -- JKR adhesion/repulsion forces, Stokes integration, and the fused mechanics substep.
// Replaces jkr_forces_cpu/gpu, apply_forces_cpu/gpu and the body of the 11-substep loop of
// CellMethods.apply_forces (reference cell_backend.py:6-172, cell_methods.py:342-439).
#include <string.h>

#include "graph.cuh"
#include "sortws.cuh"

namespace abm {

constexpr int kJkrThreads = 128;

struct Box {
    double size[3];
};

// ---- unfused: forces over existing rows, broken bonds compacted away in place ------------------------
__global__ void __launch_bounds__(kJkrThreads)
k_jkr_rows(const int *__restrict__ row_start, long long stride, int *__restrict__ row_len, int *__restrict__ cols,
           const double *__restrict__ loc, const double *__restrict__ radii, double *__restrict__ forces, int n,
           JkrConst c) {
    int i = blockIdx.x * kJkrThreads + threadIdx.x;
    if (i >= n) return;
    int len = row_len[i];
    if (len == 0) return;
    long long s = row_start ? (long long)row_start[i] : (long long)i * stride;
    double px = loc[3 * (long long)i + 0], py = loc[3 * (long long)i + 1], pz = loc[3 * (long long)i + 2];
    double pr = radii[i];
    // continue the running sum held in jkr_forces so the += order equals the serial reference's
    double fx = forces[3 * (long long)i + 0], fy = forces[3 * (long long)i + 1], fz = forces[3 * (long long)i + 2];
    int kept = 0;
    for (int k = 0; k < len; ++k) {
        int j = cols[s + k];
        double vx = px - loc[3 * (long long)j + 0], vy = py - loc[3 * (long long)j + 1],
               vz = pz - loc[3 * (long long)j + 2];
        double s2 = sumsq_nofma(vx, vy, vz);
        double gx, gy, gz;
        if (jkr_bond(vx, vy, vz, s2, pr, radii[j], c, gx, gy, gz)) {
            fx += gx;
            fy += gy;
            fz += gz;
            cols[s + kept++] = j;
        }
    }
    row_len[i] = kept;
    forces[3 * (long long)i + 0] = fx;
    forces[3 * (long long)i + 1] = fy;
    forces[3 * (long long)i + 2] = fz;
}

__global__ void __launch_bounds__(256)
k_integrate(const double *__restrict__ jkr, const double *__restrict__ mot, double *__restrict__ loc,
            const double *__restrict__ radii, int n, double stokes, Box box, double dt,
            double *__restrict__ zero_me) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    double friction = stokes_friction(radii[i], stokes);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        long long a = 3 * (long long)i + k;
        loc[a] = stokes_move(loc[a], mot[a], jkr[a], friction, dt, box.size[k]);
    }
    if (zero_me) {
#pragma unroll
        for (int k = 0; k < 3; ++k) zero_me[3 * (long long)i + k] = 0.0;
    }
}

// ---- fused: search(distance, union with old bonds) + JKR + bond breaking + integration ---------------
// Phases of one thread (one sorted slot = one cell), arranged so that the warp stays converged while it does
// the expensive work and so that the scattered memory operations are few and wide:
//   1. walk the bin stencil with the cheap test |p-q|^2 <= reach^2; partners within `distance` are parked (their
//      slot) in a per-thread column of shared memory, partners in the thin hysteresis shell (distance, reach]
//      in two registers;
//   2. the (few) lanes that have shell partners start loading their previous bond row -- its latency hides behind
//      the contact loop;
//   3. contact model over the parked partners, all lanes together (trip count = the largest hit count in the
//      warp).  The force is summed in 2^-76 N fixed point: integer addition is associative, so the result -- and
//      with it the whole trajectory -- is bit-identical whatever order the sort left the cells of a bin in;
//   4. shell partners that were bonded before go through the same contact model (bond hysteresis);
//   5. the new bond row, collected in the shared-memory column, is written with 16-byte stores; the cell is
//      integrated and pre-binned for the next substep's sort.
// Measured alternatives that did NOT help (profiles/r01_kernels_c3.md): staging the candidate bin tiles in shared
// memory, prefetching every cell's previous bond row, 64/96-register variants, ordering partners by id.
#ifndef ABM_JKR_MINBLOCKS
#define ABM_JKR_MINBLOCKS 6      // register cap: 6 -> 80/thread
#endif
constexpr int kMaxHits = 16;
constexpr int kShellFlag = 0x40000000;
constexpr int kJkrMinBlocks = ABM_JKR_MINBLOCKS;
constexpr double kForceUnit = 0x1p-76, kForceUnitInv = 0x1p76;   // fixed-point force accumulation (N)

struct BondEval {
    // radius-dependent part of the contact model, memoised on the partner radius (radii are usually uniform).
    // The two divisions of the reference's  d = ((r1+r2-mag)/1e6)/overlap_  and the force scale
    // pi*adhesion*r_hat are folded into two factors here, so the per-pair work is one rsqrt and a few FMAs.
    // This moves d by ~2 ulp: forces stay within 1e-15 relative, and the bond test d > -0.360562 can differ
    // from the reference only for |d + 0.360562| < 1e-15 (never observed; both ends of a pair use the same
    // expression, so the two rows of a bond always agree).
    double memo_r, inv_scale, force_scale, rsum;
    __device__ __forceinline__ void preset(const JkrConst &c, double rp) {
        memo_r = -1.0;
        if (rp == c.uni_r) {                 // colony of equal radii: the host evaluated the model once
            memo_r = c.uni_r;
            inv_scale = c.uni_inv_scale;
            force_scale = c.uni_force_scale;
            rsum = c.uni_r + c.uni_r;
        }
    }
    __device__ __forceinline__ void refresh(double rp, double rq, const JkrConst &c) {
        if (rq != memo_r) {
            const double kPi = 3.141592653589793;
            memo_r = rq;
            double r_hat = 1.0 / (1e6 * ((1.0 / rp) + (1.0 / rq)));
            double overlap_ = c.k23 * cbrt(r_hat);
            inv_scale = 1.0 / (1e6 * overlap_);
            force_scale = kPi * c.adhesion * r_hat;
            rsum = rp + rq;
        }
    }
};

struct NextSort {          // where to pre-bin the moved cells for the following substep (counts == nullptr: off)
    BinGeom geom;
    int *counts, *cell_bin, *rank;
};

__global__ void __launch_bounds__(kJkrThreads, kJkrMinBlocks)
k_jkr_substep(const int *__restrict__ bin_start, const int *__restrict__ sorted_id,
              const double4 *__restrict__ spos, const int *__restrict__ sorted_bin, int n, Stencil st, Radius rad,
              double reach2, GraphView old, int *__restrict__ new_len, int *__restrict__ new_cols, int stride,
              int *__restrict__ max_len, const double *__restrict__ mot, double *__restrict__ loc, JkrConst c,
              double stokes, Box box, double dt, const int *__restrict__ gid_of, NextSort next, int sort_rows) {
    // column tid: first the parked partner slots, then (overwritten from the front) the ids of the new bond row
    __shared__ int column[kMaxHits][kJkrThreads];
    const int tid = threadIdx.x;
    int s = blockIdx.x * kJkrThreads + tid;
    int cnt = 0;
    if (s < n) {
        int i = sorted_id[s];
        double4 p = spos[s];
        int bx, by, bz;
        st.decode(sorted_bin[s], bx, by, bz);
        int *row = new_cols + (long long)i * stride;
        long long fxi = 0, fyi = 0, fzi = 0;
        BondEval ev;
        ev.preset(c, p.w);
        long long a3 = 3 * (long long)i;
        const double mx = mot[a3 + 0], my = mot[a3 + 1], mz = mot[a3 + 2];   // latency hidden by the walk

        // the id bond rows are written in: the cell index, or its global id when the colony is slab-decomposed
        auto partner_id = [&](int t) {
            int j = sorted_id[t];
            return gid_of ? gid_of[j] : j;
        };
        int direct = 0;      // bonds of a crowded cell written straight to row[kMaxHits + direct]
        // contact model for one bond candidate; a surviving bond is recorded in the column (or, for a crowded cell
        // during the walk, directly in the global row)
        auto bond = [&](double4 q, int j, bool to_column) {
            double vx = p.x - q.x, vy = p.y - q.y, vz = p.z - q.z;
            double s2 = sumsq_nofma(vx, vy, vz);
            ev.refresh(p.w, q.w, c);
            double inv_mag = rsqrt(s2);                     // +inf for coincident cells, handled below
            double mag = s2 > 0.0 ? s2 * inv_mag : 0.0;
            double d = (ev.rsum - mag) * ev.inv_scale;
            if (d > -0.360562) {
                double f = (-0.0204 * (d * d * d)) + (0.4942 * (d * d)) + (1.0801 * d) - 1.324;
                if (s2 > 0.0) {
                    double w = f * ev.force_scale * inv_mag * kForceUnitInv;
                    fxi += __double2ll_rn(w * vx);
                    fyi += __double2ll_rn(w * vy);
                    fzi += __double2ll_rn(w * vz);
                }
                if (to_column && cnt < kMaxHits) {
                    column[cnt++][tid] = j;
                } else {
                    if (kMaxHits + direct < stride) row[kMaxHits + direct] = j;
                    ++direct;
                }
            }
        };
        // membership of j in this cell's previous bond row, straight from global memory (crowded cells only)
        auto was_bonded_slow = [&](int j) {
            long long os = old.start(i);
            int olen = old.row_len[i];
            bool found = false;
            for (int k = 0; k < olen; ++k) found |= (old.cols[os + k] == j);
            return found;
        };

        // phase 1: cheap tests over the stencil.  The (lo,hi) slot ranges of the three dy runs of a z-layer are
        // fetched together and walked in one loop (the warp iterates max(total) times, not sum(max(run))).
        int nh = 0, nsh = 0, sh0 = 0, sh1 = 0;
        const int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
        for (int dz = -1; dz <= 1; ++dz) {
            int z = bz + dz;
            if (z < 0 || z >= st.dims[2]) continue;
            int run_lo[3], run_n[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                int y = by + r - 1;
                bool ok = y >= 0 && y < st.dims[1];
                int rowb = ok ? (z * st.dims[1] + y) * st.dims[0] : 0;
                int lo = bin_start[rowb + x0], hi = bin_start[rowb + x1 + 1];
                run_lo[r] = lo;
                run_n[r] = ok ? hi - lo : 0;
            }
            const int n0 = run_n[0], n01 = n0 + run_n[1], tot = n01 + run_n[2];
            for (int u = 0; u < tot; ++u) {
                int t = u < n0 ? run_lo[0] + u : (u < n01 ? run_lo[1] + (u - n0) : run_lo[2] + (u - n01));
                double4 q = spos[t];
                double s2 = sumsq_nofma(p.x - q.x, p.y - q.y, p.z - q.z);
                if (s2 > reach2 || t == s) continue;          // no bond can survive beyond the reach
                if (rad.within(s2)) {
                    if (nh < kMaxHits) column[nh++][tid] = t;
                    else bond(q, partner_id(t), false);       // crowded cell: evaluate in place
                } else if (old.valid()) {                     // thin shell: only pre-existing bonds persist
                    if (nsh == 0) sh0 = t;
                    else if (nsh == 1) sh1 = t;
                    else if (nh < kMaxHits) column[nh++][tid] = t | kShellFlag;   // rare third+ shell partner
                    else {
                        int j = partner_id(t);
                        if (was_bonded_slow(j)) bond(q, j, false);
                    }
                    ++nsh;
                }
            }
        }
        // phase 2: lanes with shell partners fetch (the head of) their previous bond row now
        int olen = 0;
        int4 oa = make_int4(-1, -1, -1, -1), ob = oa;
        long long os = 0;
        if (nsh > 0) {
            os = old.start(i);
            olen = old.row_len[i];
            if ((os & 3) == 0) {                               // fixed-stride rows (stride % 4 == 0): 16-byte loads
                if (olen > 0) oa = *reinterpret_cast<const int4 *>(old.cols + os);
                if (olen > 4) ob = *reinterpret_cast<const int4 *>(old.cols + os + 4);
            }
        }
        auto in_old_row = [&](int j) {
            if ((os & 3) != 0) return was_bonded_slow(j);
            bool f = false;
            if (olen > 0) f |= (oa.x == j);
            if (olen > 1) f |= (oa.y == j);
            if (olen > 2) f |= (oa.z == j);
            if (olen > 3) f |= (oa.w == j);
            if (olen > 4) f |= (ob.x == j);
            if (olen > 5) f |= (ob.y == j);
            if (olen > 6) f |= (ob.z == j);
            if (olen > 7) f |= (ob.w == j);
            for (int k = 8; k < olen; ++k) f |= (old.cols[os + k] == j);
            return f;
        };
        // phase 3: contact model over the parked partners, all lanes together
        for (int k = 0; k < nh; ++k) {
            int t = column[k][tid];
            bool shell = t & kShellFlag;
            t &= ~kShellFlag;
            int j = partner_id(t);
            if (shell && !in_old_row(j)) continue;
            bond(spos[t], j, true);                           // cnt <= k: the column is overwritten behind the read
        }
        // phase 4: hysteresis -- shell partners that are in the previous row
        if (nsh > 0) {
            int j0 = partner_id(sh0);
            if (in_old_row(j0)) bond(spos[sh0], j0, true);
            if (nsh > 1) {
                int j1 = partner_id(sh1);
                if (in_old_row(j1)) bond(spos[sh1], j1, true);
            }
        }
        // phase 5: the bond row.  cnt entries sit in the column; a crowded cell has `direct` more already in global
        // memory from row[kMaxHits] on, which are moved down so that the row is dense.
        int total = cnt + direct;
        int kept = min(total, stride);
        if (direct == 0) {
            if (sort_rows) {                                   // ascending rows for the host-visible graph
                for (int a = 1; a < cnt; ++a) {
                    int v = column[a][tid], b = a - 1;
                    while (b >= 0 && column[b][tid] > v) {
                        column[b + 1][tid] = column[b][tid];
                        --b;
                    }
                    column[b + 1][tid] = v;
                }
            }
            int m = min(cnt, stride);
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
        } else {
            int moved = min(direct, max(stride - kMaxHits, 0));
            for (int k = 0; k < min(cnt, stride); ++k) row[k] = column[k][tid];
            if (cnt < kMaxHits)
                for (int k = 0; k < moved; ++k) row[cnt + k] = row[kMaxHits + k];
            if (sort_rows) insertion_sort(row, kept);
        }
        new_len[i] = kept;
        // what the overflow report below needs: the row length, and room for the direct entries of a crowded cell
        // (they live behind the first kMaxHits slots until they are moved down)
        cnt = direct ? max(total, kMaxHits + direct) : total;
        // integration (reference cell_backend.py:155-170); positions are read from the sorted snapshot only
        double friction = stokes_friction(p.w, stokes);
        double nx = stokes_move(p.x, mx, (double)fxi * kForceUnit, friction, dt, box.size[0]);
        double ny = stokes_move(p.y, my, (double)fyi * kForceUnit, friction, dt, box.size[1]);
        double nz = stokes_move(p.z, mz, (double)fzi * kForceUnit, friction, dt, box.size[2]);
        loc[a3 + 0] = nx;
        loc[a3 + 1] = ny;
        loc[a3 + 2] = nz;
        if (next.counts) {                                    // histogram + rank for the next substep's sort
            int b = cell_bin_of(nx, ny, nz, next.geom);
            next.cell_bin[i] = b;
            next.rank[i] = atomicAdd(&next.counts[b], 1);
        }
    }
    // overflow report: one atomic per warp, and only when a row did not fit
    int wmax = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
    if (lane_id() == 0 && wmax > stride) atomicMax(max_len, wmax);
}

}  // namespace abm

using namespace abm;

ABM_EXPORT int abm_jkr_forces(const int32_t *row_start, int64_t row_stride, int32_t *row_len, int32_t *cols,
                              const double *locations, const double *radii, double *jkr_forces, int64_t n,
                              double poisson, double youngs, double adhesion_const, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(row_len && locations && radii && jkr_forces, "null pointer");
    JkrConst c = make_jkr_const(poisson, youngs, adhesion_const, 0.0);
    k_jkr_rows<<<ceil_div(n, kJkrThreads), kJkrThreads, 0, as_stream(stream)>>>(
        row_start, (long long)row_stride, row_len, cols, locations, radii, jkr_forces, (int)n, c);
    ABM_CHECK_LAUNCH("k_jkr_rows");
    return 0;
}

ABM_EXPORT int abm_apply_forces(const double *jkr_forces, const double *motility_forces, double *locations,
                                const double *radii, int64_t n, double stokes, const double size_host[3],
                                double move_dt, double *jkr_forces_to_zero, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(jkr_forces && motility_forces && locations && radii && size_host, "null pointer");
    Box box;
    for (int k = 0; k < 3; ++k) box.size[k] = size_host[k];
    k_integrate<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(jkr_forces, motility_forces, locations, radii,
                                                                (int)n, stokes, box, move_dt, jkr_forces_to_zero);
    ABM_CHECK_LAUNCH("k_integrate");
    return 0;
}

ABM_EXPORT int abm_jkr_substep(const int32_t *bin_start, const int32_t *sorted_id, const double *sorted_pos,
                               const int32_t *sorted_bin, int64_t n, const int32_t dims_host[3], double distance,
                               double reach, const int32_t *old_row_start, int64_t old_row_stride,
                               const int32_t *old_row_len, const int32_t *old_cols, int32_t *new_row_len,
                               int32_t *new_cols, int64_t row_stride, int32_t *max_len, int sort_rows,
                               const double *motility_forces, double *locations, double poisson, double youngs,
                               double adhesion_const, double stokes, const double size_host[3], double move_dt,
                               const int32_t *global_id, double uniform_radius, const double *next_extent_host,
                               const double *next_origin_host, int32_t *next_bin_counts, void *next_sort_workspace,
                               void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(bin_start && sorted_id && sorted_pos && sorted_bin && dims_host && new_row_len && new_cols &&
                    max_len && motility_forces && locations && size_host,
                "null pointer");
    ABM_REQUIRE(row_stride > 0 && row_stride < (1 << 20), "bad row_stride");
    ABM_REQUIRE(reach >= distance && distance > 0, "reach must be >= distance > 0");
    ABM_REQUIRE(old_cols != new_cols, "old and new bond rows must be different buffers");
    JkrConst c = make_jkr_const(poisson, youngs, adhesion_const, uniform_radius);
    GraphView old{old_row_start, (long long)old_row_stride, old_row_len, old_cols};
    Stencil stc;
    Box box;
    for (int k = 0; k < 3; ++k) {
        stc.dims[k] = dims_host[k];
        box.size[k] = size_host[k];
    }
    NextSort next;
    memset(&next, 0, sizeof(next));
    if (next_bin_counts) {
        ABM_REQUIRE(next_extent_host && next_sort_workspace && next_bin_counts != bin_start,
                    "pre-binning needs the next sort's extent, workspace and a second offset table");
        int64_t nbins;
        ABM_REQUIRE(make_bin_geom(next_extent_host, next_origin_host, reach, 0, next.geom, nbins) == 0,
                    "bad next-sort geometry");
        SortWs w = sort_ws_carve(next_sort_workspace, n, nbins);
        next.counts = next_bin_counts;
        next.cell_bin = w.cell_bin;
        next.rank = w.rank;
    }
    k_jkr_substep<<<ceil_div(n, kJkrThreads), kJkrThreads, 0, as_stream(stream)>>>(
        bin_start, sorted_id, reinterpret_cast<const double4 *>(sorted_pos), sorted_bin, (int)n, stc,
        Radius(distance), reach * reach, old, new_row_len, new_cols, (int)row_stride, max_len, motility_forces,
        locations, c, stokes, box, move_dt, global_id, next, sort_rows);
    ABM_CHECK_LAUNCH("k_jkr_substep");
    return 0;
}

// The whole substep loop of CellMethods.apply_forces (reference cell_methods.py:399-436) enqueued by one call:
// n_substeps x [counting sort at bin edge `reach` + fused substep].  Substep k reads the bonds written by
// substep k-1 (substep 0 reads old_*) and writes row buffer a (k even) or b (k odd); the final bonds are in
// buffer a when n_substeps is odd, else in b.  Every substep pre-bins the moved cells, so only the first sort
// of the loop runs the histogram kernel; the two offset tables alternate.
ABM_EXPORT int abm_mechanics_loop(const double *radii, double *locations, const double *motility_forces, int64_t n,
                                  const double extent_host[3], const double *origin_host, const double size_host[3],
                                  double distance, double reach, int n_substeps, const double *dts_host,
                                  int32_t *bin_start_a, int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos,
                                  int32_t *sorted_bin, int32_t *max_count, void *sort_workspace,
                                  size_t sort_workspace_bytes, const int32_t *old_row_start, int64_t old_row_stride,
                                  const int32_t *old_row_len, const int32_t *old_cols, int32_t *len_a,
                                  int32_t *cols_a, int32_t *len_b, int32_t *cols_b, int64_t row_stride,
                                  int32_t *max_len, const int32_t *global_id, double uniform_radius, double poisson,
                                  double youngs, double adhesion_const, double stokes, void *stream) {
    ABM_REQUIRE(n_substeps >= 0 && dts_host, "bad substep list");
    ABM_REQUIRE(extent_host && size_host && bin_start_a && bin_start_b && bin_start_a != bin_start_b,
                "need two distinct offset tables");
    BinGeom g;
    int64_t nbins;
    ABM_REQUIRE(make_bin_geom(extent_host, origin_host, reach, 0, g, nbins) == 0, "bad geometry");
    cudaStream_t st = as_stream(stream);
    const int32_t *cur_start = old_row_start, *cur_len = old_row_len, *cur_cols = old_cols;
    int64_t cur_stride = old_row_stride;
    for (int k = 0; k < n_substeps; ++k) {
        int32_t *table = (k & 1) ? bin_start_b : bin_start_a, *other = (k & 1) ? bin_start_a : bin_start_b;
        int rc = abm_bin_sort(locations, radii, n, extent_host, origin_host, reach, k ? ABM_SORT_PREBINNED : 0, table,
                              sorted_id, sorted_pos, sorted_bin, max_count, sort_workspace, sort_workspace_bytes,
                              stream);
        if (rc) return rc;
        const bool prebin = k + 1 < n_substeps && n > 0;
        if (prebin) {
            rc = check_cuda(cudaMemsetAsync(other, 0, (size_t)(nbins + 1) * 4, st), "memset next offset table");
            if (rc) return rc;
        }
        int32_t *nl = (k & 1) ? len_b : len_a, *nc = (k & 1) ? cols_b : cols_a;
        rc = abm_jkr_substep(table, sorted_id, sorted_pos, sorted_bin, n, g.dims, distance, reach, cur_start,
                             cur_stride, cur_len, cur_cols, nl, nc, row_stride, max_len, k == n_substeps - 1,
                             motility_forces, locations,
                             poisson, youngs, adhesion_const, stokes, size_host, dts_host[k], global_id, uniform_radius,
                             prebin ? extent_host : nullptr, origin_host, prebin ? other : nullptr,
                             prebin ? sort_workspace : nullptr, stream);
        if (rc) return rc;
        cur_start = nullptr;
        cur_stride = row_stride;
        cur_len = nl;
        cur_cols = nc;
    }
    return 0;
}
