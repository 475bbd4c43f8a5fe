// jkr.cu -- JKR adhesion/repulsion forces, Stokes integration, and the fused mechanics substep.
// Replaces jkr_forces_cpu/gpu, apply_forces_cpu/gpu and the body of the 11-substep loop of
// CellMethods.apply_forces (reference cell_backend.py:6-172, cell_methods.py:342-439).
#include <string.h>

#include "graph.cuh"
#include "sortws.cuh"
#include "tile.cuh"

namespace abm {

#ifndef ABM_JKR_THREADS
#define ABM_JKR_THREADS 128
#endif
constexpr int kJkrThreads = ABM_JKR_THREADS;

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
//   1. filter the bin stencil: the candidates of the warp's 32 cells are staged in shared memory one z layer at a
//      time (tile.cuh) and tested in fp32 against reach^2 + error band; survivors (their slots) are parked in a
//      per-thread column of shared memory.  The filter can only let through too many, never too few;
//   2. contact loop over the parked partners, all lanes together (trip count = the largest hit count in the
//      warp): the exact fp64 squared distance decides (discard beyond the reach / bond within `distance` /
//      hysteresis shell in between), then the JKR model.  The force is summed in 2^-76 N fixed point: integer
//      addition is associative, so the result -- and with it the whole trajectory -- is bit-identical whatever
//      order the sort left the cells of a bin in;
//   3. the (few) shell partners are looked up in the cell's previous bond row and, if bonded before, go through
//      the same contact model (bond hysteresis);
//   4. the new bond row, collected in the shared-memory column, is written with 16-byte stores; the cell is
//      integrated and binned for the next substep's sort.
// A cell with more than kMaxHits partners, or whose candidate run does not fit the tile, takes a slow path that
// walks the stencil from global memory and evaluates every candidate in place.
#ifndef ABM_JKR_MINBLOCKS
#define ABM_JKR_MINBLOCKS 7      // register cap: 7 -> 72/thread (measured best of 5..8 at 500k cells)
#endif
// Build-time experiment switches (scripts/ab_substep.py builds variant libraries with -DABM_JKR_EXP=<bits>; the
// shipped library uses 0).  Measured outcomes are in profiles/r01b_kernels_c3.md.
//   1: load the motility force before the contact loop instead of at the end      (slower: register pressure)
//   2: start fetching the previous bond row at the first shell partner             (slower: register pressure)
//   4: TIMING ABLATION, WRONG RESULTS: skip the contact loop                        (attributes 23 of 56 us)
//   8: TIMING ABLATION, WRONG RESULTS: skip the candidate walk                      (attributes  9 of 56 us)
//  64: TIMING ABLATION, WRONG RESULTS (2-D only): the contact loop takes the partner's position (fp32, from the
//      warp tile) and "id" (its slot) from shared memory instead of gathering spos[] / the id table from global
//      memory -- an upper bound on what staging fp64 positions and ids with the tile could gain
//  16: bond rows keyed by sorted slot between the substeps of abm_mechanics_loop (written coalesced; the previous
//      row is found through the prev_slot index the placement kernel records); the first substep still reads and
//      the last one still writes rows keyed by cell id                               (written blind: not yet measured)
//  32: contact model with a Horner-form force polynomial and an rsqrt without special-case handling (s2 is a normal
//      number; forces move by ~1e-16 relative, inside the 1e-12 gate)
// Measured and REMOVED in round 2 (profiles/r02_kernels_c3.md; the code is in git history, commit "k_jkr_substep_flat"):
// a lane-balanced contact phase -- the warp's pairs flattened into one shared-memory list and evaluated 32 at a time,
// in-warp pairs evaluated once, forces and bond rows accumulated with 32-bit shared-memory atomics.  Bit-identical
// results (all GPU tests green), 3.0 rounds instead of 8.1 iterations, but 78.8 us instead of 56.3 us per launch:
// the ~14 conflicting shared atomics per pair cost more than the idle lanes they remove.
#ifndef ABM_JKR_EXP
#define ABM_JKR_EXP 0
#endif
#if (ABM_JKR_EXP & (12 | 64)) && !defined(ABM_ALLOW_WRONG_RESULTS)
#error "ABM_JKR_EXP bits 4, 8 and 64 are timing ablations that produce wrong results: add -DABM_ALLOW_WRONG_RESULTS"
#endif
#ifndef ABM_PDL_LATE
#define ABM_PDL_LATE 0      // 1: trigger the dependent launch after the contact loop instead of at kernel entry
#endif
#ifndef ABM_JKR_MAXHITS
#define ABM_JKR_MAXHITS 16
#endif
constexpr int kMaxHits = ABM_JKR_MAXHITS;
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

// 1/sqrt(x) for a normal, positive x (squared centre distances: 1 .. ~110 um^2): single-precision seed and two Newton
// steps y += y*(0.5 - (x/2)*y*y), no handling of zeros, infinities or subnormals (x == 0 yields NaN, which the callers
// never use).  Experiment ABM_JKR_EXP & 32.
__device__ __forceinline__ double rsqrt_normal(double x) {
    double y = (double)rsqrtf((float)x);
    const double h = 0.5 * x;
    y = fma(y, fma(-h * y, y, 0.5), y);
    y = fma(y, fma(-h * y, y, 0.5), y);
    return y;
}

struct NextSort {          // hand-over of the moved cells to the following substep's sort (counts == nullptr: off),
    BinGeom geom;          // in THIS sort's slot order: the next permutation is then nearly the identity
    int *counts, *cell_bin, *prev_id, *chunk_sums;
    double4 *next_pos;
    int compact;           // the hand-over's position records are 16-byte {x, y} (kSortCompactXY), not {x, y, z, r}
};

struct Friction {          // Stokes friction of a colony of equal radii and its correctly rounded reciprocal
    double uni, uni_inv;   // uni == 0: radii differ, every thread evaluates its own
};

// a / b rounded to nearest, given y = RN(1/b): Markstein's correction applied twice (the first brings the
// quotient within one ulp, from where the second is proven to round correctly).  Same bits as __ddiv_rn for
// normal operands, a quarter of its instructions.
__device__ __forceinline__ double div_by(double a, double b, double y) {
    double q = __dmul_rn(a, y);
    q = __fma_rn(__fma_rn(-b, q, a), y, q);
    return __fma_rn(__fma_rn(-b, q, a), y, q);
}

// Stokes integration of one axis with the division replaced by div_by (reference cell_backend.py:155-170)
__device__ __forceinline__ double stokes_move_y(double x, double f_mot, double f_jkr, double friction, double inv_f,
                                                double dt, double size_k) {
    double velocity = div_by(__dadd_rn(f_mot, f_jkr), friction, inv_f);
    double moved = __dadd_rn(x, __dmul_rn(dt, __dmul_rn(velocity, 1e6)));
    if (moved > size_k) return size_k;
    if (moved < 0.0) return 0.0;
    return moved;
}

// kUniform: all radii equal c.uni_r -- the radius-dependent contact factors and the friction are kernel constants.
template <bool kUniform>
__global__ void __launch_bounds__(kJkrThreads, kJkrMinBlocks)
k_jkr_substep(const int *__restrict__ bin_start, const int *__restrict__ sorted_id,
              const double4 *__restrict__ spos, const int *__restrict__ sorted_bin, int n, Stencil st, Radius rad,
              double reach2, BandConst bandc, GraphView old, int *__restrict__ new_len, int *__restrict__ new_cols, int stride,
              int *__restrict__ max_len, const double *__restrict__ mot, double *__restrict__ loc, JkrConst c,
              double stokes, Friction fric, Box box, double dt, const int *__restrict__ gid_of,
              const int *__restrict__ slot_pid, NextSort next, int sort_rows, const int *__restrict__ prev_slot,
              int row_keys, const int *__restrict__ n_dev, const double *__restrict__ move_window) {
    // column tid: first the parked partner slots, then (overwritten from the front) the ids of the new bond row
    __shared__ int column[kMaxHits][kJkrThreads];
    __shared__ float4 tiles[kJkrThreads / 32][kTileRows * kTileCap];
#if !ABM_PDL_LATE
    pdl_launch_dependents();
#endif
    const int tid = threadIdx.x;
    const int s = blockIdx.x * kJkrThreads + tid;
    pdl_wait();
    n = device_count(n, n_dev);
    if (s - (tid & 31) >= n) return;                           // the whole warp is past the last cell
    const bool live = s < n;
    float4 *tile = tiles[tid >> 5];
    int i = 0, bx = 0, by = 0, bz = 0;
    double4 p = make_double4(0.0, 0.0, 0.0, 1.0);
    if (live) {
        i = sorted_id[s];
        p = spos[s];
        st.decode(sorted_bin[s], bx, by, bz);
        // the motility force is loaded after the walk (registers) and used at the very end: pull its line in now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(mot + 3 * (long long)i));
    }

    // ---- phase 1: fp32 filter over the staged stencil ----------------------------------------------------------
    TileOrigin org;
    org.x = __shfl_sync(kFullWarp, p.x, 0);
    org.y = __shfl_sync(kFullWarp, p.y, 0);
    org.z = __shfl_sync(kFullWarp, p.z, 0);
    const float pfx = (float)(p.x - org.x), pfy = (float)(p.y - org.y), pfz = (float)(p.z - org.z);
    const float t_out = (float)reach2 + bandc.band(pfx, pfy, pfz);
    int nh = 0;                                                // parked partners (counts on past kMaxHits)
    bool slow = false;
    for (int dz = -1; dz <= 1; ++dz) {
#if ABM_JKR_EXP & 8
        break;                                                 // timing ablation: no walk (wrong results)
#endif
        const int z = bz + dz;
        if (__all_sync(kFullWarp, !live || z < 0 || z >= st.dims[2])) continue;
        const LayerRuns R = stage_layer(live, bx, by, z, st, bin_start, spos, org, tile);
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
            if (!((R.staged >> r) & 1u)) {                     // warp-uniform
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
#if ABM_JKR_EXP & 64
                    if (nh < kMaxHits) column[nh][tid] = R.first[r] + u;
#else
                    if (nh < kMaxHits) column[nh][tid] = t;
#endif
                    ++nh;
                }
            }
        }
        __syncwarp();                                         // the tile is overwritten by the next layer
    }
    if (!live) return;                                         // no warp-wide operation below this line
    slow |= nh > kMaxHits;
#if ABM_JKR_EXP & 4
    nh = 0;                                                    // timing ablation: no contact loop (wrong results)
    slow = false;
#endif

#if ABM_JKR_EXP & 1
    const double mx = mot[3 * (long long)i + 0], my = mot[3 * (long long)i + 1], mz = mot[3 * (long long)i + 2];
#endif
#if ABM_JKR_EXP & 16
    const int oi = (row_keys & 1) ? prev_slot[s] : i;          // key of this cell's previous bond row
    const int ni = (row_keys & 2) ? s : i;                     // key of the row it writes
#else
    const int oi = i, ni = i;
#endif
    int *row = new_cols + (long long)ni * stride;
    long long fxi = 0, fyi = 0, fzi = 0;
    BondEval ev;
    if (!kUniform) ev.preset(c, p.w);
    int cnt = 0;         // bonds recorded in the column
    int direct = 0;      // bonds of a crowded cell written straight to row[kMaxHits + direct]
    // the id bond rows are written in: the cell index, or its global id when the colony is slab-decomposed
    // (slot_pid, when given, holds that id by slot: one gather instead of two dependent ones)
    auto partner_id = [&](int t) {
        if (slot_pid) return slot_pid[t];
        int j = sorted_id[t];
        return gid_of ? gid_of[j] : j;
    };
    const int *first_hop = slot_pid ? slot_pid : sorted_id;
    const bool second_hop = !slot_pid && gid_of;
    // contact model for one bond candidate at offset v = p - q, |v|^2 = s2; a surviving bond is recorded in the
    // column (or, for a crowded cell, directly in the global row)
    auto bond = [&](double vx, double vy, double vz, double s2, double rq, int j) {
        if (!kUniform) ev.refresh(p.w, rq, c);
        const double rsum = kUniform ? c.uni_r + c.uni_r : ev.rsum;
        const double inv_scale = kUniform ? c.uni_inv_scale : ev.inv_scale;
        const double force_scale = kUniform ? c.uni_force_scale : ev.force_scale;
#if ABM_JKR_EXP & 32
        double inv_mag = rsqrt_normal(s2);                  // NaN for coincident cells, not used below
#else
        double inv_mag = rsqrt(s2);                         // +inf for coincident cells, handled below
#endif
        double mag = s2 > 0.0 ? s2 * inv_mag : 0.0;
        double d = (rsum - mag) * inv_scale;
        if (d > -0.360562) {
#if ABM_JKR_EXP & 32
            double f = fma(fma(fma(-0.0204, d, 0.4942), d, 1.0801), d, -1.324);
#else
            double f = (-0.0204 * (d * d * d)) + (0.4942 * (d * d)) + (1.0801 * d) - 1.324;
#endif
            if (s2 > 0.0) {
                double w = f * force_scale * inv_mag * kForceUnitInv;
                fxi += __double2ll_rn(w * vx);
                fyi += __double2ll_rn(w * vy);
                fzi += __double2ll_rn(w * vz);
            }
            if (cnt < kMaxHits) {
                column[cnt++][tid] = j;
            } else {
                if (kMaxHits + direct < stride) row[kMaxHits + direct] = j;
                ++direct;
            }
        }
    };
    // membership of j in this cell's previous bond row, straight from global memory (slow path only)
    auto was_bonded_slow = [&](int j) {
        long long os = old.start(oi);
        int olen = old.row_len[oi];
        bool found = false;
        for (int k = 0; k < olen; ++k) found |= (old.cols[os + k] == j);
        return found;
    };

    int nsh = 0, sh0 = 0, sh1 = 0, shj0 = 0, shj1 = 0;         // hysteresis-shell partners, looked up in phase 3
    const long long os = old.valid() ? old.start(oi) : 0;
    // fixed-stride rows of >= 8 entries (stride % 4 == 0): the first 8 ids come as two unconditional 16-byte loads
    const bool wide = old.valid() && !old.row_start && (old.stride & 3) == 0 && old.stride >= 8;
    int olen = 0;
    int4 oa = make_int4(-1, -1, -1, -1), ob = oa;
    if (!slow) {
        // ---- phase 2: exact distances and the contact model over the parked partners ---------------------------
        // (the next partner's position and id are fetched while the current one is evaluated)
#if ABM_JKR_EXP & 64
        auto fetch = [&](int e, int &t, double4 &q, int &j) {
            const float4 qf = tile[e];
            t = __float_as_int(qf.w);
            q = make_double4(org.x + (double)qf.x, org.y + (double)qf.y, org.z + (double)qf.z, p.w);
            j = t;
        };
        int t_next = s, j_next = 0;
        double4 q_next = p;
        if (nh > 0) fetch(column[0][tid], t_next, q_next, j_next);
#else
        int t_next = nh > 0 ? column[0][tid] : s;
        double4 q_next = spos[t_next];
        int j_next = first_hop[t_next];
#endif
        for (int k = 0; k < nh; ++k) {
            const int t = t_next;
            const double4 q = q_next;
            int j = j_next;
            if (k + 1 < nh) {
#if ABM_JKR_EXP & 64
                fetch(column[k + 1][tid], t_next, q_next, j_next);
#else
                t_next = column[k + 1][tid];
                q_next = spos[t_next];
                j_next = first_hop[t_next];
#endif
            }
            const double vx = p.x - q.x, vy = p.y - q.y, vz = p.z - q.z;
            const double s2 = sumsq_nofma(vx, vy, vz);
            if (s2 > reach2) continue;                        // let through by the fp32 filter's error band only
            if (rad.within(s2)) {
                if (second_hop) j = gid_of[j];
                bond(vx, vy, vz, s2, q.w, j);                 // cnt <= k: the column is overwritten behind the read
            } else if (old.valid()) {                         // thin shell: only pre-existing bonds persist
                if (second_hop) j = gid_of[j];
                if (nsh == 0) {
                    sh0 = t;
                    shj0 = j;
#if ABM_JKR_EXP & 2
                    olen = old.row_len[oi];                   // the previous row starts its way here now
                    if (wide) {
                        oa = *reinterpret_cast<const int4 *>(old.cols + os);
                        ob = *reinterpret_cast<const int4 *>(old.cols + os + 4);
                    }
#endif
                } else if (nsh == 1) {
                    sh1 = t;
                    shj1 = j;
                } else if (was_bonded_slow(j)) {              // rare third+ shell partner
                    bond(vx, vy, vz, s2, q.w, j);
                }
                ++nsh;
            }
        }
        // ---- phase 3: hysteresis -- shell partners that are in the previous row --------------------------------
        if (nsh > 0) {
#if !(ABM_JKR_EXP & 2)
            olen = old.row_len[oi];
            if (wide) {
                oa = *reinterpret_cast<const int4 *>(old.cols + os);
                ob = *reinterpret_cast<const int4 *>(old.cols + os + 4);
            }
#endif
            auto in_old_row = [&](int j) {
                if (!wide) return was_bonded_slow(j);
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
            const int nshell = min(nsh, 2);
            for (int k = 0; k < nshell; ++k) {
                const int t = k ? sh1 : sh0;
                const int j = k ? shj1 : shj0;
                if (in_old_row(j)) {
                    const double4 q = spos[t];
                    const double vx = p.x - q.x, vy = p.y - q.y, vz = p.z - q.z;
                    bond(vx, vy, vz, sumsq_nofma(vx, vy, vz), q.w, j);
                }
            }
        }
    } else {
        // ---- slow path: crowded cell or candidate run longer than the tile -------------------------------------
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
                    const double4 q = spos[t];
                    const double vx = p.x - q.x, vy = p.y - q.y, vz = p.z - q.z;
                    const double s2 = sumsq_nofma(vx, vy, vz);
                    if (s2 > reach2) continue;
                    const int j = partner_id(t);
                    if (rad.within(s2) || (old.valid() && was_bonded_slow(j))) bond(vx, vy, vz, s2, q.w, j);
                }
            }
        }
    }

#if ABM_PDL_LATE
    // the successor (the next sort's scan) may start being scheduled once every CTA of this grid has come this far:
    // its CTAs then take the place of this grid's last CTAs as they retire, instead of holding registers beside
    // a whole wave that still has its contact loop to run
    pdl_launch_dependents();
#endif
    // ---- phase 4: the bond row.  cnt entries sit in the column; a crowded cell has `direct` more already in
    // global memory from row[kMaxHits] on (cnt == kMaxHits then, so the row is dense)
    const int total = cnt + direct;
    const int kept = min(total, stride);
    if (sort_rows && direct == 0) {                            // ascending rows for the host-visible graph
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
    new_len[ni] = kept;
    if (total > stride) atomicMax(max_len, total);            // overflow report (rare): the host doubles the stride
    // integration (reference cell_backend.py:155-170); positions are read from the sorted snapshot only
    double friction = fric.uni, inv_f = fric.uni_inv;
    if (!kUniform) {
        friction = stokes_friction(p.w, stokes);
        inv_f = __drcp_rn(friction);
    }
#if !(ABM_JKR_EXP & 1)
    const double mx = mot[3 * (long long)i + 0], my = mot[3 * (long long)i + 1], mz = mot[3 * (long long)i + 2];
#endif
    const double nx = stokes_move_y(p.x, mx, (double)fxi * kForceUnit, friction, inv_f, dt, box.size[0]);
    const double ny = stokes_move_y(p.y, my, (double)fyi * kForceUnit, friction, inv_f, dt, box.size[1]);
    const double nz = stokes_move_y(p.z, mz, (double)fzi * kForceUnit, friction, inv_f, dt, box.size[2]);
    record_move(max_len, move_window,  p.x,
                (float)((nx - p.x) * (nx - p.x) + (ny - p.y) * (ny - p.y) + (nz - p.z) * (nz - p.z)));
    if (loc) {
        loc[3 * (long long)i + 0] = nx;
        loc[3 * (long long)i + 1] = ny;
        loc[3 * (long long)i + 2] = nz;
    }
    if (next.counts) {                                        // hand-over to the next substep's sort, by slot
        int b = cell_bin_of_loose(nx, ny, nz, next.geom);
        if (next.compact) reinterpret_cast<double2 *>(next.next_pos)[s] = make_double2(nx, ny);
        else next.next_pos[s] = make_double4(nx, ny, nz, p.w);
        next.cell_bin[s] = b;
        next.prev_id[s] = i;
        atomicAdd(&next.counts[b], 1);
        // the histogram's sum per scan chunk, so that the next sort needs no reduction pass: one atomic per
        // distinct chunk among the lanes here (neighbouring slots fall in the same chunk)
        const int chunk = b / kScanChunk;
        const unsigned peers = __match_any_sync(__activemask(), chunk);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&next.chunk_sums[chunk], __popc(peers));
    }
}

}  // namespace abm

#include "jkr_plane.cuh"

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

// slot_gid: optional (n) global id of the cell in every slot (written by the sort of the mechanics loop)
static int jkr_substep_impl(const int32_t *bin_start, const int32_t *sorted_id, const double *sorted_pos,
                            const int32_t *sorted_bin, int64_t n, const int32_t dims_host[3], double distance,
                            double reach, const int32_t *old_row_start, int64_t old_row_stride,
                            const int32_t *old_row_len, const int32_t *old_cols, int32_t *new_row_len,
                            int32_t *new_cols, int64_t row_stride, int32_t *max_len, int sort_rows,
                            const double *motility_forces, double *locations, double poisson, double youngs,
                            double adhesion_const, double stokes, const double size_host[3], double move_dt,
                            const int32_t *global_id, const int32_t *slot_gid, double uniform_radius,
                            const double *next_extent_host, const double *next_origin_host,
                            int32_t *next_bin_counts, void *next_sort_workspace, void *stream,
                            const int32_t *prev_slot = nullptr, int row_keys = 0, bool chained = false,
                            const int32_t *n_device = nullptr, const double *move_window = nullptr,
                            bool plane = false, bool compact_in = false, bool compact_out = false) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(!(row_keys & 1) || prev_slot, "rows keyed by the previous slot need the prev_slot index");
    ABM_REQUIRE(bin_start && sorted_id && sorted_pos && sorted_bin && dims_host && new_row_len && new_cols &&
                    max_len && motility_forces && size_host,
                "null pointer");
    ABM_REQUIRE(locations || next_bin_counts, "locations may be NULL only when the cells are handed to the next sort");
    ABM_REQUIRE(row_stride > 0 && row_stride < (1 << 20), "bad row_stride");
    ABM_REQUIRE(reach >= distance && distance > 0, "reach must be >= distance > 0");
    ABM_REQUIRE(old_cols != new_cols, "old and new bond rows must be different buffers");
    JkrConst c = make_jkr_const(poisson, youngs, adhesion_const, uniform_radius);
    GraphView old{old_row_start, (long long)old_row_stride, old_row_len, old_cols};
    Stencil stc;
    stc.set(dims_host);
    Box box;
    for (int k = 0; k < 3; ++k) box.size[k] = size_host[k];
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
        next.prev_id = w.tmp_id;
        next.chunk_sums = w.next_sums;
        next.next_pos = w.next_pos;
        next.compact = compact_out ? 1 : 0;
    }
    ABM_REQUIRE(!compact_in || (plane && uniform_radius > 0), "compact records are read by the two-axis kernel only");
    ABM_REQUIRE(!compact_out || (next_bin_counts && uniform_radius > 0 && size_host[2] == 0.0),
                "compact hand-over needs a flat world of equal radii");
    Friction fric{0.0, 0.0};
    if (uniform_radius > 0) {                                  // stokes_friction() on the host: same IEEE products
        const double kPi = 3.141592653589793;
        volatile double f = ((6.0 * kPi) * stokes) * (uniform_radius / 1e6);
        fric.uni = f;
        fric.uni_inv = 1.0 / f;
    }
#if ABM_JKR_EXP == 0
    // flat world, every z known to be exactly 0 (see jkr_plane.cuh): the two-axis kernel computes the same bits
    if (plane && stc.dims[2] == 1 && size_host[2] == 0.0) {
        auto kp = uniform_radius > 0 ? k_jkr_substep_plane<true> : k_jkr_substep_plane<false>;
        return check_cuda(launch_chained(kp, ceil_div(n, kJkrThreads), kJkrThreads, as_stream(stream), chained, bin_start,
                                         sorted_id, reinterpret_cast<const double4 *>(sorted_pos), sorted_bin, (int)n, stc,
                                         Radius(distance), reach * reach, old, new_row_len, new_cols, (int)row_stride,
                                         max_len, motility_forces, locations, c, stokes, fric, box, move_dt, global_id,
                                         global_id ? slot_gid : sorted_id, next, sort_rows, n_device, move_window,
                                         compact_in ? 1 : 0),
                          "k_jkr_substep_plane");
    }
#endif
    auto kernel = uniform_radius > 0 ? k_jkr_substep<true> : k_jkr_substep<false>;
    return check_cuda(launch_chained(kernel, ceil_div(n, kJkrThreads), kJkrThreads, as_stream(stream), chained, bin_start,
                                     sorted_id, reinterpret_cast<const double4 *>(sorted_pos), sorted_bin, (int)n, stc,
                                     Radius(distance), reach * reach, BandConst::make(reach * 1.0000001), old,
                                     new_row_len, new_cols, (int)row_stride, max_len, motility_forces, locations, c,
                                     stokes, fric, box, move_dt, global_id, global_id ? slot_gid : sorted_id, next,
                                     sort_rows, prev_slot, row_keys, n_device, move_window),
                      "k_jkr_substep");
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
                               const int32_t *n_device, const double *move_window_device, void *stream) {
    return jkr_substep_impl(bin_start, sorted_id, sorted_pos, sorted_bin, n, dims_host, distance, reach,
                            old_row_start, old_row_stride, old_row_len, old_cols, new_row_len, new_cols, row_stride,
                            max_len, sort_rows & ABM_SUBSTEP_SORT_ROWS, motility_forces, locations, poisson, youngs,
                            adhesion_const, stokes, size_host, move_dt, global_id, nullptr, uniform_radius,
                            next_extent_host, next_origin_host, next_bin_counts, next_sort_workspace, stream, nullptr, 0,
                            false, n_device, move_window_device, (sort_rows & ABM_SUBSTEP_FLAT_Z) != 0);
}

// The whole substep loop of CellMethods.apply_forces (reference cell_methods.py:399-436) enqueued by one call:
// n_substeps x [counting sort at bin edge `reach` + fused substep].  Substep k reads the bonds written by
// substep k-1 (substep 0 reads old_*) and writes row buffer a (k even) or b (k odd); the final bonds are in
// buffer a when n_substeps is odd, else in b.  Every substep pre-bins the moved cells, so only the first sort
// of the loop runs the histogram kernel; the two offset tables alternate.
// events (may be null): 2 per substep, recorded around the substep kernel (abm_mechanics_loop_timed)
static int mechanics_loop_impl(const double *radii, double *locations, const double *motility_forces, int64_t n,
                               const double extent_host[3], const double *origin_host, const double size_host[3],
                               double distance, double reach, int n_substeps, const double *dts_host,
                               int32_t *bin_start_a, int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos,
                               int32_t *sorted_bin, int32_t *max_count, void *sort_workspace,
                               size_t sort_workspace_bytes, const int32_t *old_row_start, int64_t old_row_stride,
                               const int32_t *old_row_len, const int32_t *old_cols, int32_t *len_a,
                               int32_t *cols_a, int32_t *len_b, int32_t *cols_b, int64_t row_stride,
                               int32_t *max_len, const int32_t *global_id, double uniform_radius, double poisson,
                               double youngs, double adhesion_const, double stokes, int flags,
                               const int32_t *n_device, const double *move_window_device, void *const *events,
                               void *stream) {
    ABM_REQUIRE(n_substeps >= 0 && dts_host, "bad substep list");
    ABM_REQUIRE(!(events && (flags & ABM_LOOP_FP32)), "the per-launch stopwatch is for the fp64 loop");
    if (flags & ABM_LOOP_FP32)
        return mechanics_loop32(radii, locations, motility_forces, n, extent_host, origin_host, size_host, distance,
                                reach, n_substeps, dts_host, bin_start_a, bin_start_b, sorted_id, sorted_pos,
                                sorted_bin, max_count, sort_workspace, sort_workspace_bytes, old_row_start,
                                old_row_stride, old_row_len, old_cols, len_a, cols_a, len_b, cols_b, row_stride,
                                max_len, global_id, uniform_radius, poisson, youngs, adhesion_const, stokes, flags,
                                n_device, move_window_device, stream);
    const bool carry_in = (flags & ABM_LOOP_CARRY_IN) && n > 0 && n_substeps > 0;
    const bool carry_out = (flags & ABM_LOOP_CARRY_OUT) && n > 0;
    const bool chained = flags & ABM_LOOP_CHAINED;
    // a flat world's z is exactly 0 once a substep has clamped it into [0, size_z]: from the second substep on (or from
    // the first, when the cells come from the previous call's hand-over) the two-axis kernel takes over
    const bool plane_ok = !(flags & ABM_LOOP_NO_PLANE);
    // ... and in a flat world of equal radii the position records between such substeps are 16-byte {x, y} ones
    // (kSortCompactXY): every hand-over whose consumer is the two-axis kernel, i.e. every hand-over of the loop
    const bool compact = plane_ok && uniform_radius > 0 && size_host[2] == 0.0 && !(flags & ABM_LOOP_FP32);
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
        const bool last = k + 1 == n_substeps;
        const bool prebin = n > 0 && (!last || carry_out);        // hand the moved cells to the next sort
        // (the largest bin population is reported for the first sort of the loop only)
        const bool handed = k || carry_in;                        // this sort's cells come from a hand-over
        int rc = bin_sort_impl(locations, radii, n, extent_host, origin_host, reach,
                               (handed ? ABM_SORT_PREBINNED : 0) | (chained ? ABM_SORT_CHAINED : 0) |
                                   ((handed && compact) ? kSortCompactXY : 0), table,
                               sorted_id, sorted_pos, sorted_bin, k ? nullptr : max_count, sort_workspace,
                               sort_workspace_bytes,
                               prebin ? other : nullptr, global_id, stream, n_device);   // (idle table cleared on the way)
        if (rc) return rc;
        int32_t *nl = (k & 1) ? len_b : len_a, *nc = (k & 1) ? cols_b : cols_a;
        const int32_t *slot_gid = (global_id && n > 0) ? sort_ws_carve(sort_workspace, n, nbins).slot_gid : nullptr;
        const int32_t *prev_slot = nullptr;
        int row_keys = 0;
#if ABM_JKR_EXP & 16
        if (n > 0) {
            prev_slot = sort_ws_carve(sort_workspace, n, nbins).prev_slot;
            row_keys = (k > 0 ? 1 : 0) | (last ? 0 : 2);          // first reads / last writes rows keyed by cell id
        }
#endif
        if (events && n > 0) {
            rc = check_cuda(cudaEventRecord((cudaEvent_t)events[2 * k], st), "cudaEventRecord");
            if (rc) return rc;
        }
        rc = jkr_substep_impl(table, sorted_id, sorted_pos, sorted_bin, n, g.dims, distance, reach, cur_start,
                              cur_stride, cur_len, cur_cols, nl, nc, row_stride, max_len, k == n_substeps - 1,
                              motility_forces, last ? locations : nullptr, poisson, youngs, adhesion_const, stokes,
                              size_host, dts_host[k], global_id, slot_gid, uniform_radius,
                              prebin ? extent_host : nullptr, origin_host, prebin ? other : nullptr,
                              prebin ? sort_workspace : nullptr, stream, prev_slot, row_keys, chained, n_device,
                              move_window_device, plane_ok && (handed || (flags & ABM_LOOP_FLAT_Z)), handed && compact,
                              prebin && compact);
        if (rc) return rc;
        if (events && n > 0) {
            rc = check_cuda(cudaEventRecord((cudaEvent_t)events[2 * k + 1], st), "cudaEventRecord");
            if (rc) return rc;
        }
        cur_start = nullptr;
        cur_stride = row_stride;
        cur_len = nl;
        cur_cols = nc;
    }
    return 0;
}

#define ABM_LOOP_ARGS                                                                                                 \
    radii, locations, motility_forces, n, extent_host, origin_host, size_host, distance, reach, n_substeps, dts_host, \
        bin_start_a, bin_start_b, sorted_id, sorted_pos, sorted_bin, max_count, sort_workspace, sort_workspace_bytes, \
        old_row_start, old_row_stride, old_row_len, old_cols, len_a, cols_a, len_b, cols_b, row_stride, max_len,      \
        global_id, uniform_radius, poisson, youngs, adhesion_const, stokes, flags, n_device, move_window_device

ABM_EXPORT int abm_mechanics_loop(const double *radii, double *locations, const double *motility_forces, int64_t n,
                                  const double extent_host[3], const double *origin_host, const double size_host[3],
                                  double distance, double reach, int n_substeps, const double *dts_host,
                                  int32_t *bin_start_a, int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos,
                                  int32_t *sorted_bin, int32_t *max_count, void *sort_workspace,
                                  size_t sort_workspace_bytes, const int32_t *old_row_start, int64_t old_row_stride,
                                  const int32_t *old_row_len, const int32_t *old_cols, int32_t *len_a,
                                  int32_t *cols_a, int32_t *len_b, int32_t *cols_b, int64_t row_stride,
                                  int32_t *max_len, const int32_t *global_id, double uniform_radius, double poisson,
                                  double youngs, double adhesion_const, double stokes, int flags,
                                  const int32_t *n_device, const double *move_window_device, void *stream) {
    return mechanics_loop_impl(ABM_LOOP_ARGS, nullptr, stream);
}

ABM_EXPORT int abm_mechanics_loop_timed(const double *radii, double *locations, const double *motility_forces,
                                        int64_t n, const double extent_host[3], const double *origin_host,
                                        const double size_host[3], double distance, double reach, int n_substeps,
                                        const double *dts_host, int32_t *bin_start_a, int32_t *bin_start_b,
                                        int32_t *sorted_id, double *sorted_pos, int32_t *sorted_bin,
                                        int32_t *max_count, void *sort_workspace, size_t sort_workspace_bytes,
                                        const int32_t *old_row_start, int64_t old_row_stride,
                                        const int32_t *old_row_len, const int32_t *old_cols, int32_t *len_a,
                                        int32_t *cols_a, int32_t *len_b, int32_t *cols_b, int64_t row_stride,
                                        int32_t *max_len, const int32_t *global_id, double uniform_radius,
                                        double poisson, double youngs, double adhesion_const, double stokes, int flags,
                                        const int32_t *n_device, const double *move_window_device,
                                        void *const *substep_events_host, void *stream) {
    return mechanics_loop_impl(ABM_LOOP_ARGS, substep_events_host, stream);
}
