// jkr_plane.cuh -- the fused mechanics substep for a FLAT world (included by jkr.cu after its helper structs).
//
// The reference's 2-D colonies are 3-D arrays with size_z == 0 (cell_simulation.py:136, general.yaml "size"); every
// integration clamps z into [0, size_z] = {0} (cell_backend.py:165-170).  From the second substep of a mechanics loop
// on -- and from the first one when the loop continues the previous call's hand-over -- every z is therefore exactly
// +0.0, and the z terms of the reference's arithmetic are exact no-ops:  (dx*dx + dy*dy) + 0*0  has the bits of
// dx*dx + dy*dy, the z force is 0, the moved z is 0.  k_jkr_substep_plane computes the same bits as k_jkr_substep
// without carrying the third axis, which changes the kernel's structure, not only its instruction count:
//   * a staged candidate is 16 bytes {x, y} in fp64, the size of the generic kernel's fp32 {x,y,z,slot} entry, so the
//     tile holds the EXACT positions: the candidate walk decides on the exact squared distance (no fp32 filter, no
//     second pass), and the contact loop takes the partner's position from shared memory instead of gathering it
//     from L2 -- no global load is left on the dependent chain of the loop;
//   * what is parked per partner is its tile index; the candidates' ids are staged beside their positions (one more
//     coalesced load per tile entry), so the bond row and the hysteresis lookup take the partner's id from shared
//     memory too: k_jkr_substep's id gathers held 13 % of its stall samples (profiles/r02b_kernels_c3.md);
//   * 64 registers, 8 CTAs of 128 threads per SM instead of 7.
// Measured and dropped (same profile file): the cell's motility force and its previous bond row fetched ahead with
// cp.async (slower: 56 and 67 us against 51), the force loaded into registers after the walk, two partners per
// iteration of the contact loop, 9 or 10 CTAs per SM, a CTA of 256 threads, a 12-deep column (all within +-2 %).
// Results are bit-identical to the generic kernel (tests/test_gpu_plane.py compares the two on the same colony).
#pragma once

namespace abm {

#ifndef ABM_PLANE_MINBLOCKS
#define ABM_PLANE_MINBLOCKS 8
#endif
constexpr int kPlaneMinBlocks = ABM_PLANE_MINBLOCKS;

__device__ __forceinline__ double sumsq2_nofma(double dx, double dy) {
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

template <bool kUniform>
__global__ void __launch_bounds__(kJkrThreads, kPlaneMinBlocks)
k_jkr_substep_plane(const int *__restrict__ bin_start, const int *__restrict__ sorted_id,
                    const double4 *__restrict__ spos, const int *__restrict__ sorted_bin, int n, Stencil st, Radius rad,
                    double reach2, GraphView old, int *__restrict__ new_len, int *__restrict__ new_cols, int stride,
                    int *__restrict__ max_len, const double *__restrict__ mot, double *__restrict__ loc, JkrConst c,
                    double stokes, Friction fric, Box box, double dt, const int *__restrict__ gid_of,
                    const int *__restrict__ slot_pid, NextSort next, int sort_rows, const int *__restrict__ n_dev,
                    const double *__restrict__ move_window, int compact_in) {
    // column tid: parked tile indices, then (overwritten from the front) the ids of the new bond row
    __shared__ int column[kMaxHits][kJkrThreads];
    __shared__ double2 tiles[kJkrThreads / 32][kTileRows * kTileCap];
    __shared__ int tile_ids[kJkrThreads / 32][kTileRows * kTileCap];     // first-hop id (sorted_id / slot_pid) of every candidate
    pdl_launch_dependents();
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int s = blockIdx.x * kJkrThreads + tid;
    pdl_wait();
    n = device_count(n, n_dev);
    if (s - lane >= n) return;                                 // the whole warp is past the last cell
    const bool live = s < n;
    double2 *tile = tiles[tid >> 5];
    int *tid_ids = tile_ids[tid >> 5];
    // position records: 32-byte {x, y, z, r} rows of a stand-alone sort, or 16-byte {x, y} rows between the substeps of
    // a loop (kSortCompactXY: equal radii only); either way {x, y} are the first 16 bytes of record t
    const double2 *__restrict__ sxy = reinterpret_cast<const double2 *>(spos);
    const int rec_shift = compact_in ? 0 : 1;
    const int *first_hop = slot_pid ? slot_pid : sorted_id;
    const bool second_hop = !slot_pid && gid_of;
    // fixed-stride previous rows of >= 8 entries (stride % 4 == 0): the first 8 ids come as two 16-byte loads
    const bool wide = old.valid() && !old.row_start && (old.stride & 3) == 0 && old.stride >= 8;
    int i = 0, bx = 0, by = 0;
    double px = 0.0, py = 0.0, pr = 1.0;
    if (live) {
        i = sorted_id[s];
        const double2 p = sxy[(long long)s << rec_shift];
        px = p.x;
        py = p.y;
        pr = kUniform ? c.uni_r : spos[s].w;
        const int b = sorted_bin[s];
        by = st.magic0 ? (int)__umulhi((unsigned)b, st.magic0) : b / st.dims[0];
        bx = b - by * st.dims[0];
        // the motility force is used at the very end: pull its line into L2 now
        asm volatile("prefetch.global.L2 [%0];" ::"l"(mot + 3 * (long long)i));
    }

    // ---- stage the three stencil rows: exact {x, y} of every candidate of the warp's 32 cells ----------------------
    int first[kTileRows], nr[kTileRows], L[kTileRows];
    unsigned staged = 0;
    bool slow = false;
    {
        const int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
        int W[kTileRows];
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
            const int y = by + r - 1;
            int lo = 0, hi = 0;
            if (live && y >= 0 && y < st.dims[1]) {
                const int rowb = y * st.dims[0];
                lo = bin_start[rowb + x0];
                hi = bin_start[rowb + x1 + 1];
            }
            nr[r] = hi - lo;
            L[r] = __reduce_min_sync(kFullWarp, hi > lo ? lo : 0x7fffffff);
            const int H = __reduce_max_sync(kFullWarp, hi > lo ? hi : 0);
            W[r] = H - L[r];
            if (W[r] > 0 && W[r] <= kTileCap) staged |= 1u << r;
            else slow |= nr[r] > 0;
            first[r] = lo - L[r] + r * kTileCap;
        }
#pragma unroll
        for (int r = 0; r < kTileRows; ++r) {
            if (staged & (1u << r)) {
#pragma unroll
                for (int c0 = 0; c0 < kTileCap; c0 += 32) {
                    const int cidx = c0 + lane;
                    if (cidx < W[r]) {
                        tile[r * kTileCap + cidx] = sxy[(long long)(L[r] + cidx) << rec_shift];
                        tid_ids[r * kTileCap + cidx] = first_hop[L[r] + cidx];
                    }
                }
            }
        }
        __syncwarp();
    }

    // ---- phase 1: exact squared distances over the staged runs; partners within reach are parked -----------------
    const int self_e = kTileCap + (s - L[1]);                  // the cell itself sits in the middle row's run
    int nh = 0;
#pragma unroll
    for (int r = 0; r < kTileRows; ++r) {
        if (!((staged >> r) & 1u)) continue;                   // warp-uniform
        const int f = first[r], m = nr[r];
        for (int u = 0; u < m; ++u) {
            const double2 q = tile[f + u];
            const double s2 = sumsq2_nofma(px - q.x, py - q.y);
            if (s2 <= reach2 && f + u != self_e) {
                if (nh < kMaxHits) column[nh][tid] = f + u;
                ++nh;
            }
        }
    }
    if (!live) return;                                         // no warp-wide operation with a full mask below
    slow |= nh > kMaxHits;

    int *row = new_cols + (long long)i * stride;
    long long fxi = 0, fyi = 0;
    BondEval ev;
    if (!kUniform) ev.preset(c, pr);
    int cnt = 0;         // bonds recorded in the column
    int direct = 0;      // bonds of a crowded cell written straight to row[kMaxHits + direct] (slow path only)
    auto partner_id = [&](int t) {
        if (slot_pid) return slot_pid[t];
        int j = sorted_id[t];
        return gid_of ? gid_of[j] : j;
    };
    // contact model for one bond candidate at offset v = p - q, |v|^2 = s2; a surviving bond is recorded as `token`:
    // the partner's id (its first-hop id on the fast path of a colony with a global-id table: translated after the loop)
    auto bond = [&](double vx, double vy, double s2, double rq, int token) {
        if (!kUniform) ev.refresh(pr, rq, c);
        const double rsum = kUniform ? c.uni_r + c.uni_r : ev.rsum;
        const double inv_scale = kUniform ? c.uni_inv_scale : ev.inv_scale;
        const double force_scale = kUniform ? c.uni_force_scale : ev.force_scale;
        double inv_mag = rsqrt(s2);                            // +inf for coincident cells, handled below
        double mag = s2 > 0.0 ? s2 * inv_mag : 0.0;
        double d = (rsum - mag) * inv_scale;
        if (d > -0.360562) {
            double f = (-0.0204 * (d * d * d)) + (0.4942 * (d * d)) + (1.0801 * d) - 1.324;
            if (s2 > 0.0) {
                double w = f * force_scale * inv_mag * kForceUnitInv;
                fxi += __double2ll_rn(w * vx);
                fyi += __double2ll_rn(w * vy);
            }
            if (cnt < kMaxHits) {
                column[cnt++][tid] = token;
            } else {
                if (kMaxHits + direct < stride) row[kMaxHits + direct] = token;
                ++direct;
            }
        }
    };
    const long long os = old.valid() ? old.start(i) : 0;
    auto was_bonded_slow = [&](int j) {
        int olen = old.row_len[i];
        bool found = false;
        for (int k = 0; k < olen; ++k) found |= (old.cols[os + k] == j);
        return found;
    };

    if (!slow) {
        // ---- phase 2: the contact model over the parked partners, everything from shared memory --------------------
        int nsh = 0, sh0 = 0, sh1 = 0;                         // hysteresis-shell partners (tile indices)
        int e_next = nh > 0 ? column[0][tid] : self_e;
        double2 q_next = tile[e_next];
        for (int k = 0; k < nh; ++k) {
            const int e = e_next;
            const double2 q = q_next;
            if (k + 1 < nh) {
                e_next = column[k + 1][tid];
                q_next = tile[e_next];
            }
            const double vx = px - q.x, vy = py - q.y;
            const double s2 = sumsq2_nofma(vx, vy);
            const int rsel = e >> 6;                           // kTileCap == 64
            const int t = (rsel == 0 ? L[0] : (rsel == 1 ? L[1] : L[2])) + (e & (kTileCap - 1));
            if (rad.within(s2)) {
                double rq = pr;
                if (!kUniform) rq = spos[t].w;
                bond(vx, vy, s2, rq, tid_ids[e]);              // cnt <= k: the column is overwritten behind the read
            } else if (old.valid()) {                          // thin shell: only pre-existing bonds persist
                if (nsh == 0) sh0 = e;
                else if (nsh == 1) sh1 = e;
                else {                                         // rare third+ shell partner
                    const int j1 = tid_ids[e];
                    if (was_bonded_slow(second_hop ? gid_of[j1] : j1)) {
                        double rq = pr;
                        if (!kUniform) rq = spos[t].w;
                        bond(vx, vy, s2, rq, j1);
                    }
                }
                ++nsh;
            }
        }
        // ---- phase 3: hysteresis -- shell partners that are in the previous row ------------------------------------
        if (nsh > 0) {
            const int olen = old.row_len[i];
            int4 oa = make_int4(-1, -1, -1, -1), ob = oa;
            if (wide) {
                oa = *reinterpret_cast<const int4 *>(old.cols + os);
                ob = *reinterpret_cast<const int4 *>(old.cols + os + 4);
            }
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
                const int e = k ? sh1 : sh0;
                const int rsel = e >> 6;
                const int t = (rsel == 0 ? L[0] : (rsel == 1 ? L[1] : L[2])) + (e & (kTileCap - 1));
                const int j1 = tid_ids[e];
                if (in_old_row(second_hop ? gid_of[j1] : j1)) {
                    const double2 q = tile[e];
                    const double vx = px - q.x, vy = py - q.y;
                    double rq = pr;
                    if (!kUniform) rq = spos[t].w;
                    bond(vx, vy, sumsq2_nofma(vx, vy), rq, j1);
                }
            }
        }
        // a colony with a global-id table but no per-slot copy of it (stand-alone abm_jkr_substep): the rows hold global
        // ids, one more hop (independent gathers, four in flight)
        if (second_hop) {
            auto hop = [&](int a) { return gid_of[a]; };
            for (int k = 0; k < cnt; k += 4) {
                const int a0 = column[k][tid];
                const int a1 = k + 1 < cnt ? column[k + 1][tid] : a0;
                const int a2 = k + 2 < cnt ? column[k + 2][tid] : a0;
                const int a3 = k + 3 < cnt ? column[k + 3][tid] : a0;
                const int j0 = hop(a0), j1 = hop(a1), j2 = hop(a2), j3 = hop(a3);
                column[k][tid] = j0;
                if (k + 1 < cnt) column[k + 1][tid] = j1;
                if (k + 2 < cnt) column[k + 2][tid] = j2;
                if (k + 3 < cnt) column[k + 3][tid] = j3;
            }
        }
    } else {
        // ---- slow path: crowded cell or candidate run longer than the tile -------------------------------------
        const int x0 = max(bx - 1, 0), x1 = min(bx + 1, st.dims[0] - 1);
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = by + dy;
            if (y < 0 || y >= st.dims[1]) continue;
            const int rowb = y * st.dims[0];
            const int lo = bin_start[rowb + x0], hi = bin_start[rowb + x1 + 1];
            for (int t = lo; t < hi; ++t) {
                if (t == s) continue;
                const double2 q = sxy[(long long)t << rec_shift];
                const double vx = px - q.x, vy = py - q.y;
                const double s2 = sumsq2_nofma(vx, vy);
                if (s2 > reach2) continue;
                const int j = partner_id(t);
                if (rad.within(s2) || (old.valid() && was_bonded_slow(j)))
                    bond(vx, vy, s2, kUniform ? pr : spos[t].w, j);
            }
        }
    }

    // ---- phase 4: the bond row (as in k_jkr_substep) ------------------------------------------------------------------
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
    new_len[i] = kept;
    if (total > stride) atomicMax(max_len, total);            // overflow report (rare): the host doubles the stride
    // integration (reference cell_backend.py:155-170); z stays +0.0: clamped into [0, size_z] = {0}
    double friction = fric.uni, inv_f = fric.uni_inv;
    if (!kUniform) {
        friction = stokes_friction(pr, stokes);
        inv_f = __drcp_rn(friction);
    }
    const double mx = mot[3 * (long long)i + 0], my = mot[3 * (long long)i + 1];
    const double nx = stokes_move_y(px, mx, (double)fxi * kForceUnit, friction, inv_f, dt, box.size[0]);
    const double ny = stokes_move_y(py, my, (double)fyi * kForceUnit, friction, inv_f, dt, box.size[1]);
    record_move(max_len, move_window, px, (float)((nx - px) * (nx - px) + (ny - py) * (ny - py)));
    if (loc) {
        loc[3 * (long long)i + 0] = nx;
        loc[3 * (long long)i + 1] = ny;
        loc[3 * (long long)i + 2] = 0.0;
    }
    if (next.counts) {                                        // hand-over to the next substep's sort, by slot
        const BinGeom &g = next.geom;
        const int nbx = clampi(__double2int_rd((nx - g.origin[0]) * g.inv_distance) + 1, 0, g.dims[0] - 1);
        const int nby = clampi(__double2int_rd((ny - g.origin[1]) * g.inv_distance) + 1, 0, g.dims[1] - 1);
        const int b = nby * g.dims[0] + nbx;
        if (next.compact) reinterpret_cast<double2 *>(next.next_pos)[s] = make_double2(nx, ny);
        else next.next_pos[s] = make_double4(nx, ny, 0.0, pr);
        next.cell_bin[s] = b;
        next.prev_id[s] = i;
        atomicAdd(&next.counts[b], 1);
        const int chunk = b / kScanChunk;
        const unsigned peers = __match_any_sync(__activemask(), chunk);
        if (lane == __ffs(peers) - 1) atomicAdd(&next.chunk_sums[chunk], __popc(peers));
    }
}

}  // namespace abm
