// rules.cu -- the per-cell rule methods of the model on the device: cell_division, cell_death,
// cell_pathway, cell_differentiate, cell_motility (reference cell_methods.py:11-340), so that a whole simulation step
// runs without the agent arrays ever visiting the host.  Randomness comes from the counter-based contract of
// include/abm_rng.h; cell_pathway reads its neighbours' FGF4 as of before the pass.  Both differ from the reference's
// serial Mersenne-Twister loop by design; the CPU restatement these kernels match bit for bit is
// oracle/abm_oracle.c orc_rule_* (see its header for what is and is not pinned).
//
// The colony lives in fixed-capacity arrays with its cell count in device memory (`n_device`): division appends the
// daughters behind the last cell, death compacts the arrays (stable, so array order stays id order).  Cells carry a
// persistent id (`gid`); JKR bond rows hold ids, so nothing has to be renumbered when cells die (the mechanics
// kernels take the id table, include/abm_b200.h `global_id`).  The neighbour graph of the step was searched BEFORE
// division and death, exactly as in the reference (cell_simulation.py:90-97): its rows move with their cells, its
// entries are translated through the compaction's index map (-1: the neighbour died; daughters have empty rows).
#include "abm_rng.h"
#include "scan.cuh"

namespace abm {

constexpr int kRuleThreads = 256;

struct Colony {          // == abm_colony_arrays
    double *loc, *radii, *mot;
    int *gid, *states, *GATA6, *NANOG, *FGF4, *FGFR, *ERK, *death, *diff, *div, *fds;
    int *bond_len, *bond_rows;
    int *nbr_start, *nbr_len;
};
typedef abm_rule_params Params;

// ---- cell_division (cell_methods.py:54-117) ------------------------------------------------------------------------
__global__ void __launch_bounds__(kRuleThreads)
k_division_mark(Colony c, int capacity, const int *__restrict__ n_dev, Params P, int *__restrict__ flags) {
    const int i = blockIdx.x * kRuleThreads + threadIdx.x;
    if (i > capacity) return;
    const int n = device_count(capacity, n_dev);
    int f = 0;
    if (i < n) {
        const int d = c.div[i] + abm_coin(abm_draw(P.seed, P.step, ABM_STREAM_DIV_COUNTER, (uint64_t)c.gid[i], 0));
        c.div[i] = d;
        if (c.states[i] == 0) f = d >= P.pluri_div_thresh;
        else if (d >= P.diff_div_thresh) f = c.nbr_len[i] < 6;
    }
    flags[i] = f;                                              // (flags[capacity] = 0: the scan's total lands there)
}

// pos = exclusive scan of the flags; the daughter of mother i becomes cell n + pos[i]
__global__ void __launch_bounds__(kRuleThreads)
k_division_apply(Colony c, int capacity, int stride, int *__restrict__ n_dev, int *__restrict__ next_gid, Params P,
                 const int *__restrict__ pos, int *__restrict__ needed) {
    const int i = blockIdx.x * kRuleThreads + threadIdx.x;
    const int n = device_count(capacity, n_dev);
    const int added = pos[capacity];
    if (i < n && pos[i + 1] != pos[i]) {
        const int d = n + pos[i];
        double v[3];
        abm_random_vector(P.seed, P.step, ABM_STREAM_DIV_VECTOR, (uint64_t)c.gid[i], P.three_d, v);
        const double scale = ABM_SUB(P.max_radius, P.min_radius);
        const double x[3] = {c.loc[3 * (long long)i], c.loc[3 * (long long)i + 1], c.loc[3 * (long long)i + 2]};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double dv = ABM_MUL(v[k], scale);
            c.loc[3 * (long long)i + k] = ABM_ADD(x[k], dv);              // mother  += division_position
            if (d < capacity) c.loc[3 * (long long)d + k] = ABM_SUB(x[k], dv);   // daughter -= division_position
        }
        c.div[i] = 0;
        if (d < capacity) {
            c.radii[d] = c.radii[i];
#pragma unroll
            for (int k = 0; k < 3; ++k) c.mot[3 * (long long)d + k] = c.mot[3 * (long long)i + k];
            c.gid[d] = *next_gid + pos[i];
            c.states[d] = c.states[i];
            c.GATA6[d] = c.GATA6[i];
            c.NANOG[d] = c.NANOG[i];
            c.FGF4[d] = c.FGF4[i];
            c.FGFR[d] = c.FGFR[i];
            c.ERK[d] = c.ERK[i];
            c.death[d] = c.death[i];
            c.diff[d] = c.diff[i];
            c.div[d] = 0;
            c.fds[d] = c.fds[i];
            c.bond_len[d] = 0;                                  // add_vertex(): a new vertex has no edges
            c.nbr_len[d] = 0;
            c.nbr_start[d] = 0;
        }
    }
    // the counters move once every thread of the grid has read them: last block to finish
    __shared__ bool last;
    __threadfence();
    __syncthreads();          // every warp of this block is done before thread 0 reports the block as finished
    if (threadIdx.x == 0) last = atomicAdd(&needed[1], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        needed[1] = 0;
        atomicMax(&needed[0], n + added);                       // rows needed (> capacity: daughters were dropped)
        *n_dev = min(n + added, capacity);
        *next_gid += added;
    }
}

// ---- cell_death (cell_methods.py:11-52) ------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRuleThreads)
k_death_mark(Colony c, int capacity, const int *__restrict__ n_dev, Params P, int *__restrict__ keep) {
    const int i = blockIdx.x * kRuleThreads + threadIdx.x;
    if (i > capacity) return;
    const int n = device_count(capacity, n_dev);
    int k = 0;
    if (i < n) {
        k = 1;
        if (c.states[i] == 0) {
            int d = c.nbr_len[i] < P.lonely_thresh ? c.death[i] + 1 : 0;
            c.death[i] = d;
            if (d >= P.death_thresh) k = 0;
        }
    }
    keep[i] = k;
}

// stable compaction of every per-cell array; index_map[old] = new index or -1
__global__ void __launch_bounds__(kRuleThreads)
k_compact(Colony a, Colony o, int capacity, int stride, const int *__restrict__ n_dev, const int *__restrict__ pos,
          int *__restrict__ index_map, int *__restrict__ n_out) {
    const int i = blockIdx.x * kRuleThreads + threadIdx.x;
    const int n = device_count(capacity, n_dev);
    if (i == 0) *n_out = pos[capacity];
    if (i >= capacity) return;
    if (i >= n) {
        index_map[i] = -1;
        return;
    }
    const bool kept = pos[i + 1] != pos[i];
    const int d = pos[i];
    index_map[i] = kept ? d : -1;
    if (!kept) return;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o.loc[3 * (long long)d + k] = a.loc[3 * (long long)i + k];
        o.mot[3 * (long long)d + k] = a.mot[3 * (long long)i + k];
    }
    o.radii[d] = a.radii[i];
    o.gid[d] = a.gid[i];
    o.states[d] = a.states[i];
    o.GATA6[d] = a.GATA6[i];
    o.NANOG[d] = a.NANOG[i];
    o.FGF4[d] = a.FGF4[i];
    o.FGFR[d] = a.FGFR[i];
    o.ERK[d] = a.ERK[i];
    o.death[d] = a.death[i];
    o.diff[d] = a.diff[i];
    o.div[d] = a.div[i];
    o.fds[d] = a.fds[i];
    o.nbr_start[d] = a.nbr_start[i];
    o.nbr_len[d] = a.nbr_len[i];
    const int len = min(a.bond_len[i], stride);
    o.bond_len[d] = len;
    const int *src = a.bond_rows + (long long)i * stride;
    int *dst = o.bond_rows + (long long)d * stride;
    for (int k = 0; k < len; ++k) dst[k] = src[k];
}

// ---- cell_pathway + cell_differentiate (cell_methods.py:176-244) ------------------------------------------------------
__global__ void __launch_bounds__(kRuleThreads)
k_pathway(Colony c, int capacity, const int *__restrict__ n_dev, const int *__restrict__ nbr_cols,
          const int *__restrict__ index_map, const int *__restrict__ fgf4_old, Params P) {
    const int i = blockIdx.x * kRuleThreads + threadIdx.x;
    const int n = device_count(capacity, n_dev);
    if (i >= n) return;
    int x4 = c.GATA6[i], x5 = c.NANOG[i];
    if (P.step >= P.dox_step) {
        const int start = c.nbr_start[i], len = c.nbr_len[i];
        int num = 1;
        for (int e = 0; e < len; ++e) num += index_map[nbr_cols[start + e]] >= 0;
        const double dnum = (double)num;
        const uint64_t me = (uint64_t)c.gid[i];
        double perceived = 0.0;
        for (int e = 0; e <= len; ++e) {                        // the neighbours in ascending order, then the cell
            const int j = e < len ? index_map[nbr_cols[start + e]] : i;
            if (j < 0) continue;
            uint64_t used;
            const double g = abm_gauss(P.seed, P.step, ABM_STREAM_PATHWAY, me, (uint64_t)c.gid[j] << 6, &used);
            perceived = ABM_ADD(perceived, ABM_MUL(ABM_ADD(1.0, g), ABM_DIV((double)fgf4_old[j], dnum)));
        }
        int x1 = (int)perceived;                                // int(): truncation toward zero
        x1 = x1 > P.field - 1 ? P.field - 1 : (x1 < 0 ? 0 : x1);
        const int fds = c.fds[i];
        if (fds % P.fds_thresh == 0) {
            const int x2 = c.FGFR[i], x3 = c.ERK[i];
            int n4, n5;
            if (P.field == 2) {
                c.FGF4[i] = x5;
                c.FGFR[i] = (1 + x5 + (x4 * x5)) % 2;
                c.ERK[i] = (x1 * x2) % 2;
                n4 = (x3 + x4 + (x3 * x4) + (x3 * x5) + (x4 * x5) + (x3 * x4 * x5)) % 2;
                n5 = (x5 + (x3 * x5) + (x4 * x5) + (x3 * x4 * x5)) % 2;
            } else {
                c.FGF4[i] = x5;
                c.FGFR[i] = (x1 * x4 * ((2 * x1 + 1) * (2 * x4 + 1) + x1 * x4)) % 3;
                c.ERK[i] = x2 % 3;
                n4 = ((x4 * x4) * (x5 + 1) + (x5 * x5) * (x4 + 1) + 2 * x5 + 1) % 3;
                n5 = (x5 * x5 + x5 * (x5 + 1) * (x3 * (2 * x4 * x4 + 2 * x3 + 1) + x4 * (2 * x3 * x3 + 2 * x4 + 1)) +
                      (2 * x3 * x3 + 1) * (2 * x4 * x4 + 1)) % 3;
            }
            x4 = n4;
            x5 = n5;
            c.GATA6[i] = x4;
        }
        c.fds[i] = fds + 1;
    }
    // cell_differentiate: own values only, after the pathway update of this cell
    if (x4 > x5 && c.states[i] == 0) {
        const int d = c.diff[i] + abm_coin(abm_draw(P.seed, P.step, ABM_STREAM_DIFF_COUNTER, (uint64_t)c.gid[i], 0));
        c.diff[i] = d;
        if (d >= P.pluri_to_diff) {
            c.states[i] = 1;
            x5 = 0;
        }
    }
    c.NANOG[i] = x5;
}

// ---- cell_motility (cell_methods.py:246-340) --------------------------------------------------------------------------
__global__ void __launch_bounds__(kRuleThreads)
k_motility(Colony c, int capacity, const int *__restrict__ n_dev, const int *__restrict__ nbr_cols,
           const int *__restrict__ index_map, Params P) {
    const int i = blockIdx.x * kRuleThreads + threadIdx.x;
    const int n = device_count(capacity, n_dev);
    if (i >= n) return;
    const int start = c.nbr_start[i], len = c.nbr_len[i];
    int alive = 0;
    for (int e = 0; e < len; ++e) alive += index_map[nbr_cols[start + e]] >= 0;
    if (alive >= 6) return;
    const int state = c.states[i], g6 = c.GATA6[i], ng = c.NANOG[i];
    int mode;                                   // 0: random only; 1: NANOG-high neighbours; 2: the cell's OWN state (:287)
    double sign = 0.0;
    if (state == 1) { mode = 1; sign = -0.8; }
    else if (g6 > ng) {
        if (P.guye_move) { mode = 2; sign = 0.8; }
        else { mode = 1; sign = -0.8; }
    } else if (g6 < ng) { mode = 1; sign = 0.8; }
    else mode = 0;
    int count = 0;
    double hold[3] = {0.0, 0.0, 0.0};
    const double x[3] = {c.loc[3 * (long long)i], c.loc[3 * (long long)i + 1], c.loc[3 * (long long)i + 2]};
    if (mode != 0)
        for (int e = 0; e < len; ++e) {
            const int j = index_map[nbr_cols[start + e]];
            if (j < 0) continue;
            const bool take = mode == 1 ? (c.NANOG[j] > c.GATA6[j]) : (state == 1);
            if (take) {
                ++count;
#pragma unroll
                for (int k = 0; k < 3; ++k) hold[k] = ABM_ADD(hold[k], ABM_SUB(c.loc[3 * (long long)j + k], x[k]));
            }
        }
    double rnd[3];
    abm_random_vector(P.seed, P.step, ABM_STREAM_MOTILITY, (uint64_t)c.gid[i], P.three_d, rnd);
    if (count > 0) {
        const double mag = ABM_SQRT(ABM_ADD(ABM_ADD(ABM_MUL(hold[0], hold[0]), ABM_MUL(hold[1], hold[1])),
                                            ABM_MUL(hold[2], hold[2])));
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double normal = mag == 0.0 ? 0.0 : ABM_DIV(hold[k], mag);
            const double f = ABM_MUL(ABM_ADD(ABM_MUL(normal, sign), ABM_MUL(rnd[k], 0.2)), P.motility_force);
            c.mot[3 * (long long)i + k] = ABM_ADD(c.mot[3 * (long long)i + k], f);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 3; ++k)
            c.mot[3 * (long long)i + k] = ABM_ADD(c.mot[3 * (long long)i + k], ABM_MUL(rnd[k], P.motility_force));
    }
}

static Colony as_colony(const abm_colony_arrays *a) {
    Colony c{a->locations, a->radii, a->motility_forces, a->gid, a->states, a->GATA6, a->NANOG, a->FGF4, a->FGFR,
             a->ERK, a->death_counters, a->diff_counters, a->div_counters, a->fds_counters, a->bond_len,
             a->bond_rows, a->nbr_start, a->nbr_len};
    return c;
}
static bool complete(const abm_colony_arrays *a) {
    return a && a->locations && a->radii && a->motility_forces && a->gid && a->states && a->GATA6 && a->NANOG &&
           a->FGF4 && a->FGFR && a->ERK && a->death_counters && a->diff_counters && a->div_counters &&
           a->fds_counters && a->bond_len && a->bond_rows && a->nbr_start && a->nbr_len;
}

}  // namespace abm

using namespace abm;

ABM_EXPORT size_t abm_rules_workspace_bytes(int64_t capacity) {
    if (capacity < 0) return 0;
    return (((size_t)capacity + 1 + 63) / 64 * 64 + (size_t)scan_blocks(capacity + 1) + 64) * 4 + 256;
}

ABM_EXPORT int abm_rule_division(const abm_colony_arrays *cells, int64_t capacity, int64_t bond_stride,
                                 int32_t *n_device, int32_t *next_gid_device, const abm_rule_params *params,
                                 int32_t *needed_device, void *workspace, size_t workspace_bytes, void *stream) {
    ABM_REQUIRE(capacity > 0 && capacity < 0x7ffffff0 && bond_stride > 0, "bad sizes");
    ABM_REQUIRE(complete(cells) && n_device && next_gid_device && params && needed_device && workspace, "null pointer");
    ABM_REQUIRE(workspace_bytes >= abm_rules_workspace_bytes(capacity), "workspace too small");
    cudaStream_t st = as_stream(stream);
    Colony c = as_colony(cells);
    int *flags = (int *)workspace;
    int *sums = flags + ((size_t)capacity + 1 + 63) / 64 * 64;
    const int blocks = ceil_div(capacity + 1, kRuleThreads);
    k_division_mark<<<blocks, kRuleThreads, 0, st>>>(c, (int)capacity, n_device, *params, flags);
    ABM_CHECK_LAUNCH("k_division_mark");
    int rc = exclusive_scan_inplace(flags, capacity + 1, sums, nullptr, nullptr, st);
    if (rc) return rc;
    k_division_apply<<<blocks, kRuleThreads, 0, st>>>(c, (int)capacity, (int)bond_stride, n_device, next_gid_device,
                                                      *params, flags, needed_device);
    ABM_CHECK_LAUNCH("k_division_apply");
    return 0;
}

ABM_EXPORT int abm_rule_death(const abm_colony_arrays *cells, const abm_colony_arrays *cells_out, int64_t capacity,
                              int64_t bond_stride, const int32_t *n_device, int32_t *n_device_out,
                              const abm_rule_params *params, int32_t *index_map, void *workspace,
                              size_t workspace_bytes, void *stream) {
    ABM_REQUIRE(capacity > 0 && capacity < 0x7ffffff0 && bond_stride > 0, "bad sizes");
    ABM_REQUIRE(complete(cells) && complete(cells_out) && n_device && n_device_out && params && index_map && workspace,
                "null pointer");
    ABM_REQUIRE(cells->locations != cells_out->locations && n_device != n_device_out, "compaction is out of place");
    ABM_REQUIRE(workspace_bytes >= abm_rules_workspace_bytes(capacity), "workspace too small");
    cudaStream_t st = as_stream(stream);
    int *keep = (int *)workspace;
    int *sums = keep + ((size_t)capacity + 1 + 63) / 64 * 64;
    const int blocks = ceil_div(capacity + 1, kRuleThreads);
    k_death_mark<<<blocks, kRuleThreads, 0, st>>>(as_colony(cells), (int)capacity, n_device, *params, keep);
    ABM_CHECK_LAUNCH("k_death_mark");
    int rc = exclusive_scan_inplace(keep, capacity + 1, sums, nullptr, nullptr, st);
    if (rc) return rc;
    k_compact<<<blocks, kRuleThreads, 0, st>>>(as_colony(cells), as_colony(cells_out), (int)capacity,
                                               (int)bond_stride, n_device, keep, index_map, n_device_out);
    ABM_CHECK_LAUNCH("k_compact");
    return 0;
}

ABM_EXPORT int abm_rule_update(const abm_colony_arrays *cells, int64_t capacity, const int32_t *n_device,
                               const int32_t *nbr_cols, const int32_t *index_map, const abm_rule_params *params,
                               int32_t *fgf4_scratch, void *stream) {
    ABM_REQUIRE(capacity > 0 && capacity < 0x7ffffff0, "bad sizes");
    ABM_REQUIRE(complete(cells) && n_device && nbr_cols && index_map && params && fgf4_scratch, "null pointer");
    cudaStream_t st = as_stream(stream);
    Colony c = as_colony(cells);
    int rc = check_cuda(cudaMemcpyAsync(fgf4_scratch, c.FGF4, (size_t)capacity * 4, cudaMemcpyDeviceToDevice, st),
                        "snapshot of FGF4");
    if (rc) return rc;
    const int blocks = ceil_div(capacity, kRuleThreads);
    k_pathway<<<blocks, kRuleThreads, 0, st>>>(c, (int)capacity, n_device, nbr_cols, index_map, fgf4_scratch, *params);
    ABM_CHECK_LAUNCH("k_pathway");
    k_motility<<<blocks, kRuleThreads, 0, st>>>(c, (int)capacity, n_device, nbr_cols, index_map, *params);
    ABM_CHECK_LAUNCH("k_motility");
    return 0;
}
