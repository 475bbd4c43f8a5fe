// sortws.cuh -- layout of the counting-sort scratch buffer, shared by sort.cu and the fused mechanics kernel
// (which can pre-bin the cells for the NEXT sort while it integrates them).
#pragma once
#include "scan.cuh"

namespace abm {

// internal flag of bin_sort_impl (not part of the public ABI; above every ABM_SORT_* bit): with ABM_SORT_PREBINNED, the
// hand-over in the workspace (next_pos) AND the sorted_pos this sort writes are 16-byte {x, y} records instead of
// 32-byte {x, y, z, r} ones -- the layout the two-axis substep kernel of a flat world with equal radii reads and
// writes (z is 0, r is a kernel parameter there): 40 % less traffic through placement and hand-over.
constexpr int kSortCompactXY = 1 << 16;

struct SortWs {
    int *cursor;      // nbins+1 : copy of the offset table; the placement kernels claim slots from it atomically
    int *cell_bin;    // n   : linear bin of cell i (ABM_SORT_PREBINNED: of the cell in PREVIOUS slot s)
    int *tmp_id;      // n   : unstable placement (stable mode); ABM_SORT_PREBINNED: id of the cell in previous slot s
    int *block_sums;  // scan scratch
    int *next_sums;   // ABM_SORT_PREBINNED: per-scan-chunk sums of the histogram, accumulated by the mechanics kernel
                      // (the placement kernel of every sort leaves them zeroed for the next one)
    double4 *next_pos;// n   : ABM_SORT_PREBINNED: moved {x,y,z,r} of the cell in previous slot s
    int *slot_gid;    // n   : global id of the cell in slot s (mechanics loop of a slab-decomposed colony)
    int *prev_slot;   // n   : ABM_SORT_PREBINNED: the slot the cell in slot s had in the previous sort (experiment
                      //       ABM_JKR_EXP & 16: bond rows keyed by slot between the substeps of a loop)
};

inline size_t sort_align(size_t v) { return (v + 255) & ~size_t(255); }

inline size_t sort_ws_bytes(int64_t n, int64_t nbins) {
    return 2 * sort_align((size_t)n * 4) + sort_align((size_t)(nbins + 1) * 4) +
           2 * sort_align((size_t)scan_blocks(nbins + 1) * 4 + 4) + sort_align((size_t)n * 32) + 2 * sort_align((size_t)n * 4) + 512;
}

inline SortWs sort_ws_carve(void *ws, int64_t n, int64_t nbins) {
    char *p = (char *)ws;
    SortWs w;
    w.cell_bin = (int *)p;   p += sort_align((size_t)n * 4);
    w.tmp_id = (int *)p;     p += sort_align((size_t)n * 4);
    w.cursor = (int *)p;     p += sort_align((size_t)(nbins + 1) * 4);
    w.block_sums = (int *)p; p += sort_align((size_t)scan_blocks(nbins + 1) * 4 + 4);
    w.next_sums = (int *)p;  p += sort_align((size_t)scan_blocks(nbins + 1) * 4 + 4);
    p = (char *)(((uintptr_t)p + 255) & ~uintptr_t(255));    // the caller's buffer need not be aligned
    w.next_pos = (double4 *)p; p += sort_align((size_t)n * 32);
    w.slot_gid = (int *)p;   p += sort_align((size_t)n * 4);
    w.prev_slot = (int *)p;
    return w;
}

// Bin lattice of a sort.  A flat world (extent_z == 0, the reference's 2-D runs) gets ONE z layer instead of the
// reference's three (two of which are always empty): every cell is a stencil candidate of its (x,y) neighbours
// whatever its z, and the distance predicate is exact, so the result is unchanged and the offset table is 3x
// shorter.  ABM_SORT_REFERENCE keeps the reference's lattice (what assign_bins() returns).
inline int make_bin_geom(const double extent[3], const double *origin, double distance, int flags, BinGeom &g,
                         int64_t &nbins) {
    if (!(distance > 0)) return -1;
    g.distance = distance;
    g.inv_distance = 1.0 / distance;
    nbins = 1;
    for (int k = 0; k < 3; ++k) {
        g.origin[k] = origin ? origin[k] : 0.0;
        g.dims[k] = (int)ceil(extent[k] / distance) + 3;
        if (k == 2 && extent[2] == 0.0 && !(flags & ABM_SORT_REFERENCE)) g.dims[k] = 1;
        nbins *= g.dims[k];
    }
    return nbins < (int64_t)0x7fffffff ? 0 : -1;
}

// abm_bin_sort with one rider (sort.cu)
int bin_sort_impl(const double *locations, const double *radii, int64_t n, const double size_host[3],
                  const double *origin_host, double distance, int flags, int32_t *bin_start, int32_t *sorted_id,
                  double *sorted_pos, int32_t *sorted_bin, int32_t *max_count, void *workspace,
                  size_t workspace_bytes, int32_t *zero_table, const int32_t *global_id, void *stream,
                  const int32_t *n_device = nullptr);

// the substep loop in fp32 on bin-relative coordinates (jkr32.cu); same arguments as abm_mechanics_loop
int mechanics_loop32(const double *radii, double *locations, const double *motility_forces, int64_t n,
                     const double extent_host[3], const double *origin_host, const double size_host[3],
                     double distance, double reach, int n_substeps, const double *dts_host, int32_t *bin_start_a,
                     int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos, int32_t *sorted_bin,
                     int32_t *max_count, void *sort_workspace, size_t sort_workspace_bytes,
                     const int32_t *old_row_start, int64_t old_row_stride, const int32_t *old_row_len,
                     const int32_t *old_cols, int32_t *len_a, int32_t *cols_a, int32_t *len_b, int32_t *cols_b,
                     int64_t row_stride, int32_t *max_len, const int32_t *global_id, double uniform_radius,
                     double poisson, double youngs, double adhesion_const, double stokes, int flags,
                     const int32_t *n_device, const double *move_window, void *stream);

__device__ __forceinline__ int cell_bin_of(double x, double y, double z, const BinGeom &g) {
    return lin_bin(bin_coord_fast(x - g.origin[0], g.distance, g.inv_distance, g.dims[0]),
                   bin_coord_fast(y - g.origin[1], g.distance, g.inv_distance, g.dims[1]),
                   bin_coord_fast(z - g.origin[2], g.distance, g.inv_distance, g.dims[2]), g.dims);
}

// Bin of a moved cell for the NEXT sort of the mechanics loop, without the exact-floor correction: the plain
// floor((x - origin) * (1/d)) can be off by one bin only for a coordinate within ~1e-12 um of a bin boundary.  The
// mechanics bins are `reach` = (largest bond distance) * (1 + 1e-9) wide, i.e. ~1e-8 um wider than any pair that has
// to meet in adjacent bins, so every such pair still does -- the partition need not be the reference's, only
// consistent within one sort.  (abm_bin_sort itself keeps the exact floor: there the bin edge may equal the search
// radius and the reference-shaped outputs are defined by it.)
__device__ __forceinline__ int cell_bin_of_loose(double x, double y, double z, const BinGeom &g) {
    int bx = clampi(__double2int_rd((x - g.origin[0]) * g.inv_distance) + 1, 0, g.dims[0] - 1);
    int by = clampi(__double2int_rd((y - g.origin[1]) * g.inv_distance) + 1, 0, g.dims[1] - 1);
    int bz = clampi(__double2int_rd((z - g.origin[2]) * g.inv_distance) + 1, 0, g.dims[2] - 1);
    return lin_bin(bx, by, bz, g.dims);
}

}  // namespace abm
