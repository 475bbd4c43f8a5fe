/*
 * abm_b200.h -- C-ABI of libabm_b200.so: the B200 (sm_100a) implementation of the per-step
 * cell-update hot path of Python-hiPSC-ABM.
 *
 * Every entry point replaces one of the reference's JIT functions or the numpy/igraph host work
 * around it (file:line under /root/reference/ABM is cited per function).  The reference binds this
 * path through Numba; the replacement binding is ctypes.
 *
 * Conventions
 *   - All pointers are DEVICE pointers (e.g. torch tensor.data_ptr()) unless the name ends in _host.
 *   - The caller owns every buffer, including workspaces; the library allocates nothing persistent.
 *   - Every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises,
 *     except the *_sync helpers that must hand a count back to the host.
 *   - Return value: 0 ok; <0 invalid argument (-1) or -(1000+cudaError_t).  Nothing throws across the ABI;
 *     abm_last_error() describes the last failure.
 *   - Capacity overflow is NOT a return code (a call never waits for the device, so it cannot know): a kernel
 *     that runs out of room writes what fits, and records the size it needed in a device counter the caller
 *     passed in (`cursor` of abm_neighbor_search, `max_len` of abm_jkr_substep / abm_mechanics_loop, `counts` of
 *     abm_slab_pack).  The caller reads the counter when it next synchronises (abm_read_i32_sync), and if it
 *     exceeds the capacity it grows the buffer and calls again -- the reference's grow-and-retry loops
 *     (simulation.py:217-221, 277-281) with the check moved off the critical path.
 *   - Device-side cell counts: every per-cell entry point takes `n` (int64, on the host: sizes the launch and the
 *     buffers) and `n_device` (may be NULL).  If n_device is given, the number of cells is min(n, *n_device), read
 *     by the kernels themselves: a pipeline whose cell count is decided on the device (cells migrating between
 *     slabs) enqueues -- or replays as a CUDA graph -- whole steps without bringing the count to the host.
 *   - No process-wide mutable state: every option is an argument of the call it affects (flags below); the only
 *     things the library caches are per-device kernel attributes and the per-thread text of the last error.
 *   - Cell ids, bin ids, row offsets are int32; positions/forces/fields are float64 exactly as the
 *     reference holds them (simulation.py:136-149 dtypes).
 *   - Positions are (N,3) row-major float64 ("AoS", the reference's `locations`), radii (N,) float64.
 */
#ifndef ABM_B200_H
#define ABM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABM_B200_ABI_VERSION 2

/* flags of abm_bin_dims_host / abm_bin_sort */
#define ABM_SORT_STABLE 1     /* ascending cell id inside a bin (the reference's fill order) */
#define ABM_SORT_REFERENCE 2  /* keep the reference's 3 z-layers in a flat (size_z == 0) world */
#define ABM_LOOP_CARRY_OUT 1   /* abm_mechanics_loop: the last substep also hands its cells to the NEXT call's first sort */
#define ABM_LOOP_CARRY_IN 2    /* abm_mechanics_loop: the previous call did so (same n, lattice, workspace, uniform_radius,
                                  ABM_LOOP_FP32 / ABM_LOOP_NO_PLANE flags -- they decide the layout of the handed-over
                                  records; its idle offset table is passed as bin_start_a): the first sort needs no
                                  histogram pass */
#define ABM_SORT_PREBINNED 4  /* histogram + per-cell bin already written by abm_jkr_substep */
#define ABM_SORT_CHAINED 8    /* abm_bin_sort(ABM_SORT_PREBINNED): launch its kernels with programmatic stream
                                 serialization (see ABM_LOOP_CHAINED) */
#define ABM_LOOP_FP32 8       /* abm_mechanics_loop: the dtype flag of SURVEY 8(b).  The substeps run in SINGLE precision on
                                 bin-relative coordinates (position = bin + fp32 offset inside the bin: one ulp is
                                 1e-6 um whatever the colony size), forces and positions within rel 1e-5 of the fp64
                                 path per step; clamp and re-binning stay exact (fp64).  Inputs and outputs keep their
                                 fp64 types.  Bond sets may differ from the fp64 path for pairs within ~1e-6 um of a
                                 threshold.  (csrc/jkr32.cu) */
#define ABM_SUBSTEP_SORT_ROWS 1 /* bits of abm_jkr_substep's sort_rows argument: ascending bond rows */
#define ABM_SUBSTEP_FLAT_Z 2    /* ... the caller vouches that size[2] == 0 AND every z in sorted_pos is exactly +0.0 (true for
                                   positions any substep of a flat world has produced): the two-axis kernel
                                   k_jkr_substep_plane runs, which computes the same bits (csrc/jkr_plane.cuh) */
#define ABM_LOOP_NO_PLANE 16  /* abm_mechanics_loop: keep the three-axis kernel also in a flat (size[2] == 0) world.  By
                                 default the substeps whose input z is known to be exactly 0 (all but the first of a
                                 call without ABM_LOOP_CARRY_IN) run k_jkr_substep_plane, which computes the same bits
                                 (csrc/jkr_plane.cuh); the flag exists for the test that compares the two */
#define ABM_LOOP_FLAT_Z 32    /* abm_mechanics_loop in a flat (size[2] == 0) world: the caller vouches that every z in `locations` is
                                 exactly +0.0 at entry (true for anything a previous loop has produced), so the first
                                 substep runs the two-axis kernel as well */
#define ABM_LOOP_CHAINED 4    /* abm_mechanics_loop: launch the kernels of the loop with programmatic stream
                                 serialization, so that a kernel's CTAs become resident while the last wave of its
                                 predecessor drains (they wait for its completion before touching memory) */

/* ---- library ----------------------------------------------------------------------------------- */
int abm_abi_version(void);          /* == ABM_B200_ABI_VERSION */
int abm_compiled_arch(void);        /* 100 : built for sm_100a only */
const char *abm_last_error(void);   /* thread-local text of the last non-zero return */

/* fills sm_count / cc_major / cc_minor of the current device; <0 if there is no CUDA device */
int abm_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ---- spatial binning: counting sort ------------------------------------------------------------
 * Replaces Simulation.assign_bins + assign_bins_jit (simulation.py:190-223, backend.py:31-50).
 * dims[k] = ceil(size[k]/distance)+3 and bin = floor_divide(x,distance)+1 exactly as the reference;
 * coordinates are clamped into [0,dims-1] (the reference wraps negative indices to an always-empty
 * bin; the edge sets agree, DESIGN.md "bins").  Linear bin id = (bz*dims[1]+by)*dims[0]+bx.
 * Unless ABM_SORT_REFERENCE is set, a flat world (size[2] == 0) gets one z layer instead of three (the other
 * two are always empty; candidates are a superset and the predicate is exact, so results are unchanged).
 * Unless ABM_SORT_STABLE is set, the order of the cells INSIDE a bin is unspecified (every consumer in this
 * library orders rows by cell id and sums forces in fixed point, so results do not depend on it).    */
int abm_bin_dims_host(const double size_host[3], double distance, int flags, int32_t dims_host[3]);

/* bytes of scratch abm_bin_sort needs for n cells and nbins bins */
size_t abm_bin_sort_workspace_bytes(int64_t n, int64_t nbins);

/* Outputs:
 *   bin_start  (nbins+1) int32  exclusive prefix of per-bin counts, bin_start[nbins] = n
 *   sorted_id  (n)       int32  sorted slot -> cell id; with ABM_SORT_STABLE ascending id inside a bin (the
 *                               reference's serial fill order, backend.py:36-48)
 *   sorted_pos (n,4)     f64    {x,y,z,radius} of the cell in that slot (32-byte rows)
 *   sorted_bin (n)       int32  linear bin of the slot
 *   max_count  (1)       int32  largest bin population (what drives the reference's max_agents)
 * With ABM_SORT_PREBINNED `locations` and `radii` are not read (may be NULL): the cells come from the workspace,
 * where the preceding abm_jkr_substep left them in the previous sort's slot order.
 * origin_host (may be NULL = 0,0,0) shifts the bin lattice only: bin = floor_divide(x-origin,distance)+1 over
 * an extent of size_host.  A slab-decomposed colony bins its slab (+halo) this way while positions, the
 * distance predicate and the [0,size] clamp stay in global coordinates.                              */
int abm_bin_sort(const double *locations, const double *radii, int64_t n, const double size_host[3],
                 const double *origin_host, double distance, int flags, int32_t *bin_start, int32_t *sorted_id, double *sorted_pos,
                 int32_t *sorted_bin, int32_t *max_count, void *workspace, size_t workspace_bytes,
                 const int32_t *n_device, void *stream);

/* Reference-shaped dense outputs of assign_bins (simulation.py:206-213) from a finished sort:
 *   bins (dims0,dims1,dims2,max_agents) int64 [reference index order x,y,z], bins_help (dims0,dims1,dims2)
 *   int64, bin_locations (n,3) int64.  bins/bins_help must be zero-filled by the caller.           */
int abm_bins_dense(const double *locations, int64_t n, const double size_host[3], double distance,
                   const int32_t *bin_start, const int32_t *sorted_id, int64_t max_agents, int64_t *bins,
                   int64_t *bins_help, int64_t *bin_locations, void *stream);

/* ---- fixed-radius neighbour search -> CSR -------------------------------------------------------
 * Replaces get_neighbors_cpu/gpu + the edges[if_edge] compaction + graph.add_edges (+ simplify when
 * clear=False) (backend.py:64-160, simulation.py:225-288).
 * Graph format ("row CSR"): for cell i its neighbours are cols[row_start[i] .. row_start[i]+row_len[i]),
 * both directions stored (j in row i <=> i in row j), ascending.  Rows are packed without gaps but in
 * no particular order (warp-aggregated reservation), so row_start is explicit.  Wherever a graph is an
 * INPUT, row_start may be NULL, meaning fixed-stride rows: row i starts at i*row_stride.
 * An edge exists iff sqrt(dx*dx+dy*dy+dz*dz) <= distance evaluated in fp64 without FMA (bit-identical
 * to the oracle).  If old_* is non-NULL the result is the union with that graph (clear=False).
 *   cursor (1) int32 : zeroed by the call, ends as the number of column entries needed (= 2*edges).
 * Returns 0; the caller reads *cursor (abm_read_i32_sync) and, if it exceeds cols_capacity, grows
 * cols and calls again (rows are only valid when cursor <= cols_capacity).                         */
int abm_neighbor_search(const int32_t *bin_start, const int32_t *sorted_id, const double *sorted_pos,
                        const int32_t *sorted_bin, int64_t n, const int32_t dims_host[3], double distance,
                        const double *locations, const int32_t *old_row_start, int64_t old_row_stride,
                        const int32_t *old_row_len, const int32_t *old_cols, int32_t *row_start,
                        int32_t *row_len, int32_t *cols, int64_t cols_capacity, int32_t *cursor,
                        const int32_t *n_device, void *stream);

/* synchronises `stream` and returns *device_value through value_host */
int abm_read_i32_sync(const int32_t *device_value, int32_t *value_host, void *stream);

/* (i<j) edge list of a row-CSR graph in canonical order (sorted by i, then j), as the reference's
 * get_edgelist() after simplify() (cell_methods.py:352).  scan_ws: abm_graph_edge_list_workspace_bytes(n).
 * edges (E,2) int64.  n_edges (1) int32 receives E.  Two launches, no host sync.                  */
size_t abm_graph_edge_list_workspace_bytes(int64_t n);
int abm_graph_edge_list(const int32_t *row_start, int64_t row_stride, const int32_t *row_len,
                        const int32_t *cols, int64_t n, int32_t *scan_ws, size_t scan_ws_bytes, int64_t *edges,
                        int64_t edges_capacity, int32_t *n_edges, void *stream);

/* ---- JKR forces over graph rows ------------------------------------------------------------------
 * Replaces jkr_forces_cpu/gpu + jkr_graph.delete_edges (cell_backend.py:6-115, cell_methods.py:342-384).
 * Row-owner accumulation (each cell sums its own bonds in ascending partner id = the serial
 * reference's += order), so no atomics and run-to-run identical.  Broken bonds (d <= -0.360562) are
 * removed from both rows in place (row_len shrinks).  jkr_forces (N,3) is ACCUMULATED into.       */
int abm_jkr_forces(const int32_t *row_start, int64_t row_stride, int32_t *row_len, int32_t *cols,
                   const double *locations, const double *radii, double *jkr_forces, int64_t n, double poisson,
                   double youngs, double adhesion_const, void *stream);

/* Replaces apply_forces_cpu/gpu (cell_backend.py:118-172): Stokes velocity, Euler step, clamp to
 * [0,size] per axis, same operation order as the reference (bit-identical for identical inputs).
 * If jkr_forces_to_zero is non-NULL that (N,3) array is zeroed afterwards (cell_methods.py:436).   */
int abm_apply_forces(const double *jkr_forces, const double *motility_forces, double *locations,
                     const double *radii, int64_t n, double stokes, const double size_host[3], double move_dt,
                     double *jkr_forces_to_zero, void *stream);

/* ---- fused mechanics substep ---------------------------------------------------------------------
 * One launch = get_neighbors("jkr_graph", distance, clear=False) + calculate_jkr() + apply_forces
 * integration for one of the 11 substeps of CellMethods.apply_forces (cell_methods.py:399-436).
 * Input: a sort of the CURRENT locations made with bin edge >= abm_jkr_reach() (so every bond that
 * can survive lies inside the 27-bin stencil), the bond graph from the previous substep (old_*; rows
 * may be unsorted) and the motility forces.  Output: the new bond graph in fixed-stride rows
 * (row i at new_cols[i*row_stride], length new_row_len[i], ascending if sort_rows & 1) and the moved locations
 * (written in place; the kernel only reads positions from sorted_pos).  The force on a cell is summed in
 * 2^-76 N fixed point (integer adds are associative), so it does not depend on the order inside the bins.
 * A bond (i,j) exists afterwards iff (dist <= distance or it existed before) and d > -0.360562.
 *   global_id        : NULL, or (n) int32 ids to write into / look up in the bond rows instead of the cell
 *                      index (a slab-decomposed colony keeps its bonds by global id; rows still belong to the
 *                      local cell index).
 *   uniform_radius   : the common radius if all radii are equal (the contact constants are then evaluated once
 *                      on the host), else 0.
 *   next_bin_counts  : NULL, or a zeroed (nbins+1) table (not bin_start): the kernel then also bins every moved
 *                      cell on the lattice (next_origin, next_extent, edge `reach`) and leaves histogram,
 *                      bin, moved position and id of every cell (in the current slot order) and the histogram's
 *                      per-chunk sums where abm_bin_sort(ABM_SORT_PREBINNED, same workspace, SAME lattice as the
 *                      sort that produced the current order) expects them.  `locations` may then be NULL (the
 *                      moved cells are only handed to the next sort).
 *   reach            : bin edge of the sort, >= abm_jkr_reach(distance, max(radii), ...).
 *   max_len (2) int32: [0] is raised (atomicMax) by rows that needed more than row_stride entries; a value
 *   > row_stride afterwards means those rows were truncated: grow row_stride and redo the substep
 *   from the saved locations.  [1] is raised to the bit pattern of the largest squared displacement (float, um^2)
 *   of a cell in one substep -- what the slab decomposition's halo depth has to be checked against.  The caller
 *   zeroes both.
 *   move_window_device (2) double, may be NULL: only cells with x < [0] or x > [1] count for max_len[1] (a slab's
 *   ghost zones: the only place a wrong cell can be while the halo holds); NULL = every cell.        */
double abm_jkr_reach(double distance, double max_radius, double poisson, double youngs, double adhesion_const);

int abm_jkr_substep(const int32_t *bin_start, const int32_t *sorted_id, const double *sorted_pos,
                    const int32_t *sorted_bin, int64_t n, const int32_t dims_host[3], double distance,
                    double reach, const int32_t *old_row_start, int64_t old_row_stride,
                    const int32_t *old_row_len, const int32_t *old_cols, int32_t *new_row_len, int32_t *new_cols,
                    int64_t row_stride, int32_t *max_len, int sort_rows, const double *motility_forces,
                    double *locations,
                    double poisson, double youngs, double adhesion_const, double stokes,
                    const double size_host[3], double move_dt, const int32_t *global_id, double uniform_radius,
                    const double *next_extent_host, const double *next_origin_host, int32_t *next_bin_counts,
                    void *next_sort_workspace, const int32_t *n_device, const double *move_window_device,
                    void *stream);

/* The whole substep loop of CellMethods.apply_forces (cell_methods.py:399-436) enqueued by ONE call:
 * n_substeps x [abm_bin_sort at bin edge `reach` over (origin, extent) + abm_jkr_substep with dts_host[k]].
 * Substep 0 reads the bonds old_*; substep k writes row buffer a (k even) or b (k odd), so the final bonds are in
 * buffer a when n_substeps is odd, else in b (the last substep writes ascending rows).  Every substep pre-bins the moved cells for the next sort
 * (ABM_SORT_PREBINNED), alternating between the two offset tables; the sort outputs are scratch.
 * No host synchronisation; overflow is reported through max_len as in abm_jkr_substep.              */
int abm_mechanics_loop(const double *radii, double *locations, const double *motility_forces, int64_t n,
                       const double extent_host[3], const double *origin_host, const double size_host[3],
                       double distance, double reach, int n_substeps, const double *dts_host,
                       int32_t *bin_start_a, int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos,
                       int32_t *sorted_bin, int32_t *max_count, void *sort_workspace, size_t sort_workspace_bytes,
                       const int32_t *old_row_start, int64_t old_row_stride, const int32_t *old_row_len,
                       const int32_t *old_cols, int32_t *len_a, int32_t *cols_a, int32_t *len_b, int32_t *cols_b,
                       int64_t row_stride, int32_t *max_len, const int32_t *global_id, double uniform_radius,
                       double poisson, double youngs, double adhesion_const, double stokes, int flags,
                       const int32_t *n_device, const double *move_window_device, void *stream);

/* abm_mechanics_loop with a stopwatch on its dominant kernel: substep_events_host (may be NULL) is a HOST array of
 * 2*n_substeps cudaEvent_t handles (as void*); events [2k] and [2k+1] are recorded on `stream` right before and right
 * after substep k's k_jkr_substep launch, so the caller reads every launch's duration inside the running loop
 * (cudaEventElapsedTime) -- what bench.py's roofline needs.  Not with ABM_LOOP_FP32.  Same results as
 * abm_mechanics_loop.                                                                               */
int abm_mechanics_loop_timed(const double *radii, double *locations, const double *motility_forces, int64_t n,
                             const double extent_host[3], const double *origin_host, const double size_host[3],
                             double distance, double reach, int n_substeps, const double *dts_host,
                             int32_t *bin_start_a, int32_t *bin_start_b, int32_t *sorted_id, double *sorted_pos,
                             int32_t *sorted_bin, int32_t *max_count, void *sort_workspace,
                             size_t sort_workspace_bytes, const int32_t *old_row_start, int64_t old_row_stride,
                             const int32_t *old_row_len, const int32_t *old_cols, int32_t *len_a, int32_t *cols_a,
                             int32_t *len_b, int32_t *cols_b, int64_t row_stride, int32_t *max_len,
                             const int32_t *global_id, double uniform_radius, double poisson, double youngs,
                             double adhesion_const, double stokes, int flags, const int32_t *n_device,
                             const double *move_window_device, void *const *substep_events_host, void *stream);

/* ---- morphogen diffusion ---------------------------------------------------------------------------
 * Replaces CellMethods.update_diffusion + update_diffusion_jit (cell_methods.py:441-468,
 * cell_backend.py:175-204) for one (gx,gy) float64 plane, row-major (the reference's field[:,:,0]).
 * clamp to [0,max_concentration] -> `steps` FTCS Jacobi substeps with reflecting walls, the last one
 * with last_dt -> multiply by (1-degradation).  Same operation order as the reference, no FMA:
 * bit-identical.  scratch: gx*gy doubles.  use_tma: 1 = TMA-staged tiles (needs gy even and
 * 16-byte aligned pointers, returns -1 otherwise), 0 = plain staged loads.  Substeps ping-pong between
 * field and scratch: if result_in_scratch_host is NULL the result is copied back into `field` when
 * `steps` is odd; otherwise no copy is made and *result_in_scratch_host says where the result is.
 * max_fused_substeps (1..5; 0 = 5): diffusion substeps advanced per kernel launch (temporal blocking in shared
 * memory, k_diffuse_fused; results are bit-identical for every value); chains of one use k_diffuse.   */
int abm_diffusion(double *field, double *scratch, int64_t gx, int64_t gy, int64_t steps, double diffuse_dt,
                  double last_dt, double diffuse_const, double spat_res2, double max_concentration,
                  double degradation, int use_tma, int max_fused_substeps, int *result_in_scratch_host,
                  void *stream);
/* number of kernel launches abm_diffusion enqueues for `steps` substeps with that max_fused_substeps (it has the
 * parity of `steps`, so the result lands where the substep-by-substep ping-pong would leave it) */
int64_t abm_diffusion_launches(int64_t steps, int max_fused_substeps);

/* 3-D extension (no reference counterpart: the reference's diffusion is 2-D only, cell_backend.py:184 "currently 2D";
 * oracle: oracle/abm_oracle.c orc_update_diffusion3d, bit-identical).  Same contract as abm_diffusion for a
 * (gx,gy,gz) float64 field with z contiguous (the reference's gradient[x][y][z]): clamp -> `steps` 7-point FTCS Jacobi
 * substeps (a = dt*D/h^2, b = 1-6a, reflecting walls, the last one with last_dt) -> multiply by (1-degradation).
 * One launch per substep; use_tma: 1 = planes staged by TMA (cp.async.bulk.tensor.3d; needs gz even and 16-byte
 * aligned fields), 0 = plain staged loads.                                                                    */
int abm_diffusion3d(double *field, double *scratch, int64_t gx, int64_t gy, int64_t gz, int64_t steps,
                    double diffuse_dt, double last_dt, double diffuse_const, double spat_res2,
                    double max_concentration, double degradation, int use_tma, int *result_in_scratch_host,
                    void *stream);

/* Replaces per-cell get_concentration (cell_methods.py:470-483) for all cells: nearest grid point
 * idx = ceil(floor(2x/h)/2) on the (gx,gy,gz) grid.  `field` holds the global rows [row_lo,row_hi) of that grid,
 * row-major (the whole grid: row_lo=0,row_hi=gx; a slab of it for a slab-decomposed colony).  sampled (N,) f64. */
int abm_sample_concentrations(const double *locations, int64_t n, const double *field,
                              const int64_t gsize_host[3], double spat_res, int64_t row_lo, int64_t row_hi,
                              double *sampled, const int32_t *n_device, void *stream);

/* Replaces per-cell adjust_morphogens (cell_methods.py:485-521) for all cells: amounts[i] is shared
 * equally by the in-bounds corner points closer than spat_res (fp64 atomic adds); amounts == NULL: every cell
 * deposits uniform_amount.  Shares that fall on rows
 * outside [row_lo,row_hi) belong to another slab and are skipped (the share size still counts them).  */
int abm_deposit_morphogens(const double *locations, const double *amounts, double uniform_amount, int64_t n,
                           double *field, const int64_t gsize_host[3], double spat_res, int64_t row_lo,
                           int64_t row_hi, const int32_t *n_device, void *stream);

/* 3-D extension of the deposit: amounts[i] is shared by the in-bounds corner points of the cell's grid CUBE (eight)
 * closer than spat_res (the reference deposits on the z = 0 plane only, cell_methods.py:499; oracle:
 * orc_adjust_morphogens3d).  Same arguments as abm_deposit_morphogens.                                          */
int abm_deposit_morphogens3d(const double *locations, const double *amounts, double uniform_amount, int64_t n,
                             double *field, const int64_t gsize_host[3], double spat_res, int64_t row_lo,
                             int64_t row_hi, const int32_t *n_device, void *stream);

/* ---- device-resident rule methods ----------------------------------------------------------------
 * cell_division, cell_death, cell_pathway, cell_differentiate, cell_motility (cell_methods.py:11-340) on a colony
 * that lives in fixed-capacity device arrays with its cell count in device memory.  Random draws follow
 * include/abm_rng.h (counter-based: seed, step, rule, persistent cell id, draw index); cell_pathway reads its
 * neighbours' FGF4 as of before the pass.  Both differ from the reference's serial Mersenne-Twister loop by design;
 * the CPU restatement these calls reproduce bit for bit is oracle/abm_oracle.c orc_rule_*.
 * Every array of abm_colony_arrays has `capacity` rows (bond_rows: capacity x bond_stride ids).  nbr_start / nbr_len
 * are the rows of the step's neighbour graph (abm_neighbor_search, searched BEFORE division and death as in
 * cell_simulation.py:90-97) and travel with their cells; its column entries keep the indices of search time and are
 * translated through index_map (written by abm_rule_death: new index, or -1 for a cell that died).            */
typedef struct abm_colony_arrays {
    double *locations, *radii, *motility_forces;
    int32_t *gid;                                  /* persistent cell id; array order is id order */
    int32_t *states, *GATA6, *NANOG, *FGF4, *FGFR, *ERK;
    int32_t *death_counters, *diff_counters, *div_counters, *fds_counters;
    int32_t *bond_len, *bond_rows;                 /* JKR bonds by id, fixed stride (abm_mechanics_loop, global_id) */
    int32_t *nbr_start, *nbr_len;
} abm_colony_arrays;

typedef struct abm_rule_params {                  /* cell_simulation.py:33-57, templates/experimental.yaml */
    int64_t seed, step;
    int32_t field, dox_step, lonely_thresh, guye_move;
    int32_t pluri_div_thresh, diff_div_thresh, pluri_to_diff, death_thresh, fds_thresh, three_d;
    double max_radius, min_radius, motility_force;
} abm_rule_params;

size_t abm_rules_workspace_bytes(int64_t capacity);

/* cell_division: stochastic division counters, thresholds, contact inhibition; every dividing cell's daughter is
 * appended behind the last cell (a copy of the mother, both displaced by +-random_vector*(max_radius-min_radius), new
 * id *next_gid_device + rank, no bonds, no neighbours).  *n_device and *next_gid_device are advanced on the device;
 * needed_device (int32[2], zeroed by the caller once) keeps the largest row count needed in [0] (daughters beyond
 * `capacity` are dropped: the caller checks it when it next synchronises).                                       */
int abm_rule_division(const abm_colony_arrays *cells, int64_t capacity, int64_t bond_stride, int32_t *n_device,
                      int32_t *next_gid_device, const abm_rule_params *params, int32_t *needed_device,
                      void *workspace, size_t workspace_bytes, void *stream);

/* cell_death: death counters of lonely pluripotent cells; the survivors are compacted (stable) into cells_out,
 * *n_device_out = their number, index_map (capacity) int32 = old index -> new index or -1.                        */
int abm_rule_death(const abm_colony_arrays *cells, const abm_colony_arrays *cells_out, int64_t capacity,
                   int64_t bond_stride, const int32_t *n_device, int32_t *n_device_out,
                   const abm_rule_params *params, int32_t *index_map, void *workspace, size_t workspace_bytes,
                   void *stream);

/* cell_pathway + cell_differentiate + cell_motility on the compacted colony.  nbr_cols: column buffer of the
 * neighbour search; fgf4_scratch: `capacity` int32.  motility_forces are ADDED to, as in the reference.            */
int abm_rule_update(const abm_colony_arrays *cells, int64_t capacity, const int32_t *n_device,
                    const int32_t *nbr_cols, const int32_t *index_map, const abm_rule_params *params,
                    int32_t *fgf4_scratch, void *stream);

/* ---- slab decomposition plumbing (no reference counterpart; see hipsc_abm_b200/slab.py) -------------------
 * A cell travels between ranks as one byte record:
 *   [0,24) position  [24,32) radius  [32,56) motility force  [56,60) global id  [60,64) bond_len
 *   [64,64+4*row_stride) bond row (global ids).
 * abm_slab_pack writes (in no particular order) the records of the owned cells with x < x_send_left (to the left
 * neighbour) and with x >= x_send_right (to the right one); counts (2) int32 receives the two record counts
 * (records beyond a capacity are dropped: the caller compares and retries with larger buffers).
 * n_own_device (may be NULL): device int32 -- only the first *n_own_device of the n_own rows are owned cells, the
 * rest (last step's halo cells) are skipped; lets a caller enqueue the pack before the owned count of the
 * finishing step has reached the host.
 * abm_slab_merge partitions the n_own owned cells followed by the received records (two segments, e.g. the
 * messages of the left and of the right neighbour; either may be empty) into
 * [cells owned after the exchange | halo cells] -- owned iff x in [x_lo,x_hi) (open towards a missing
 * neighbour) -- keeping input order inside each part, and writes them to the out_* arrays (capacity
 * n_own + n_received + n_received_b); n_owned_out (1) int32 receives the size of the first part.
 * workspace (merge only): abm_slab_workspace_bytes(n_own + n_received + n_received_b).  No host synchronisation.
 * Device-side counts (merge): n_own_device (may be NULL) caps the owned input rows at min(n_own, *n_own_device);
 * received_counts_device (may be NULL) = int32[2], the valid records of the two segments (clipped to n_received /
 * n_received_b, which are then the segments' capacities); n_local_out (may be NULL) receives the total number of
 * cells written (owned + halo).  With them a whole slab step is enqueued without a host round trip.
 * out_capacity (0 = n_own + n_received + n_received_b): rows of the out_* arrays; cells that do not fit are
 * dropped and *n_local_out (> out_capacity then) tells the caller how many rows it needed.                       */
size_t abm_slab_workspace_bytes(int64_t n);
int abm_slab_pack(const double *locations, const double *radii, const double *motility_forces,
                  const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows, int64_t n_own,
                  const int32_t *n_own_device, int64_t row_stride, double x_send_left, double x_send_right,
                  int has_left, int has_right, uint8_t *records_left, int64_t capacity_left, uint8_t *records_right, int64_t capacity_right,
                  int32_t *counts, void *workspace, size_t workspace_bytes, void *stream);
int abm_slab_merge(const double *locations, const double *radii, const double *motility_forces,
                   const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows, int64_t n_own,
                   const uint8_t *received, int64_t n_received, const uint8_t *received_b, int64_t n_received_b,
                   int64_t row_stride, double x_lo, double x_hi,
                   int has_left, int has_right, double *out_locations, double *out_radii, double *out_motility,
                   int32_t *out_global_id, int32_t *out_bond_len, int32_t *out_bond_rows, int32_t *n_owned_out,
                   void *workspace, size_t workspace_bytes, const int32_t *n_own_device,
                   const int32_t *received_counts_device, int32_t *n_local_out, int64_t out_capacity,
                   void *stream);

/* The same two steps for the pipeline that keeps every count on the device (hipsc_abm_b200/slab_device.py):
 *   - a message is `capacity` records followed by one header row whose first int32 is the number of valid records
 *     (buffers of (capacity+1) x (64+4*row_stride) bytes), so an exchange is one send and one receive per neighbour;
 *   - abm_slab_exchange_pack: one launch; counters (int32[3], zeroed by the caller ONCE, left zeroed by every call);
 *     worst (int32[2]) keeps the largest record counts ever claimed (> capacity: records were dropped);
 *   - abm_slab_exchange_merge: two launches; the owned rows (min(rows, *n_own_device)) and the valid records of the two
 *     messages are partitioned into [owned after the exchange | halo] of the out_* arrays (out_rows rows) in NO
 *     particular order inside each part, and their positions are also written to `snapshot` (out_rows x 3) -- the
 *     positions the step starts with.  tally (int32[5], zeroed by the caller once, left zeroed by every call);
 *     *n_owned_out / *n_local_out receive the two counts (clipped to out_rows), worst_rows (int32[1]) the largest
 *     total ever needed.                                                                                          */
int abm_slab_exchange_pack(const double *locations, const double *radii, const double *motility_forces,
                           const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows, int64_t rows,
                           const int32_t *n_own_device, int64_t row_stride, double x_send_left, double x_send_right,
                           int has_left, int has_right, uint8_t *message_left, uint8_t *message_right,
                           int64_t capacity, int32_t *counters, int32_t *worst, void *stream);
int abm_slab_exchange_merge(const double *locations, const double *radii, const double *motility_forces,
                            const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows, int64_t rows,
                            const int32_t *n_own_device, const uint8_t *message_left, const uint8_t *message_right,
                            int64_t capacity, int64_t row_stride, double x_lo, double x_hi, int has_left,
                            int has_right, double *out_locations, double *out_radii, double *out_motility,
                            int32_t *out_global_id, int32_t *out_bond_len, int32_t *out_bond_rows, double *snapshot,
                            int64_t out_rows, int32_t *n_owned_out, int32_t *n_local_out, int32_t *tally,
                            int32_t *worst_rows, void *stream);

/* Cell exchange over NVLink peer memory (slab_device.PeerSlab; one node, one process per GPU; the neighbours' receive
 * buffers and flag words are mapped through CUDA IPC).  Protocol per step k: a rank stores its two messages into the
 * neighbours' receive buffers and releases k into their "data" flag words (abm_peer_exchange, ONE launch; a side first
 * polls the "ack" word the neighbour wrote for k-1, so a buffer is never overwritten before it was merged); its stream
 * then waits for k in its own two data words (abm_peer_flags with wait_*), merges (abm_slab_exchange_merge), and
 * releases k into the neighbours' ack words (abm_peer_flags with release_*).  No host synchronisation.
 *   abm_ipc_open: cudaIpcOpenMemHandle of a neighbour's allocation (handle_host: the 64 bytes of its
 *     cudaIpcMemHandle_t) with THIS process's device current; *base_host receives the mapped base address.
 *   abm_peer_exchange: only the valid records (their count is in the header row behind `capacity` records of
 *     `record_bytes`) and the header row travel.  A side whose peer_message_* is NULL does not exist.
 *     block_counters: int32[2] on this device, zeroed once by the caller.
 *   Every poll reads a word of the polling device's own memory, gives up after max_wait_ns and then raises
 *   *timed_out (int32, device) to the step it waited for instead of hanging the GPU.  A flag value is
 *   value + *value_device when value_device is given: a captured CUDA graph, whose arguments are frozen, reads the
 *   step number from a counter it increments itself.                                                          */
int abm_ipc_open(const uint8_t *handle_host, int64_t handle_bytes, void **base_host);
int abm_peer_exchange(uint8_t *peer_message_left, const uint8_t *message_left, int32_t *peer_data_flag_left,
                      const int32_t *ack_flag_left, uint8_t *peer_message_right, const uint8_t *message_right,
                      int32_t *peer_data_flag_right, const int32_t *ack_flag_right, int64_t capacity,
                      int64_t record_bytes, int32_t value, const int32_t *value_device, int32_t *block_counters,
                      int64_t max_wait_ns, int32_t *timed_out, void *stream);
int abm_peer_flags(const int32_t *wait_a, const int32_t *wait_b, int32_t *release_a, int32_t *release_b,
                   int32_t value, const int32_t *value_device, int64_t max_wait_ns, int32_t *timed_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ABM_B200_H */
