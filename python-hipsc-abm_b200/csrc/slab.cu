// slab.cu -- device-side plumbing of the slab decomposition (hipsc_abm_b200/slab.py): selecting and packing the
// cells a rank sends to its x-neighbours, and merging what it receives with the cells it keeps.  The reference has
// no multi-device path; these kernels exist because doing the same with a few dozen framework tensor ops per
// step is host-launch-bound (about 1.2 ms per rank and step at 500k cells).
//
// Record layout (bytes), one per cell, shared with slab.py:
//   [0,24) position  [24,32) radius  [32,56) motility force  [56,60) global id  [60,64) bond_len  [64,64+4*stride) bond row
#include "scan.cuh"

namespace abm {

struct CellArrays {
    double *loc, *radii, *mot;
    int *gid, *bond_len, *bond_rows;
};

__device__ __forceinline__ bool owns(double x, double x_lo, double x_hi, int has_left, int has_right) {
    return (!has_left || x >= x_lo) && (!has_right || x < x_hi);
}

// flags[i] = 1 if owned cell i goes to the left (side 0) / right (side 1) neighbour
// n_own_dev (may be nullptr): only the first *n_own_dev of the n rows are owned cells (the rest are last step's
// halo cells, which are not sent) -- lets the caller enqueue the pack before it knows the owned count itself
__global__ void __launch_bounds__(256) k_slab_flags(const double *__restrict__ loc, int n,
                                                    const int *__restrict__ n_own_dev, double x_send_left,
                                                    double x_send_right, int has_left, int has_right,
                                                    int *__restrict__ flag_l, int *__restrict__ flag_r) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i > n) return;
    int l = 0, r = 0;
    if (i < n && (!n_own_dev || i < *n_own_dev)) {
        double x = loc[3 * (long long)i];
        l = has_left && x < x_send_left;
        r = has_right && x >= x_send_right;
    }
    flag_l[i] = l;   // element n stays 0 so that the scan leaves the total there
    flag_r[i] = r;
}

__device__ __forceinline__ void write_record(unsigned char *rec, const CellArrays &a, int i, int stride) {
    double *d = reinterpret_cast<double *>(rec);
    d[0] = a.loc[3 * (long long)i + 0];
    d[1] = a.loc[3 * (long long)i + 1];
    d[2] = a.loc[3 * (long long)i + 2];
    d[3] = a.radii[i];
    d[4] = a.mot[3 * (long long)i + 0];
    d[5] = a.mot[3 * (long long)i + 1];
    d[6] = a.mot[3 * (long long)i + 2];
    int *w = reinterpret_cast<int *>(rec + 56);
    w[0] = a.gid[i];
    w[1] = a.bond_len[i];
    const int *row = a.bond_rows + (long long)i * stride;
    for (int k = 0; k < stride; ++k) w[2 + k] = row[k];
}

__device__ __forceinline__ void read_record(const unsigned char *rec, const CellArrays &a, int dst, int stride) {
    const double *d = reinterpret_cast<const double *>(rec);
    a.loc[3 * (long long)dst + 0] = d[0];
    a.loc[3 * (long long)dst + 1] = d[1];
    a.loc[3 * (long long)dst + 2] = d[2];
    a.radii[dst] = d[3];
    a.mot[3 * (long long)dst + 0] = d[4];
    a.mot[3 * (long long)dst + 1] = d[5];
    a.mot[3 * (long long)dst + 2] = d[6];
    const int *w = reinterpret_cast<const int *>(rec + 56);
    a.gid[dst] = w[0];
    a.bond_len[dst] = w[1];
    int *row = a.bond_rows + (long long)dst * stride;
    for (int k = 0; k < stride; ++k) row[k] = w[2 + k];
}

// pos_* hold the exclusive scans of the flags: a flagged cell writes its record at that position
__global__ void __launch_bounds__(256) k_slab_pack(CellArrays a, int n, int stride, int rec_bytes,
                                                   const int *__restrict__ pos_l, const int *__restrict__ pos_r,
                                                   unsigned char *__restrict__ out_l, unsigned char *__restrict__ out_r,
                                                   long long cap_l, long long cap_r) {
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    if (pos_l[i + 1] != pos_l[i] && pos_l[i] < cap_l) write_record(out_l + (long long)pos_l[i] * rec_bytes, a, i, stride);
    if (pos_r[i + 1] != pos_r[i] && pos_r[i] < cap_r) write_record(out_r + (long long)pos_r[i] * rec_bytes, a, i, stride);
}

// keep[i] for the n_own owned cells followed by the k received records: 1 = owned after the exchange
__global__ void __launch_bounds__(256) k_slab_keep(const double *__restrict__ loc, int n_own,
                                                   const unsigned char *__restrict__ recv, int k, int rec_bytes,
                                                   double x_lo, double x_hi, int has_left, int has_right,
                                                   int *__restrict__ keep) {
    int i = blockIdx.x * 256 + threadIdx.x;
    int total = n_own + k;
    if (i > total) return;
    int f = 0;
    if (i < n_own) f = owns(loc[3 * (long long)i], x_lo, x_hi, has_left, has_right);
    else if (i < total)
        f = owns(*reinterpret_cast<const double *>(recv + (long long)(i - n_own) * rec_bytes), x_lo, x_hi, has_left,
                 has_right);
    keep[i] = f;
}

// stable two-way partition: owned-next cells first (in input order), then everything else (halo cells)
__global__ void __launch_bounds__(256) k_slab_merge(CellArrays in, int n_own, const unsigned char *__restrict__ recv,
                                                    int k, int rec_bytes, int stride, const int *__restrict__ pos,
                                                    CellArrays out) {
    int i = blockIdx.x * 256 + threadIdx.x;
    int total = n_own + k;
    if (i >= total) return;
    int n_keep = pos[total];
    bool kept = pos[i + 1] != pos[i];
    int dst = kept ? pos[i] : n_keep + (i - pos[i]);
    if (i < n_own) {
        out.loc[3 * (long long)dst + 0] = in.loc[3 * (long long)i + 0];
        out.loc[3 * (long long)dst + 1] = in.loc[3 * (long long)i + 1];
        out.loc[3 * (long long)dst + 2] = in.loc[3 * (long long)i + 2];
        out.radii[dst] = in.radii[i];
        out.mot[3 * (long long)dst + 0] = in.mot[3 * (long long)i + 0];
        out.mot[3 * (long long)dst + 1] = in.mot[3 * (long long)i + 1];
        out.mot[3 * (long long)dst + 2] = in.mot[3 * (long long)i + 2];
        out.gid[dst] = in.gid[i];
        out.bond_len[dst] = in.bond_len[i];
        const int *src = in.bond_rows + (long long)i * stride;
        int *row = out.bond_rows + (long long)dst * stride;
        for (int c = 0; c < stride; ++c) row[c] = src[c];
    } else {
        read_record(recv + (long long)(i - n_own) * rec_bytes, out, dst, stride);
    }
}

static size_t scan_ws_ints(int64_t n) { return ((size_t)n + 63) / 64 * 64 + (size_t)scan_blocks(n) + 64; }

}  // namespace abm

using namespace abm;

ABM_EXPORT size_t abm_slab_workspace_bytes(int64_t n) {
    if (n < 0) return 0;
    return 2 * scan_ws_ints(n + 1) * 4 + 256;
}

ABM_EXPORT int abm_slab_pack(const double *locations, const double *radii, const double *motility_forces,
                             const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows,
                             int64_t n_own, const int32_t *n_own_device, int64_t row_stride, double x_send_left,
                             double x_send_right, int has_left,
                             int has_right, uint8_t *records_left, int64_t capacity_left, uint8_t *records_right,
                             int64_t capacity_right, int32_t *counts, void *workspace, size_t workspace_bytes,
                             void *stream) {
    ABM_REQUIRE(n_own >= 0 && n_own < 0x7fffffff && row_stride > 0, "bad sizes");
    ABM_REQUIRE(counts && workspace && workspace_bytes >= abm_slab_workspace_bytes(n_own), "workspace too small");
    cudaStream_t st = as_stream(stream);
    int rc = check_cuda(cudaMemsetAsync(counts, 0, 8, st), "memset counts");
    if (rc || n_own == 0) return rc;
    ABM_REQUIRE(locations && radii && motility_forces && global_id && bond_len && bond_rows, "null pointer");
    size_t half = scan_ws_ints(n_own + 1);
    int *flag_l = (int *)workspace, *flag_r = flag_l + half;
    int *sums_l = flag_l + ((size_t)(n_own + 1) + 63) / 64 * 64, *sums_r = flag_r + ((size_t)(n_own + 1) + 63) / 64 * 64;
    k_slab_flags<<<ceil_div(n_own + 1, 256), 256, 0, st>>>(locations, (int)n_own, n_own_device, x_send_left,
                                                           x_send_right, has_left, has_right, flag_l, flag_r);
    ABM_CHECK_LAUNCH("k_slab_flags");
    rc = exclusive_scan_inplace(flag_l, n_own + 1, sums_l, nullptr, nullptr, st);
    if (rc) return rc;
    rc = exclusive_scan_inplace(flag_r, n_own + 1, sums_r, nullptr, nullptr, st);
    if (rc) return rc;
    rc = check_cuda(cudaMemcpyAsync(counts, flag_l + n_own, 4, cudaMemcpyDeviceToDevice, st), "count left");
    if (rc) return rc;
    rc = check_cuda(cudaMemcpyAsync(counts + 1, flag_r + n_own, 4, cudaMemcpyDeviceToDevice, st), "count right");
    if (rc) return rc;
    CellArrays a{const_cast<double *>(locations), const_cast<double *>(radii), const_cast<double *>(motility_forces),
                 const_cast<int *>(global_id), const_cast<int *>(bond_len), const_cast<int *>(bond_rows)};
    int rec_bytes = 64 + 4 * (int)row_stride;
    k_slab_pack<<<ceil_div(n_own, 256), 256, 0, st>>>(a, (int)n_own, (int)row_stride, rec_bytes, flag_l, flag_r,
                                                      records_left, records_right, (long long)capacity_left,
                                                      (long long)capacity_right);
    ABM_CHECK_LAUNCH("k_slab_pack");
    return 0;
}

ABM_EXPORT int abm_slab_merge(const double *locations, const double *radii, const double *motility_forces,
                              const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows,
                              int64_t n_own, const uint8_t *received, int64_t n_received, int64_t row_stride,
                              double x_lo, double x_hi, int has_left, int has_right, double *out_locations,
                              double *out_radii, double *out_motility, int32_t *out_global_id, int32_t *out_bond_len,
                              int32_t *out_bond_rows, int32_t *n_owned_out, void *workspace, size_t workspace_bytes,
                              void *stream) {
    int64_t total = n_own + n_received;
    ABM_REQUIRE(n_own >= 0 && n_received >= 0 && total < 0x7fffffff && row_stride > 0, "bad sizes");
    ABM_REQUIRE(n_owned_out && workspace && workspace_bytes >= abm_slab_workspace_bytes(total), "workspace too small");
    cudaStream_t st = as_stream(stream);
    int rc = check_cuda(cudaMemsetAsync(n_owned_out, 0, 4, st), "memset n_owned_out");
    if (rc || total == 0) return rc;
    ABM_REQUIRE(out_locations && out_radii && out_motility && out_global_id && out_bond_len && out_bond_rows,
                "null output");
    ABM_REQUIRE(out_locations != locations && out_bond_rows != bond_rows, "merge is out of place");
    int *keep = (int *)workspace;
    int *sums = keep + ((size_t)(total + 1) + 63) / 64 * 64;
    int rec_bytes = 64 + 4 * (int)row_stride;
    k_slab_keep<<<ceil_div(total + 1, 256), 256, 0, st>>>(locations, (int)n_own, received, (int)n_received, rec_bytes,
                                                          x_lo, x_hi, has_left, has_right, keep);
    ABM_CHECK_LAUNCH("k_slab_keep");
    rc = exclusive_scan_inplace(keep, total + 1, sums, nullptr, nullptr, st);
    if (rc) return rc;
    rc = check_cuda(cudaMemcpyAsync(n_owned_out, keep + total, 4, cudaMemcpyDeviceToDevice, st), "owned count");
    if (rc) return rc;
    CellArrays in{const_cast<double *>(locations), const_cast<double *>(radii), const_cast<double *>(motility_forces),
                  const_cast<int *>(global_id), const_cast<int *>(bond_len), const_cast<int *>(bond_rows)};
    CellArrays out{out_locations, out_radii, out_motility, out_global_id, out_bond_len, out_bond_rows};
    k_slab_merge<<<ceil_div(total, 256), 256, 0, st>>>(in, (int)n_own, received, (int)n_received, rec_bytes,
                                                       (int)row_stride, keep, out);
    ABM_CHECK_LAUNCH("k_slab_merge");
    return 0;
}
