// slab.cu -- device-side plumbing of the slab decomposition (hipsc_abm_b200/slab.py): selecting and packing the
// cells a rank sends to its x-neighbours, and merging what it receives with the cells it keeps.  The reference has
// no multi-device path; these kernels exist because doing the same with a few dozen framework tensor ops per
// step is host-launch-bound (about 1.2 ms per rank and step at 500k cells).
//
// Record layout (bytes), one per cell, shared with slab.py:
//   [0,24) position  [24,32) radius  [32,56) motility force  [56,60) global id  [60,64) bond_len  [64,64+4*stride) bond row
#include <string.h>

#include "scan.cuh"

namespace abm {

struct CellArrays {
    double *loc, *radii, *mot;
    int *gid, *bond_len, *bond_rows;
};

__device__ __forceinline__ bool owns(double x, double x_lo, double x_hi, int has_left, int has_right) {
    return (!has_left || x >= x_lo) && (!has_right || x < x_hi);
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
    // only the live part of the row travels (entries beyond bond_len are never read); 16 bytes at a time when the
    // record and the row are aligned (record = 64 + 4*stride bytes: a multiple of 16 when stride % 4 == 0)
    const int len = min(w[1], stride);
    if ((stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(rec) | reinterpret_cast<uintptr_t>(row)) & 15) == 0) {
        for (int k = 0; k < len; k += 4)
            *reinterpret_cast<int4 *>(w + 2 + k) = *reinterpret_cast<const int4 *>(row + k);
    } else {
        for (int k = 0; k < len; ++k) w[2 + k] = row[k];
    }
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
    const int len = min(w[1], stride);
    if ((stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(rec) | reinterpret_cast<uintptr_t>(row)) & 15) == 0) {
        for (int k = 0; k < len; k += 4)
            *reinterpret_cast<int4 *>(row + k) = *reinterpret_cast<const int4 *>(w + 2 + k);
    } else {
        for (int k = 0; k < len; ++k) row[k] = w[2 + k];
    }
}

// One pass: a cell to be sent claims its record slot from the message's counter (one atomic per warp and side).
// The order of the records inside a message is therefore unspecified -- nothing downstream depends on the local
// order of the cells (rows are kept by global id, force sums are integer sums).
// n_own_dev (may be nullptr): only the first *n_own_dev of the n rows are owned cells (the rest are last step's
// halo cells, which are not sent) -- lets the caller enqueue the pack before it knows the owned count itself.
__global__ void __launch_bounds__(256) k_slab_pack(CellArrays a, int n, const int *__restrict__ n_own_dev, int stride,
                                                   int rec_bytes, double x_send_left, double x_send_right,
                                                   int has_left, int has_right, unsigned char *__restrict__ out_l,
                                                   unsigned char *__restrict__ out_r, long long cap_l, long long cap_r,
                                                   int *__restrict__ counts) {
    int i = blockIdx.x * 256 + threadIdx.x;
    bool l = false, r = false;
    if (i < n && (!n_own_dev || i < *n_own_dev)) {
        double x = a.loc[3 * (long long)i];
        l = has_left && x < x_send_left;
        r = has_right && x >= x_send_right;
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const bool f = side ? r : l;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (m == 0) continue;
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&counts[side], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (f) {
            long long pos = base + __popc(m & ((1u << lane) - 1));
            if (pos < (side ? cap_r : cap_l)) write_record((side ? out_r : out_l) + pos * rec_bytes, a, i, stride);
        }
    }
}

// The received records come as two segments (from the left and from the right neighbour): record j of the
// concatenation lives in segment a for j < k_a, else in segment b.
struct Received {
    const unsigned char *a, *b;
    int k_a, k_b;
    // the valid record counts may live on the device (counts[0], counts[1], clipped to the capacities k_a / k_b) and
    // so may the number of owned input rows: resolve() is the first thing a kernel does
    const int *counts, *n_own_dev;
    __device__ __forceinline__ void resolve(int &n_own) {
        n_own = device_count(n_own, n_own_dev);
        if (counts) {
            k_a = min(k_a, max(counts[0], 0));
            k_b = min(k_b, max(counts[1], 0));
        }
    }
    __device__ __forceinline__ const unsigned char *record(int j, int rec_bytes) const {
        return j < k_a ? a + (long long)j * rec_bytes : b + (long long)(j - k_a) * rec_bytes;
    }
};

// keep[i] for the n_own owned cells followed by the k received records: 1 = owned after the exchange
__global__ void __launch_bounds__(256) k_slab_keep(const double *__restrict__ loc, int n_own, Received recv,
                                                   int rec_bytes, double x_lo, double x_hi, int has_left,
                                                   int has_right, int *__restrict__ keep, int capacity,
                                                   int *__restrict__ n_local_out) {
    int i = blockIdx.x * 256 + threadIdx.x;
    recv.resolve(n_own);
    int total = n_own + recv.k_a + recv.k_b;
    if (i == 0 && n_local_out) *n_local_out = total;
    if (i > capacity) return;                      // flags up to the capacity: the scan runs over all of them
    int f = 0;
    if (i < n_own) f = owns(loc[3 * (long long)i], x_lo, x_hi, has_left, has_right);
    else if (i < total)
        f = owns(*reinterpret_cast<const double *>(recv.record(i - n_own, rec_bytes)), x_lo, x_hi, has_left,
                 has_right);
    keep[i] = f;
}

// stable two-way partition: owned-next cells first (in input order), then everything else (halo cells)
__global__ void __launch_bounds__(256) k_slab_merge(CellArrays in, int n_own, Received recv, int rec_bytes,
                                                    int stride, const int *__restrict__ pos, CellArrays out,
                                                    int out_capacity) {
    int i = blockIdx.x * 256 + threadIdx.x;
    recv.resolve(n_own);
    int total = n_own + recv.k_a + recv.k_b;
    if (i >= total) return;
    int n_keep = pos[total];
    bool kept = pos[i + 1] != pos[i];
    int dst = kept ? pos[i] : n_keep + (i - pos[i]);
    if (dst >= out_capacity) return;               // does not fit: dropped; n_local_out tells the caller how many rows it needed
    if (i < n_own) {
        out.loc[3 * (long long)dst + 0] = in.loc[3 * (long long)i + 0];
        out.loc[3 * (long long)dst + 1] = in.loc[3 * (long long)i + 1];
        out.loc[3 * (long long)dst + 2] = in.loc[3 * (long long)i + 2];
        out.radii[dst] = in.radii[i];
        out.mot[3 * (long long)dst + 0] = in.mot[3 * (long long)i + 0];
        out.mot[3 * (long long)dst + 1] = in.mot[3 * (long long)i + 1];
        out.mot[3 * (long long)dst + 2] = in.mot[3 * (long long)i + 2];
        out.gid[dst] = in.gid[i];
        const int len = in.bond_len[i];
        out.bond_len[dst] = len;
        const int *src = in.bond_rows + (long long)i * stride;
        int *row = out.bond_rows + (long long)dst * stride;
        // only the live part of the row travels (entries beyond bond_len are never read), 16 bytes at a time when
        // the rows are aligned
        if ((stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(in.bond_rows) | reinterpret_cast<uintptr_t>(out.bond_rows)) & 15) == 0) {
            for (int c = 0; c < len; c += 4)
                *reinterpret_cast<int4 *>(row + c) = *reinterpret_cast<const int4 *>(src + c);
        } else {
            for (int c = 0; c < len; ++c) row[c] = src[c];
        }
    } else {
        read_record(recv.record(i - n_own, rec_bytes), out, dst, stride);
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
    ABM_REQUIRE(counts, "null counts");
    (void)workspace;
    (void)workspace_bytes;
    cudaStream_t st = as_stream(stream);
    int rc = check_cuda(cudaMemsetAsync(counts, 0, 8, st), "memset counts");
    if (rc || n_own == 0) return rc;
    ABM_REQUIRE(locations && radii && motility_forces && global_id && bond_len && bond_rows, "null pointer");
    ABM_REQUIRE((!has_left || records_left || capacity_left == 0) && (!has_right || records_right || capacity_right == 0),
                "null record buffer");
    CellArrays a{const_cast<double *>(locations), const_cast<double *>(radii), const_cast<double *>(motility_forces),
                 const_cast<int *>(global_id), const_cast<int *>(bond_len), const_cast<int *>(bond_rows)};
    int rec_bytes = 64 + 4 * (int)row_stride;
    k_slab_pack<<<ceil_div(n_own, 256), 256, 0, st>>>(a, (int)n_own, n_own_device, (int)row_stride, rec_bytes,
                                                      x_send_left, x_send_right, has_left, has_right, records_left,
                                                      records_right, (long long)capacity_left,
                                                      (long long)capacity_right, counts);
    ABM_CHECK_LAUNCH("k_slab_pack");
    return 0;
}

ABM_EXPORT int abm_slab_merge(const double *locations, const double *radii, const double *motility_forces,
                              const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows,
                              int64_t n_own, const uint8_t *received, int64_t n_received, const uint8_t *received_b,
                              int64_t n_received_b, int64_t row_stride,
                              double x_lo, double x_hi, int has_left, int has_right, double *out_locations,
                              double *out_radii, double *out_motility, int32_t *out_global_id, int32_t *out_bond_len,
                              int32_t *out_bond_rows, int32_t *n_owned_out, void *workspace, size_t workspace_bytes,
                              const int32_t *n_own_device, const int32_t *received_counts_device,
                              int32_t *n_local_out, int64_t out_capacity, void *stream) {
    int64_t total = n_own + n_received + n_received_b;          // the capacity when the counts live on the device
    ABM_REQUIRE(n_own >= 0 && n_received >= 0 && n_received_b >= 0 && total < 0x7fffffff && row_stride > 0, "bad sizes");
    ABM_REQUIRE((received || n_received == 0) && (received_b || n_received_b == 0), "null record segment");
    Received recv{received, received_b, (int)n_received, (int)n_received_b, received_counts_device, n_own_device};
    ABM_REQUIRE(n_owned_out && workspace && workspace_bytes >= abm_slab_workspace_bytes(total), "workspace too small");
    cudaStream_t st = as_stream(stream);
    int rc = check_cuda(cudaMemsetAsync(n_owned_out, 0, 4, st), "memset n_owned_out");
    if (!rc && n_local_out) rc = check_cuda(cudaMemsetAsync(n_local_out, 0, 4, st), "memset n_local_out");
    if (rc || total == 0) return rc;
    ABM_REQUIRE(out_locations && out_radii && out_motility && out_global_id && out_bond_len && out_bond_rows,
                "null output");
    ABM_REQUIRE(out_locations != locations && out_bond_rows != bond_rows, "merge is out of place");
    int *keep = (int *)workspace;
    int *sums = keep + ((size_t)(total + 1) + 63) / 64 * 64;
    int rec_bytes = 64 + 4 * (int)row_stride;
    k_slab_keep<<<ceil_div(total + 1, 256), 256, 0, st>>>(locations, (int)n_own, recv, rec_bytes, x_lo, x_hi, has_left,
                                                          has_right, keep, (int)total, n_local_out);
    ABM_CHECK_LAUNCH("k_slab_keep");
    rc = exclusive_scan_inplace(keep, total + 1, sums, nullptr, nullptr, st);
    if (rc) return rc;
    rc = check_cuda(cudaMemcpyAsync(n_owned_out, keep + total, 4, cudaMemcpyDeviceToDevice, st), "owned count");
    if (rc) return rc;
    CellArrays in{const_cast<double *>(locations), const_cast<double *>(radii), const_cast<double *>(motility_forces),
                  const_cast<int *>(global_id), const_cast<int *>(bond_len), const_cast<int *>(bond_rows)};
    CellArrays out{out_locations, out_radii, out_motility, out_global_id, out_bond_len, out_bond_rows};
    if (out_capacity <= 0 || out_capacity > total) out_capacity = total;
    k_slab_merge<<<ceil_div(total, 256), 256, 0, st>>>(in, (int)n_own, recv, rec_bytes, (int)row_stride, keep, out,
                                                       (int)out_capacity);
    ABM_CHECK_LAUNCH("k_slab_merge");
    return 0;
}

// ---- the device-count pipeline (hipsc_abm_b200/slab_device.py): leaner variants ---------------------------------
// A message is `capacity` records followed by ONE header row whose first int32 is the number of valid records, so an
// exchange is one send and one receive per neighbour.  Nothing here depends on the order of the cells inside the
// owned / halo parts (rows are kept by global id, force sums are integer sums), which allows placement by atomic
// cursors instead of a stable scan: two launches per merge instead of five.
namespace abm {

// pack + header + overflow record in one launch (the header is written by the last block to finish)
__global__ void __launch_bounds__(256)
k_slab_pack_hdr(CellArrays a, int rows, const int *__restrict__ n_own_dev, int stride, int rec_bytes,
                double x_send_left, double x_send_right, int has_left, int has_right,
                unsigned char *__restrict__ out_l, unsigned char *__restrict__ out_r, long long cap,
                int *__restrict__ counts /*[0,1]: rows claimed, [2]: block counter*/, int *__restrict__ worst) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = device_count(rows, n_own_dev);
    bool l = false, r = false;
    if (i < n) {
        const double x = a.loc[3 * (long long)i];
        l = has_left && x < x_send_left;
        r = has_right && x >= x_send_right;
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int side = 0; side < 2; ++side) {
        const bool f = side ? r : l;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        if (m == 0) continue;
        int base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(&counts[side], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (f) {
            const long long pos = base + __popc(m & ((1u << lane) - 1));
            if (pos < cap) write_record((side ? out_r : out_l) + pos * rec_bytes, a, i, stride);
        }
    }
    __shared__ bool last;
    __threadfence();
    __syncthreads();          // every warp of this block is done before thread 0 reports the block as finished
    if (threadIdx.x == 0) last = atomicAdd(&counts[2], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        const int cl = counts[0], cr = counts[1];
        *reinterpret_cast<int *>(out_l + cap * rec_bytes) = cl;
        *reinterpret_cast<int *>(out_r + cap * rec_bytes) = cr;
        atomicMax(&worst[0], cl);
        atomicMax(&worst[1], cr);
        counts[0] = counts[1] = counts[2] = 0;                 // ready for the next step
    }
}

struct Incoming {
    const unsigned char *l, *r;      // messages (capacity records + header row), or nullptr
    int cap;
    __device__ __forceinline__ void counts(int rec_bytes, int &kl, int &kr) const {
        kl = l ? min(max(*reinterpret_cast<const int *>(l + (long long)cap * rec_bytes), 0), cap) : 0;
        kr = r ? min(max(*reinterpret_cast<const int *>(r + (long long)cap * rec_bytes), 0), cap) : 0;
    }
    __device__ __forceinline__ const unsigned char *record(int j, int kl, int rec_bytes) const {
        return j < kl ? l + (long long)j * rec_bytes : r + (long long)(j - kl) * rec_bytes;
    }
};

// pass 1: how many of (owned rows + received records) are owned after the exchange
__global__ void __launch_bounds__(256)
k_slab_count(const double *__restrict__ loc, int rows, const int *__restrict__ n_own_dev, Incoming in, int rec_bytes,
             double x_lo, double x_hi, int has_left, int has_right, int *__restrict__ tally /*[0] owned, [1] total*/) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = device_count(rows, n_own_dev);
    int kl, kr;
    in.counts(rec_bytes, kl, kr);
    const int total = n + kl + kr;
    bool f = false;
    if (i < n) f = owns(loc[3 * (long long)i], x_lo, x_hi, has_left, has_right);
    else if (i < total)
        f = owns(*reinterpret_cast<const double *>(in.record(i - n, kl, rec_bytes)), x_lo, x_hi, has_left, has_right);
    // one atomic per BLOCK (16 k warps adding to one word serialise in L2)
    __shared__ int warp_owned[8];
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if ((threadIdx.x & 31) == 0) warp_owned[threadIdx.x >> 5] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int sum = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += warp_owned[w];
        if (sum) atomicAdd(&tally[0], sum);
    }
    if (i == 0) tally[1] = total;
}

// pass 2: owned cells claim slots from the front, halo cells behind them; also the position snapshot of the step
__global__ void __launch_bounds__(256)
k_slab_place(CellArrays a, int rows, const int *__restrict__ n_own_dev, Incoming in, int rec_bytes, int stride,
             double x_lo, double x_hi, int has_left, int has_right, CellArrays out, double *__restrict__ snap,
             int out_rows, int *__restrict__ tally /*[0] owned, [1] total, [2],[3] cursors, [4] block counter*/,
             int *__restrict__ n_owned_out, int *__restrict__ n_local_out, int *__restrict__ worst_rows) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n = device_count(rows, n_own_dev);
    int kl, kr;
    in.counts(rec_bytes, kl, kr);
    const int total = n + kl + kr, n_keep = tally[0];
    const bool valid = i < total;
    bool f = false;
    if (i < n) f = owns(a.loc[3 * (long long)i], x_lo, x_hi, has_left, has_right);
    else if (valid)
        f = owns(*reinterpret_cast<const double *>(in.record(i - n, kl, rec_bytes)), x_lo, x_hi, has_left, has_right);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // slots are claimed per BLOCK: one atomic per part and block (the order inside a part is free), ranks inside the
    // block from the warps' ballots
    __shared__ int part_count[2][8];     // cells of part p in warp w, then the exclusive prefix over the warps
    __shared__ int part_base[2];
    unsigned ballots[2];
#pragma unroll
    for (int part = 0; part < 2; ++part) {
        ballots[part] = __ballot_sync(0xffffffffu, valid && (part == 0 ? f : !f));
        if (lane == 0) part_count[part][warp] = __popc(ballots[part]);
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int part = threadIdx.x;
        int sum = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int c = part_count[part][w];
            part_count[part][w] = sum;
            sum += c;
        }
        part_base[part] = sum ? atomicAdd(&tally[2 + part], sum) : 0;
    }
    __syncthreads();
    int dst = -1;
    if (valid) {
        const int part = f ? 0 : 1;
        dst = (part ? n_keep : 0) + part_base[part] + part_count[part][warp] +
              __popc(ballots[part] & ((1u << lane) - 1));
    }
    if (valid && dst < out_rows) {
        if (i < n) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double x = a.loc[3 * (long long)i + k];
                out.loc[3 * (long long)dst + k] = x;
                snap[3 * (long long)dst + k] = x;
                out.mot[3 * (long long)dst + k] = a.mot[3 * (long long)i + k];
            }
            out.radii[dst] = a.radii[i];
            out.gid[dst] = a.gid[i];
            const int len = min(a.bond_len[i], stride);
            out.bond_len[dst] = len;
            const int *src = a.bond_rows + (long long)i * stride;
            int *row = out.bond_rows + (long long)dst * stride;
            if ((stride & 3) == 0 &&
                ((reinterpret_cast<uintptr_t>(a.bond_rows) | reinterpret_cast<uintptr_t>(out.bond_rows)) & 15) == 0) {
                for (int c = 0; c < len; c += 4)
                    *reinterpret_cast<int4 *>(row + c) = *reinterpret_cast<const int4 *>(src + c);
            } else {
                for (int c = 0; c < len; ++c) row[c] = src[c];
            }
        } else {
            read_record(in.record(i - n, kl, rec_bytes), out, dst, stride);
#pragma unroll
            for (int k = 0; k < 3; ++k) snap[3 * (long long)dst + k] = out.loc[3 * (long long)dst + k];
        }
    }
    __shared__ bool last;
    __threadfence();
    __syncthreads();          // every warp of this block is done before thread 0 reports the block as finished
    if (threadIdx.x == 0) last = atomicAdd(&tally[4], 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *n_owned_out = min(n_keep, out_rows);
        *n_local_out = min(total, out_rows);
        atomicMax(worst_rows, total);
        tally[0] = tally[1] = tally[2] = tally[3] = tally[4] = 0;   // ready for the next step
    }
}

// ---- cell exchange over NVLink peer memory (slab_device.PeerSlab) ----------------------------------------------------
// A rank stores the valid records of its two messages straight into its neighbours' receive buffers (CUDA-IPC
// mappings, NVLink) and then releases the step number into each neighbour's "data" flag word; its own stream waits
// until both neighbours' step numbers have arrived in ITS flag words, merges, and acknowledges with the step number in
// the neighbours' "ack" words -- which a sender polls before it overwrites a buffer.  Every poll reads the polling
// device's own memory and gives up after max_wait_ns (a neighbour that died must not hang the GPU), recording the
// step in *timed_out, which the caller reads at its next check.  A flag value is value + *value_dev when value_dev is
// given: a captured graph, whose arguments are frozen, reads the step number from a counter it increments itself.
__device__ __forceinline__ void poll_flag(const int *flag, int value, long long max_wait_ns, int *timed_out) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        int v;
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= value) return;
        __nanosleep(100);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if ((long long)(t - t0) > max_wait_ns) {
            atomicMax(timed_out, value > 0 ? value : 1);
            return;
        }
    }
}

struct PeerSide {
    int4 *dst;               // the neighbour's receive buffer (nullptr: no neighbour on this side)
    const int4 *src;         // my message for it
    int *data_flag;          // the neighbour's "data" flag word for messages from my side
    const int *ack_flag;     // MY flag word the neighbour writes when it has merged a message of mine
};

constexpr int kPeerBlocks = kSMsB200;        // blocks per side

// both sides in one launch: blocks [0, kPeerBlocks) serve the left neighbour, the rest the right one
__global__ void __launch_bounds__(256)
k_peer_exchange(PeerSide left, PeerSide right, long long capacity, int rec_bytes, int value,
                const int *__restrict__ value_dev, int *__restrict__ counters, long long max_wait_ns,
                int *__restrict__ timed_out) {
    const int side = blockIdx.x / kPeerBlocks, blk = blockIdx.x % kPeerBlocks;
    const PeerSide p = side ? right : left;
    if (!p.dst) return;
    if (value_dev) value += *value_dev;
    if (threadIdx.x == 0) poll_flag(p.ack_flag, value - 1, max_wait_ns, timed_out);   // my previous message was merged
    __syncthreads();
    const long long header = capacity * rec_bytes / 16;                              // in 16-byte units
    int count = *reinterpret_cast<const int *>(p.src + header);
    count = count < 0 ? 0 : (count > capacity ? (int)capacity : count);
    const long long body = (long long)count * rec_bytes / 16;
    for (long long k = (long long)blk * 256 + threadIdx.x; k < body; k += (long long)kPeerBlocks * 256) p.dst[k] = p.src[k];
    if (blk == 0 && threadIdx.x < rec_bytes / 16) p.dst[header + threadIdx.x] = p.src[header + threadIdx.x];
    __shared__ bool last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(&counters[side], 1) == kPeerBlocks - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        counters[side] = 0;
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p.data_flag), "r"(value) : "memory");
    }
}

// thread k < 2: wait until my flag word k has reached the value (wait_flags), or release the value into the
// neighbour's word k (release_flags); a null pointer = no neighbour on that side
__global__ void k_peer_flags(const int *wait_a, const int *wait_b, int *release_a, int *release_b, int value,
                             const int *__restrict__ value_dev, long long max_wait_ns, int *timed_out) {
    if (value_dev) value += *value_dev;
    const int *w = threadIdx.x ? wait_b : wait_a;
    int *r = threadIdx.x ? release_b : release_a;
    if (w) poll_flag(w, value, max_wait_ns, timed_out);
    if (r) asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(r), "r"(value) : "memory");
}

}  // namespace abm

ABM_EXPORT int abm_ipc_open(const uint8_t *handle_host, int64_t handle_bytes, void **base_host) {
    ABM_REQUIRE(handle_host && base_host && handle_bytes == (int64_t)sizeof(cudaIpcMemHandle_t), "bad IPC handle");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    // opened with the CONSUMER's device current: the mapping then belongs to this device's context and, the exporter
    // being another GPU, goes over NVLink (peer access is enabled on the way)
    return check_cuda(cudaIpcOpenMemHandle(base_host, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

ABM_EXPORT int abm_peer_exchange(uint8_t *peer_message_left, const uint8_t *message_left, int32_t *peer_data_flag_left,
                                 const int32_t *ack_flag_left, uint8_t *peer_message_right,
                                 const uint8_t *message_right, int32_t *peer_data_flag_right,
                                 const int32_t *ack_flag_right, int64_t capacity, int64_t record_bytes, int32_t value,
                                 const int32_t *value_device, int32_t *block_counters, int64_t max_wait_ns,
                                 int32_t *timed_out, void *stream) {
    ABM_REQUIRE(block_counters && timed_out && max_wait_ns > 0 && capacity >= 0 && record_bytes > 0 &&
                    record_bytes % 16 == 0,
                "bad arguments");
    ABM_REQUIRE(!peer_message_left || (message_left && peer_data_flag_left && ack_flag_left), "left side incomplete");
    ABM_REQUIRE(!peer_message_right || (message_right && peer_data_flag_right && ack_flag_right),
                "right side incomplete");
    ABM_REQUIRE((((uintptr_t)peer_message_left | (uintptr_t)message_left | (uintptr_t)peer_message_right |
                  (uintptr_t)message_right) & 15) == 0, "messages must be 16-byte aligned");
    if (!peer_message_left && !peer_message_right) return 0;
    PeerSide l{reinterpret_cast<int4 *>(peer_message_left), reinterpret_cast<const int4 *>(message_left),
               peer_data_flag_left, ack_flag_left};
    PeerSide r{reinterpret_cast<int4 *>(peer_message_right), reinterpret_cast<const int4 *>(message_right),
               peer_data_flag_right, ack_flag_right};
    k_peer_exchange<<<2 * kPeerBlocks, 256, 0, as_stream(stream)>>>(l, r, (long long)capacity, (int)record_bytes, value,
                                                                    value_device, block_counters,
                                                                    (long long)max_wait_ns, timed_out);
    ABM_CHECK_LAUNCH("k_peer_exchange");
    return 0;
}

ABM_EXPORT int abm_peer_flags(const int32_t *wait_a, const int32_t *wait_b, int32_t *release_a, int32_t *release_b,
                              int32_t value, const int32_t *value_device, int64_t max_wait_ns, int32_t *timed_out,
                              void *stream) {
    ABM_REQUIRE(timed_out && max_wait_ns > 0, "bad arguments");
    if (!wait_a && !wait_b && !release_a && !release_b) return 0;
    k_peer_flags<<<1, 2, 0, as_stream(stream)>>>(wait_a, wait_b, release_a, release_b, value, value_device,
                                                 (long long)max_wait_ns, timed_out);
    ABM_CHECK_LAUNCH("k_peer_flags");
    return 0;
}

ABM_EXPORT int abm_slab_exchange_pack(const double *locations, const double *radii, const double *motility_forces,
                                      const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows,
                                      int64_t rows, const int32_t *n_own_device, int64_t row_stride,
                                      double x_send_left, double x_send_right, int has_left, int has_right,
                                      uint8_t *message_left, uint8_t *message_right, int64_t capacity,
                                      int32_t *counters, int32_t *worst, void *stream) {
    ABM_REQUIRE(rows > 0 && rows < 0x7fffffff && row_stride > 0 && capacity >= 0, "bad sizes");
    ABM_REQUIRE(locations && radii && motility_forces && global_id && bond_len && bond_rows && n_own_device &&
                    message_left && message_right && counters && worst,
                "null pointer");
    CellArrays a{const_cast<double *>(locations), const_cast<double *>(radii), const_cast<double *>(motility_forces),
                 const_cast<int *>(global_id), const_cast<int *>(bond_len), const_cast<int *>(bond_rows)};
    k_slab_pack_hdr<<<ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(
        a, (int)rows, n_own_device, (int)row_stride, 64 + 4 * (int)row_stride, x_send_left, x_send_right, has_left,
        has_right, message_left, message_right, (long long)capacity, counters, worst);
    ABM_CHECK_LAUNCH("k_slab_pack_hdr");
    return 0;
}

ABM_EXPORT int abm_slab_exchange_merge(const double *locations, const double *radii, const double *motility_forces,
                                       const int32_t *global_id, const int32_t *bond_len, const int32_t *bond_rows,
                                       int64_t rows, const int32_t *n_own_device, const uint8_t *message_left,
                                       const uint8_t *message_right, int64_t capacity, int64_t row_stride, double x_lo,
                                       double x_hi, int has_left, int has_right, double *out_locations,
                                       double *out_radii, double *out_motility, int32_t *out_global_id,
                                       int32_t *out_bond_len, int32_t *out_bond_rows, double *snapshot,
                                       int64_t out_rows, int32_t *n_owned_out, int32_t *n_local_out, int32_t *tally,
                                       int32_t *worst_rows, void *stream) {
    ABM_REQUIRE(rows > 0 && rows + 2 * capacity < 0x7fffffff && row_stride > 0 && capacity >= 0 && out_rows > 0,
                "bad sizes");
    ABM_REQUIRE(locations && radii && motility_forces && global_id && bond_len && bond_rows && n_own_device &&
                    out_locations && out_radii && out_motility && out_global_id && out_bond_len && out_bond_rows &&
                    snapshot && n_owned_out && n_local_out && tally && worst_rows,
                "null pointer");
    ABM_REQUIRE(out_locations != locations && out_bond_rows != bond_rows, "merge is out of place");
    ABM_REQUIRE((!has_left || message_left) && (!has_right || message_right), "missing message");
    CellArrays in{const_cast<double *>(locations), const_cast<double *>(radii), const_cast<double *>(motility_forces),
                  const_cast<int *>(global_id), const_cast<int *>(bond_len), const_cast<int *>(bond_rows)};
    CellArrays out{out_locations, out_radii, out_motility, out_global_id, out_bond_len, out_bond_rows};
    Incoming msg{has_left ? message_left : nullptr, has_right ? message_right : nullptr, (int)capacity};
    const int rec_bytes = 64 + 4 * (int)row_stride;
    const int blocks = ceil_div(rows + 2 * capacity, 256);
    cudaStream_t st = as_stream(stream);
    k_slab_count<<<blocks, 256, 0, st>>>(locations, (int)rows, n_own_device, msg, rec_bytes, x_lo, x_hi, has_left,
                                         has_right, tally);
    ABM_CHECK_LAUNCH("k_slab_count");
    k_slab_place<<<blocks, 256, 0, st>>>(in, (int)rows, n_own_device, msg, rec_bytes, (int)row_stride, x_lo, x_hi,
                                         has_left, has_right, out, snapshot, (int)out_rows, tally, n_owned_out,
                                         n_local_out, worst_rows);
    ABM_CHECK_LAUNCH("k_slab_place");
    return 0;
}
