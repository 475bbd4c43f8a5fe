// diffusion3d.cu -- morphogen diffusion on a 3-D grid: 7-point FTCS Jacobi substeps with reflecting walls, and the
// 8-corner per-cell deposit.  The reference's update_diffusion is 2-D only ("currently 2D", cell_backend.py:184);
// this is the extension SURVEY 8 f-4 / the north star ("TMA-staged 2D/3D finite-difference stencil") ask for:
// the algorithm of cell_backend.py:175-204 with a third axis (a = dt*D/h^2, b = 1-6a), restated on the CPU in
// oracle/abm_oracle.c:orc_update_diffusion3d, which this kernel matches bit for bit (same operation order, no FMA).
//
// Layout: field (gx,gy,gz) float64, z contiguous -- the reference's gradient[x][y][z] (cell_methods.py:470-483).
// One launch = one substep.  A CTA owns a (kJ x kK) pencil in (y,z) and marches along x over a chunk of planes: the
// (kJ+2) x (kK+4) neighbourhood of every plane is staged in shared memory exactly once, through a ring of kStages
// planes filled by TMA (cp.async.bulk.tensor.3d, box {z: kK+4, y: kJ+2, x: 1}, mbarrier complete_tx, one elected
// thread) or by plain cooperative loads; a point reads its x-neighbours from the adjacent ring slots.  Per substep
// the field is read 1.33x (pencil halo) and written 1x.  TMA zero-fills what lies outside the field; the reflecting
// wall is applied when reading: a neighbour outside the field is the point itself, which is what re-copying the ghost
// faces every substep does in the oracle.  (The z halo is 2 wide: a TMA box must start on a 16-byte boundary.)
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace abm {

constexpr int k3J = 8;                       // pencil extent in y
constexpr int k3K = 64;                      // pencil extent in z (contiguous)
constexpr int k3HaloK = 2;
constexpr int k3BoxJ = k3J + 2;
constexpr int k3BoxK = k3K + 2 * k3HaloK;    // 68 doubles = 544 B, a multiple of 16
constexpr int k3Stages = 4;
constexpr int k3PlaneBytes = ((k3BoxJ * k3BoxK * 8 + 127) / 128) * 128;
constexpr int k3Threads = 256;
constexpr int k3ChunkX = 32;                 // planes one CTA marches through
constexpr size_t k3Smem = (size_t)k3Stages * k3PlaneBytes + 128 + k3Stages * 8;

struct Diff3Params {
    long long gx, gy, gz;
    int tiles_j, tiles_k, chunks;
    double a, b, cmax, keep;
    int clamp_in, degrade_out;
};

template <bool kTma>
__global__ void __launch_bounds__(k3Threads)
k_diffuse3d(const __grid_constant__ CUtensorMap src_map, const double *__restrict__ src, double *__restrict__ dst,
            Diff3Params P) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~uintptr_t(127));
    uint64_t *bars = (uint64_t *)(base + (size_t)k3Stages * k3PlaneBytes);
    const int tid = threadIdx.x;
    const int chunk = blockIdx.x % P.chunks;
    const int t = blockIdx.x / P.chunks;
    const int tk = t % P.tiles_k, tj = t / P.tiles_k;
    const long long x0 = (long long)chunk * k3ChunkX;
    const long long x1 = min(x0 + k3ChunkX, P.gx);           // planes [x0, x1) are computed
    const long long y0 = (long long)tj * k3J, z0 = (long long)tk * k3K;
    const long long first = max(x0 - 1, 0ll);                // first / last plane that is read; plane p lives in ring
    const long long last = min(x1, P.gx - 1);                // slot (p - first) % kStages

    if (kTma) {
        if (tid == 0) {
            for (int s = 0; s < k3Stages; ++s) mbar_init(&bars[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    auto slot = [&](long long p) { return (int)((p - first) % k3Stages); };
    auto issue = [&](long long p) {                          // plane p -> its ring slot (p inside the field)
        double *box = (double *)(base + (size_t)slot(p) * k3PlaneBytes);
        if (kTma) {
            if (tid == 0) {
                mbar_expect_tx(&bars[slot(p)], k3BoxJ * k3BoxK * 8);
                tma_load_3d(box, &src_map, &bars[slot(p)], (int)z0 - k3HaloK, (int)y0 - 1, (int)p);
            }
        } else {
            for (int e = tid; e < k3BoxJ * k3BoxK; e += k3Threads) {
                int lj = e / k3BoxK, lk = e % k3BoxK;
                long long gj = y0 - 1 + lj, gk = z0 - k3HaloK + lk;
                double v = 0.0;
                if (gj >= 0 && gj < P.gy && gk >= 0 && gk < P.gz) v = src[(p * P.gy + gj) * P.gz + gk];
                box[e] = v;
            }
        }
    };
    auto wait = [&](long long p) {                           // plane p has landed in its slot
        if (kTma) mbar_wait(&bars[slot(p)], (uint32_t)(((p - first) / k3Stages) & 1));
    };

    // prologue: the first kStages planes of the chunk's neighbourhood
    for (long long p = first; p <= last && p < first + k3Stages; ++p) issue(p);
    if (!kTma) __syncthreads();
    wait(first);
    if (x0 > first) wait(x0);

    const int lk = tid & (k3K - 1);
    const long long gk = z0 + lk;
    for (long long x = x0; x < x1; ++x) {
        if (x + 1 <= last) wait(x + 1);
        const double *pc = (const double *)(base + (size_t)slot(x) * k3PlaneBytes);
        const double *pm = x > 0 ? (const double *)(base + (size_t)slot(x - 1) * k3PlaneBytes) : pc;
        const double *pp = x < P.gx - 1 ? (const double *)(base + (size_t)slot(x + 1) * k3PlaneBytes) : pc;
        if (gk < P.gz) {
            const int kl = (gk == 0) ? 0 : -1, kr = (gk == P.gz - 1) ? 0 : 1;
            for (int lj = tid / k3K; lj < k3J; lj += k3Threads / k3K) {
                const long long gj = y0 + lj;
                if (gj >= P.gy) break;
                const int jd = (gj == 0) ? 0 : -1, ju = (gj == P.gy - 1) ? 0 : 1;
                const int o = (lj + 1) * k3BoxK + (lk + k3HaloK);
                double vc = pc[o], vxp = pp[o], vxm = pm[o], vyp = pc[o + ju * k3BoxK], vym = pc[o + jd * k3BoxK],
                       vzp = pc[o + kr], vzm = pc[o + kl];
                if (P.clamp_in) {
                    vc = fmin(fmax(vc, 0.0), P.cmax);
                    vxp = fmin(fmax(vxp, 0.0), P.cmax);
                    vxm = fmin(fmax(vxm, 0.0), P.cmax);
                    vyp = fmin(fmax(vyp, 0.0), P.cmax);
                    vym = fmin(fmax(vym, 0.0), P.cmax);
                    vzp = fmin(fmax(vzp, 0.0), P.cmax);
                    vzm = fmin(fmax(vzm, 0.0), P.cmax);
                }
                // oracle order: t = a*(((((x+1 + x-1) + y+1) + y-1) + z+1) + z-1); new = centre*b + t   (no FMA)
                double sum = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(vxp, vxm), vyp), vym), vzp), vzm);
                double v = __dadd_rn(__dmul_rn(vc, P.b), __dmul_rn(P.a, sum));
                if (P.degrade_out) v = __dmul_rn(v, P.keep);
                dst[(x * P.gy + gj) * P.gz + gk] = v;
            }
        }
        __syncthreads();                                     // plane x-1 is no longer needed by anyone
        const long long nxt = x - 1 + k3Stages;              // ... its slot takes the plane kStages further on
        if (x - 1 >= first && nxt <= last) {                 // (x-1 < first only for the very first plane of the field)
            issue(nxt);
            if (!kTma) __syncthreads();
        }
    }
}

// adjust_morphogens over the EIGHT corners of the cell's grid cube (oracle: orc_adjust_morphogens3d): the reference
// deposits on the z = 0 plane only (cell_methods.py:499), which k_deposit keeps for (gx,gy,1) fields.
__global__ void k_deposit3d(const double *__restrict__ loc, const double *__restrict__ amounts, double uniform_amount,
                            int n, double *__restrict__ field, long long g0, long long g1, long long g2, double h,
                            long long row_lo, long long row_hi, const int *__restrict__ n_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    n = device_count(n, n_dev);
    if (i >= n) return;
    const double l[3] = {loc[3 * (long long)i + 0], loc[3 * (long long)i + 1], loc[3 * (long long)i + 2]};
    const long long g[3] = {g0, g1, g2};
    long long c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) c[k] = (long long)floor(l[k] / h);
    unsigned near = 0;
    int total = 0;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        long long q[3] = {c[0] + (p & 1), c[1] + ((p >> 1) & 1), c[2] + ((p >> 2) & 1)};
        if (q[0] < g[0] && q[1] < g[1] && q[2] < g[2]) {
            double dx = l[0] - __dmul_rn((double)q[0], h), dy = l[1] - __dmul_rn((double)q[1], h),
                   dz = l[2] - __dmul_rn((double)q[2], h);
            if (__dsqrt_rn(sumsq_nofma(dx, dy, dz)) < h) {
                near |= 1u << p;
                ++total;
            }
        }
    }
    if (total == 0) return;
    double share = __ddiv_rn(amounts ? amounts[i] : uniform_amount, (double)total);
#pragma unroll
    for (int p = 0; p < 8; ++p) {
        if (!((near >> p) & 1u)) continue;
        long long q[3] = {c[0] + (p & 1), c[1] + ((p >> 1) & 1), c[2] + ((p >> 2) & 1)};
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (q[k] < 0) q[k] += g[k];                       // python negative-index wraparound, as the oracle
        if (q[0] < row_lo || q[0] >= row_hi || q[1] < 0 || q[2] < 0) continue;      // planes held by another slab
        atomicAdd(&field[((q[0] - row_lo) * g1 + q[1]) * g2 + q[2]], share);
    }
}

static int make_field_map3d(CUtensorMap *map, const double *ptr, long long gx, long long gy, long long gz) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return -1;
    cuuint64_t gdim[3] = {(cuuint64_t)gz, (cuuint64_t)gy, (cuuint64_t)gx};
    cuuint64_t gstride[2] = {(cuuint64_t)gz * 8, (cuuint64_t)gy * gz * 8};
    cuuint32_t box[3] = {(cuuint32_t)k3BoxK, (cuuint32_t)k3BoxJ, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)ptr, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d (gx=%lld gy=%lld gz=%lld)", (int)r, gx, gy, gz);
        return -1;
    }
    return 0;
}

}  // namespace abm

using namespace abm;

ABM_EXPORT int abm_diffusion3d(double *field, double *scratch, int64_t gx, int64_t gy, int64_t gz, int64_t steps,
                               double diffuse_dt, double last_dt, double diffuse_const, double spat_res2,
                               double max_concentration, double degradation, int use_tma,
                               int *result_in_scratch_host, void *stream) {
    ABM_REQUIRE(field && scratch && field != scratch, "field/scratch must be distinct non-null buffers");
    ABM_REQUIRE(gx >= 1 && gy >= 1 && gz >= 1 && gx * gy * gz < (1ll << 40) && gx < 0x7fffffff, "bad grid size");
    ABM_REQUIRE(steps >= 1, "steps must be >= 1");
    cudaStream_t st = as_stream(stream);
    CUtensorMap maps[2];
    memset(maps, 0, sizeof(maps));
    if (use_tma) {
        ABM_REQUIRE((gz % 2) == 0, "TMA path needs an even gz (row pitch must be a multiple of 16 bytes)");
        ABM_REQUIRE((((uintptr_t)field | (uintptr_t)scratch) & 15) == 0, "TMA path needs 16-byte aligned fields");
        int rc = make_field_map3d(&maps[0], field, gx, gy, gz);
        if (rc) return rc;
        rc = make_field_map3d(&maps[1], scratch, gx, gy, gz);
        if (rc) return rc;
    }
    static bool attr_set[kMaxDevices][2] = {};
    const int slot = current_device_slot();
    if (slot < 0 || !attr_set[slot][use_tma ? 1 : 0]) {
        cudaError_t e = use_tma ? cudaFuncSetAttribute(k_diffuse3d<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)k3Smem)
                                : cudaFuncSetAttribute(k_diffuse3d<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)k3Smem);
        int rc = check_cuda(e, "cudaFuncSetAttribute(k_diffuse3d)");
        if (rc) return rc;
        if (slot >= 0) attr_set[slot][use_tma ? 1 : 0] = true;
    }
    Diff3Params P;
    P.gx = gx;
    P.gy = gy;
    P.gz = gz;
    P.tiles_j = ceil_div(gy, k3J);
    P.tiles_k = ceil_div(gz, k3K);
    P.chunks = ceil_div(gx, k3ChunkX);
    P.cmax = max_concentration;
    P.keep = 1.0 - degradation;
    const long long grid = (long long)P.tiles_j * P.tiles_k * P.chunks;
    ABM_REQUIRE(grid < 0x7fffffff, "grid too large");
    double *bufs[2] = {field, scratch};
    for (int64_t k = 0; k < steps; ++k) {
        const int from = (int)(k & 1), to = from ^ 1;
        const double dt = (k == steps - 1) ? last_dt : diffuse_dt;
        P.a = dt * diffuse_const / spat_res2;
        P.b = 1 - 6 * P.a;
        P.clamp_in = (k == 0);
        P.degrade_out = (k == steps - 1);
        if (use_tma)
            k_diffuse3d<true><<<(unsigned)grid, k3Threads, k3Smem, st>>>(maps[from], bufs[from], bufs[to], P);
        else
            k_diffuse3d<false><<<(unsigned)grid, k3Threads, k3Smem, st>>>(maps[from], bufs[from], bufs[to], P);
        ABM_CHECK_LAUNCH("k_diffuse3d");
    }
    const int in_scratch = (int)(steps & 1);
    if (result_in_scratch_host) {
        *result_in_scratch_host = in_scratch;
    } else if (in_scratch) {
        int rc = check_cuda(cudaMemcpyAsync(field, scratch, (size_t)gx * gy * gz * 8, cudaMemcpyDeviceToDevice, st),
                            "copy diffusion result back");
        if (rc) return rc;
    }
    return 0;
}

ABM_EXPORT int abm_deposit_morphogens3d(const double *locations, const double *amounts, double uniform_amount,
                                        int64_t n, double *field, const int64_t gsize_host[3], double spat_res,
                                        int64_t row_lo, int64_t row_hi, const int32_t *n_device, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(locations && field && gsize_host, "null pointer");
    ABM_REQUIRE(spat_res > 0 && gsize_host[0] > 0 && gsize_host[1] > 0 && gsize_host[2] > 0, "bad grid");
    ABM_REQUIRE(row_lo >= 0 && row_hi <= gsize_host[0] && row_lo < row_hi, "bad row window");
    k_deposit3d<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(locations, amounts, uniform_amount, (int)n, field,
                                                                 gsize_host[0], gsize_host[1], gsize_host[2],
                                                                 spat_res, row_lo, row_hi, n_device);
    ABM_CHECK_LAUNCH("k_deposit3d");
    return 0;
}
