// diffusion.cu -- FGF4 morphogen field: FTCS 5-point Jacobi substeps with reflecting walls, plus the
// per-cell sample / deposit coupling.  Replaces CellMethods.update_diffusion + update_diffusion_jit,
// get_concentration and adjust_morphogens (reference cell_methods.py:441-521, cell_backend.py:175-204).
//
// One launch = one substep.  Persistent CTAs walk (kTileI x kTileJ) output tiles; the (kTileI+2) x
// (kTileJ+4) input box of each tile is staged in shared memory through a kStages-deep ring, filled either
// by TMA (cp.async.bulk.tensor.2d + mbarrier complete_tx, one elected thread) or by plain cooperative loads.
// TMA zero-fills the part of a box outside the field; the reflecting wall is applied when reading (a
// neighbour outside the field is the point itself, which is what re-copying the ghost ring every substep
// does in the reference, cell_backend.py:192-195).
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace abm {

constexpr int kTileI = 32;                 // rows (slow axis, stride gy)
constexpr int kTileJ = 128;                // columns (contiguous axis)
constexpr int kHaloJ = 2;                  // TMA needs the box to START on a 16-byte boundary in global memory:
                                           // with 8-byte elements the inner start coordinate must be even, so the
                                           // column halo is 2 wide on each side although the stencil needs 1
constexpr int kBoxI = kTileI + 2;
constexpr int kBoxJ = kTileJ + 2 * kHaloJ; // 132 doubles = 1056 B, a multiple of 16 as TMA requires
constexpr int kStages = 3;
constexpr int kStageBytes = ((kBoxI * kBoxJ * 8 + 127) / 128) * 128;
constexpr int kDiffThreads = 256;
constexpr size_t kDiffSmem = (size_t)kStages * kStageBytes + 128 /*alignment slack*/ + kStages * 8;

struct DiffParams {
    long long gx, gy;
    int tiles_i, tiles_j;
    double a, b;            // a = dt*D/h^2, b = 1-4a (reference cell_backend.py:181-182)
    double cmax;            // clamp applied to the SOURCE when clamp_in
    double keep;            // 1-degradation applied to the RESULT when degrade_out
    int clamp_in, degrade_out;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

template <bool kTma>
__global__ void __launch_bounds__(kDiffThreads)
k_diffuse(const __grid_constant__ CUtensorMap src_map, const double *__restrict__ src, double *__restrict__ dst,
          DiffParams P) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~uintptr_t(127));
    uint64_t *bars = (uint64_t *)(base + (size_t)kStages * kStageBytes);
    const int tid = threadIdx.x;
    const int ntiles = P.tiles_i * P.tiles_j;

    if (kTma) {
        if (tid == 0) {
            for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    auto issue = [&](int tile, int stage) {
        int ti = tile / P.tiles_j, tj = tile % P.tiles_j;
        double *box = (double *)(base + (size_t)stage * kStageBytes);
        if (kTma) {
            if (tid == 0) {
                mbar_expect_tx(&bars[stage], kBoxI * kBoxJ * 8);
                tma_load_2d(box, &src_map, &bars[stage], tj * kTileJ - kHaloJ, ti * kTileI - 1);
            }
        } else {
            long long i0 = (long long)ti * kTileI - 1, j0 = (long long)tj * kTileJ - kHaloJ;
            for (int e = tid; e < kBoxI * kBoxJ; e += kDiffThreads) {
                int li = e / kBoxJ, lj = e % kBoxJ;
                long long gi = i0 + li, gj = j0 + lj;
                double v = 0.0;
                if (gi >= 0 && gi < P.gx && gj >= 0 && gj < P.gy) v = src[gi * P.gy + gj];
                box[e] = v;
            }
        }
    };

    // prologue: fill the ring
    int my_count = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_count;
    for (int k = 0; k < kStages && k < my_count; ++k) issue(blockIdx.x + k * gridDim.x, k);

    for (int k = 0; k < my_count; ++k) {
        int tile = blockIdx.x + k * gridDim.x;
        int stage = k % kStages;
        if (kTma) mbar_wait(&bars[stage], (k / kStages) & 1);
        else __syncthreads();
        const double *box = (const double *)(base + (size_t)stage * kStageBytes);
        int ti = tile / P.tiles_j, tj = tile % P.tiles_j;
        // thread -> column lj (0..127), rows interleaved by 2
        int lj = tid & (kTileJ - 1);
        long long gj = (long long)tj * kTileJ + lj;
        if (gj < P.gy) {
            int jl = (gj == 0) ? 0 : -1;              // reflecting wall: outside neighbour = the point itself
            int jr = (gj == P.gy - 1) ? 0 : 1;
            for (int li = tid / kTileJ; li < kTileI; li += kDiffThreads / kTileJ) {
                long long gi = (long long)ti * kTileI + li;
                if (gi >= P.gx) break;
                int iu = (gi == 0) ? 0 : -1;
                int id = (gi == P.gx - 1) ? 0 : 1;
                const double *c = box + (li + 1) * kBoxJ + (lj + kHaloJ);
                double vc = c[0], vd = c[id * kBoxJ], vu = c[iu * kBoxJ], vr = c[jr], vl = c[jl];
                if (P.clamp_in) {
                    vc = fmin(fmax(vc, 0.0), P.cmax);
                    vd = fmin(fmax(vd, 0.0), P.cmax);
                    vu = fmin(fmax(vu, 0.0), P.cmax);
                    vr = fmin(fmax(vr, 0.0), P.cmax);
                    vl = fmin(fmax(vl, 0.0), P.cmax);
                }
                // reference order: temp = a*(((down+up)+right)+left); new = center*b + temp   (no FMA)
                double t = __dmul_rn(P.a, __dadd_rn(__dadd_rn(__dadd_rn(vd, vu), vr), vl));
                double v = __dadd_rn(__dmul_rn(vc, P.b), t);
                if (P.degrade_out) v = __dmul_rn(v, P.keep);
                dst[gi * P.gy + gj] = v;
            }
        }
        __syncthreads();                               // everyone is done reading this stage
        if (k + kStages < my_count) issue(blockIdx.x + (k + kStages) * gridDim.x, stage);
    }
}

// `field` holds global rows [row_lo,row_hi) of a (g0,g1,g2) grid (a slab of it, or all of it); indices are computed
// on the global grid exactly as the reference does and then shifted.
__global__ void k_sample(const double *__restrict__ loc, int n, const double *__restrict__ field, long long g0,
                         long long g1, long long g2, double h, long long row_lo, long long row_hi,
                         double *__restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long g[3] = {g0, g1, g2};
    long long idx[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double half = floor(__ddiv_rn(__dmul_rn(2.0, loc[3 * (long long)i + k]), h));
        long long v = (long long)ceil(half / 2.0);
        if (v < 0) v += g[k];                          // python negative-index wraparound
        v = v < 0 ? 0 : (v >= g[k] ? g[k] - 1 : v);   // the reference would raise IndexError beyond this
        idx[k] = v;
    }
    long long r = idx[0] < row_lo ? row_lo : (idx[0] >= row_hi ? row_hi - 1 : idx[0]);
    out[i] = field[((r - row_lo) * g1 + idx[1]) * g2 + idx[2]];
}

// amounts == nullptr: every cell deposits uniform_amount
__global__ void k_deposit(const double *__restrict__ loc, const double *__restrict__ amounts, double uniform_amount,
                          int n, double *__restrict__ field, long long g0, long long g1, long long g2, double h,
                          long long row_lo, long long row_hi) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double lx = loc[3 * (long long)i + 0], ly = loc[3 * (long long)i + 1], lz = loc[3 * (long long)i + 2];
    long long x = (long long)floor(lx / h), y = (long long)floor(ly / h);
    int total = 0;
    bool near[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        long long px = x + (p & 1), py = y + (p >> 1);
        near[p] = false;
        if (px < g0 && py < g1) {
            double dx = lx - __dmul_rn((double)px, h), dy = ly - __dmul_rn((double)py, h), dz = lz - 0.0;
            if (__dsqrt_rn(sumsq_nofma(dx, dy, dz)) < h) {
                near[p] = true;
                ++total;
            }
        }
    }
    if (total == 0) return;
    double share = __ddiv_rn(amounts ? amounts[i] : uniform_amount, (double)total);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        if (!near[p]) continue;
        long long px = x + (p & 1), py = y + (p >> 1);
        if (px < 0) px += g0;
        if (py < 0) py += g1;
        if (px < row_lo || px >= row_hi || py < 0) continue;   // rows held by another slab
        atomicAdd(&field[((px - row_lo) * g1 + py) * g2 + 0], share);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_field_map(CUtensorMap *map, const double *ptr, long long gx, long long gy) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        int rc = check_cuda(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q),
                            "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
        if (rc) return rc;
        if (!fn || q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available from this driver");
            return -1;
        }
        encode = (EncodeTiledFn)fn;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)gy, (cuuint64_t)gx};
    cuuint64_t gstride[1] = {(cuuint64_t)gy * 8};
    cuuint32_t box[2] = {(cuuint32_t)kBoxJ, (cuuint32_t)kBoxI};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)ptr, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (gx=%lld gy=%lld)", (int)r, gx, gy);
        return -1;
    }
    return 0;
}

}  // namespace abm

using namespace abm;

ABM_EXPORT int abm_diffusion(double *field, double *scratch, int64_t gx, int64_t gy, int64_t steps,
                             double diffuse_dt, double last_dt, double diffuse_const, double spat_res2,
                             double max_concentration, double degradation, int use_tma,
                             int *result_in_scratch_host, void *stream) {
    ABM_REQUIRE(field && scratch && field != scratch, "field/scratch must be distinct non-null buffers");
    ABM_REQUIRE(gx >= 1 && gy >= 1 && gx * gy < (1ll << 40), "bad grid size");
    ABM_REQUIRE(steps >= 1, "steps must be >= 1 (the reference always runs int(step_dt // diffuse_dt) + 1)");
    cudaStream_t st = as_stream(stream);
    static int sm_count = 0;
    if (!sm_count) {
        int rc = abm_device_info(&sm_count, nullptr, nullptr);
        if (rc) return rc;
    }
    CUtensorMap maps[2];
    memset(maps, 0, sizeof(maps));
    if (use_tma) {
        ABM_REQUIRE((gy % 2) == 0, "TMA path needs an even gy (row pitch must be a multiple of 16 bytes)");
        ABM_REQUIRE((((uintptr_t)field | (uintptr_t)scratch) & 15) == 0, "TMA path needs 16-byte aligned planes");
        int rc = make_field_map(&maps[0], field, gx, gy);
        if (rc) return rc;
        rc = make_field_map(&maps[1], scratch, gx, gy);
        if (rc) return rc;
    }
    static bool attr_set[2] = {false, false};
    if (!attr_set[use_tma ? 1 : 0]) {
        cudaError_t e = use_tma ? cudaFuncSetAttribute(k_diffuse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)kDiffSmem)
                                : cudaFuncSetAttribute(k_diffuse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)kDiffSmem);
        int rc = check_cuda(e, "cudaFuncSetAttribute(k_diffuse)");
        if (rc) return rc;
        attr_set[use_tma ? 1 : 0] = true;
    }
    DiffParams P;
    P.gx = gx;
    P.gy = gy;
    P.tiles_i = ceil_div(gx, kTileI);
    P.tiles_j = ceil_div(gy, kTileJ);
    P.cmax = max_concentration;
    P.keep = 1.0 - degradation;
    int ntiles = P.tiles_i * P.tiles_j;
    int grid = ntiles < 2 * sm_count ? ntiles : 2 * sm_count;   // two resident CTAs per SM (106 KB each)
    double *bufs[2] = {field, scratch};
    for (int64_t k = 0; k < steps; ++k) {
        double dt = (k == steps - 1) ? last_dt : diffuse_dt;
        P.a = dt * diffuse_const / spat_res2;
        P.b = 1 - 4 * P.a;
        P.clamp_in = (k == 0);
        P.degrade_out = (k == steps - 1);
        int from = (int)(k & 1), to = from ^ 1;
        if (use_tma)
            k_diffuse<true><<<grid, kDiffThreads, kDiffSmem, st>>>(maps[from], bufs[from], bufs[to], P);
        else
            k_diffuse<false><<<grid, kDiffThreads, kDiffSmem, st>>>(maps[from], bufs[from], bufs[to], P);
        ABM_CHECK_LAUNCH("k_diffuse");
    }
    int in_scratch = (int)(steps & 1);
    if (result_in_scratch_host) {
        *result_in_scratch_host = in_scratch;
    } else if (in_scratch) {
        int rc = check_cuda(cudaMemcpyAsync(field, scratch, (size_t)gx * gy * 8, cudaMemcpyDeviceToDevice, st),
                            "copy diffusion result back");
        if (rc) return rc;
    }
    return 0;
}

ABM_EXPORT int abm_sample_concentrations(const double *locations, int64_t n, const double *field,
                                         const int64_t gsize_host[3], double spat_res, int64_t row_lo,
                                         int64_t row_hi, double *sampled, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(locations && field && gsize_host && sampled, "null pointer");
    ABM_REQUIRE(spat_res > 0 && gsize_host[0] > 0 && gsize_host[1] > 0 && gsize_host[2] > 0, "bad grid");
    ABM_REQUIRE(row_lo >= 0 && row_hi <= gsize_host[0] && row_lo < row_hi, "bad row window");
    k_sample<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(locations, (int)n, field, gsize_host[0],
                                                              gsize_host[1], gsize_host[2], spat_res, row_lo,
                                                              row_hi, sampled);
    ABM_CHECK_LAUNCH("k_sample");
    return 0;
}

ABM_EXPORT int abm_deposit_morphogens(const double *locations, const double *amounts, double uniform_amount,
                                      int64_t n, double *field, const int64_t gsize_host[3], double spat_res,
                                      int64_t row_lo, int64_t row_hi, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(locations && field && gsize_host, "null pointer");
    ABM_REQUIRE(spat_res > 0 && gsize_host[0] > 0 && gsize_host[1] > 0 && gsize_host[2] > 0, "bad grid");
    ABM_REQUIRE(row_lo >= 0 && row_hi <= gsize_host[0] && row_lo < row_hi, "bad row window");
    k_deposit<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(locations, amounts, uniform_amount, (int)n, field,
                                                               gsize_host[0],
                                                               gsize_host[1], gsize_host[2], spat_res, row_lo,
                                                               row_hi);
    ABM_CHECK_LAUNCH("k_deposit");
    return 0;
}
