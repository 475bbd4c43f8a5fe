// diffusion.cu -- FGF4 morphogen field: FTCS 5-point Jacobi substeps with reflecting walls, plus the
// per-cell sample / deposit coupling.  Replaces CellMethods.update_diffusion + update_diffusion_jit,
// get_concentration and adjust_morphogens (reference cell_methods.py:441-521, cell_backend.py:175-204).
//
// Two kernels.  k_diffuse_fused<N> advances N = 2..5 substeps per launch (temporal blocking): a CTA stages the
// (kTileI+2N) x (kTileJ+2N) neighbourhood of its output tile in shared memory (one TMA box, or plain loads), runs
// the substeps there on a region that shrinks by one ring per substep, and writes the tile once -- 2wG bytes of
// field traffic and one launch per N substeps, for ~1.4x the arithmetic.  Every point still sees exactly the
// reference's operations in the reference's order, so the result is bit-identical to substep-by-substep launches.
// k_diffuse does a single substep (chains of length 1, and max_fused_substeps = 1):
// One launch = one substep.  Persistent CTAs walk (kTileI x kTileJ) output tiles; the (kTileI+2) x
// (kTileJ+4) input box of each tile is staged in shared memory through a kStages-deep ring, filled either
// by TMA (cp.async.bulk.tensor.2d + mbarrier complete_tx, one elected thread) or by plain cooperative loads.
// TMA zero-fills the part of a box outside the field; the reflecting wall is applied when reading (a
// neighbour outside the field is the point itself, which is what re-copying the ghost ring every substep
// does in the reference, cell_backend.py:192-195).
#include <cuda.h>
#include <string.h>

#include "common.cuh"
#include "tma.cuh"

namespace abm {

// Build-time tile shape (scripts/ab_substep.py builds variants): kTileJ must be a power of two <= 256 that divides
// kDiffThreads or is a multiple of it; the shipped library uses 32 x 128.
#ifndef ABM_DIFF_TILE_I
#define ABM_DIFF_TILE_I 32
#endif
#ifndef ABM_DIFF_TILE_J
#define ABM_DIFF_TILE_J 128
#endif
constexpr int kTileI = ABM_DIFF_TILE_I;    // rows (slow axis, stride gy)
constexpr int kTileJ = ABM_DIFF_TILE_J;    // columns (contiguous axis)
static_assert((kTileJ & (kTileJ - 1)) == 0 && kTileJ >= 32 && kTileJ <= 256 && kTileI >= 1 && kTileI <= 200, "tile shape");
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

// ---- temporal blocking: N substeps per launch -------------------------------------------------------------------
constexpr int kMaxFuse = 5;

struct FuseParams {
    long long gx, gy;
    int tiles_j;
    double a[kMaxFuse], b[kMaxFuse];   // per substep of this launch (the chain's last substep has its own dt)
    double cmax, keep;
    int clamp_in, degrade_out;         // this launch starts / ends the chain
};

template <int N>
struct FuseGeom {
    static constexpr int kHaloI = N;
    static constexpr int kHaloJ = (N + 1) & ~1;          // even: the TMA box must start on a 16-byte boundary
    static constexpr int kColSkip = kHaloJ - N;           // staged but unused columns on each side (0 or 1)
    static constexpr int kBoxI = kTileI + 2 * kHaloI;
    static constexpr int kBoxJ = kTileJ + 2 * kHaloJ;
    static constexpr int kBufBytes = ((kBoxI * kBoxJ * 8 + 127) / 128) * 128;
    static constexpr size_t kSmem = 2 * (size_t)kBufBytes + 128 /*alignment slack*/ + 8;
};

template <int N, bool kTma>
__global__ void __launch_bounds__(kDiffThreads)
k_diffuse_fused(const __grid_constant__ CUtensorMap src_map, const double *__restrict__ src, double *__restrict__ dst,
                FuseParams P) {
    using G = FuseGeom<N>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *base = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~uintptr_t(127));
    double *cur = (double *)base, *nxt = (double *)(base + G::kBufBytes);
    uint64_t *bar = (uint64_t *)(base + 2 * (size_t)G::kBufBytes);
    const int tid = threadIdx.x;
    const int ti = blockIdx.x / P.tiles_j, tj = blockIdx.x % P.tiles_j;
    // global coordinates of the box element (0,0)
    const long long i0 = (long long)ti * kTileI - G::kHaloI, j0 = (long long)tj * kTileJ - G::kHaloJ;

    if (kTma) {
        if (tid == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(bar, G::kBoxI * G::kBoxJ * 8);
            tma_load_2d(cur, &src_map, bar, (int)j0, (int)i0);       // outside the field: zero fill (never read)
        }
        mbar_wait(bar, 0);
    } else {
        for (int e = tid; e < G::kBoxI * G::kBoxJ; e += kDiffThreads) {
            const int li = e / G::kBoxJ, lj = e % G::kBoxJ;
            const long long gi = i0 + li, gj = j0 + lj;
            double v = 0.0;
            if (gi >= 0 && gi < P.gx && gj >= 0 && gj < P.gy) v = src[gi * P.gy + gj];
            cur[e] = v;
        }
        __syncthreads();
    }
    if (P.clamp_in) {                                      // reference cell_methods.py:454-455, once per chain
        for (int e = tid; e < G::kBoxI * G::kBoxJ; e += kDiffThreads) cur[e] = fmin(fmax(cur[e], 0.0), P.cmax);
        __syncthreads();
    }

#pragma unroll
    for (int s = 1; s <= N; ++s) {
        // substep s is evaluated on the tile grown by N-s rings; its inputs (substep s-1 on N-s+1 rings) are in `cur`
        const int rows = G::kBoxI - 2 * s, cols = kTileJ + 2 * (N - s);
        const int l0 = s, c0 = G::kColSkip + s;
        const double a = P.a[s - 1], b = P.b[s - 1];
        for (int e = tid; e < rows * cols; e += kDiffThreads) {
            const int r = e / cols;
            const int li = l0 + r, lj = c0 + (e - r * cols);
            const long long gi = i0 + li, gj = j0 + lj;
            if (gi < 0 || gi >= P.gx || gj < 0 || gj >= P.gy) continue;
            const int iu = (gi == 0) ? 0 : -1, id = (gi == P.gx - 1) ? 0 : 1;     // reflecting wall: the point itself
            const int jl = (gj == 0) ? 0 : -1, jr = (gj == P.gy - 1) ? 0 : 1;
            const double *c = cur + li * G::kBoxJ + lj;
            const double vc = c[0], vd = c[id * G::kBoxJ], vu = c[iu * G::kBoxJ], vr = c[jr], vl = c[jl];
            // reference order: temp = a*(((down+up)+right)+left); new = center*b + temp   (no FMA)
            const double t = __dmul_rn(a, __dadd_rn(__dadd_rn(__dadd_rn(vd, vu), vr), vl));
            double v = __dadd_rn(__dmul_rn(vc, b), t);
            if (s == N) {
                if (P.degrade_out) v = __dmul_rn(v, P.keep);
                dst[gi * P.gy + gj] = v;
            } else {
                nxt[li * G::kBoxJ + lj] = v;
            }
        }
        if (s < N) {
            __syncthreads();
            double *tmp = cur;
            cur = nxt;
            nxt = tmp;
        }
    }
}

// `field` holds global rows [row_lo,row_hi) of a (g0,g1,g2) grid (a slab of it, or all of it); indices are computed
// on the global grid exactly as the reference does and then shifted.
__global__ void k_sample(const double *__restrict__ loc, int n, const double *__restrict__ field, long long g0,
                         long long g1, long long g2, double h, long long row_lo, long long row_hi,
                         double *__restrict__ out, const int *__restrict__ n_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    n = device_count(n, n_dev);
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
                          long long row_lo, long long row_hi, const int *__restrict__ n_dev) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    n = device_count(n, n_dev);
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

static int make_field_map(CUtensorMap *map, const double *ptr, long long gx, long long gy, int box_j = kBoxJ,
                          int box_i = kBoxI) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode) return -1;
    cuuint64_t gdim[2] = {(cuuint64_t)gy, (cuuint64_t)gx};
    cuuint64_t gstride[1] = {(cuuint64_t)gy * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_j, (cuuint32_t)box_i};
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


template <int N>
static int launch_fused(const double *from, double *to, long long gx, long long gy, int use_tma, const FuseParams &P,
                        int ntiles, cudaStream_t st) {
    using G = FuseGeom<N>;
    static bool attr_set[kMaxDevices][2] = {};          // per device (a function attribute is per device)
    const int slot = current_device_slot();
    if (slot < 0 || !attr_set[slot][use_tma ? 1 : 0]) {
        cudaError_t e = use_tma ? cudaFuncSetAttribute(k_diffuse_fused<N, true>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem)
                                : cudaFuncSetAttribute(k_diffuse_fused<N, false>,
                                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::kSmem);
        int rc = check_cuda(e, "cudaFuncSetAttribute(k_diffuse_fused)");
        if (rc) return rc;
        if (slot >= 0) attr_set[slot][use_tma ? 1 : 0] = true;
    }
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (use_tma) {
        int rc = make_field_map(&map, from, gx, gy, G::kBoxJ, G::kBoxI);
        if (rc) return rc;
        k_diffuse_fused<N, true><<<ntiles, kDiffThreads, G::kSmem, st>>>(map, from, to, P);
    } else {
        k_diffuse_fused<N, false><<<ntiles, kDiffThreads, G::kSmem, st>>>(map, from, to, P);
    }
    ABM_CHECK_LAUNCH("k_diffuse_fused");
    return 0;
}

}  // namespace abm

using namespace abm;

ABM_EXPORT int abm_diffusion(double *field, double *scratch, int64_t gx, int64_t gy, int64_t steps,
                             double diffuse_dt, double last_dt, double diffuse_const, double spat_res2,
                             double max_concentration, double degradation, int use_tma, int max_fused_substeps,
                             int *result_in_scratch_host, void *stream) {
    ABM_REQUIRE(field && scratch && field != scratch, "field/scratch must be distinct non-null buffers");
    ABM_REQUIRE(gx >= 1 && gy >= 1 && gx * gy < (1ll << 40), "bad grid size");
    ABM_REQUIRE(steps >= 1, "steps must be >= 1 (the reference always runs int(step_dt // diffuse_dt) + 1)");
    cudaStream_t st = as_stream(stream);
    static int sm_counts[kMaxDevices] = {};             // per device
    const int slot = current_device_slot();
    int sm_count = slot >= 0 ? sm_counts[slot] : 0;
    if (!sm_count) {
        int rc = abm_device_info(&sm_count, nullptr, nullptr);
        if (rc) return rc;
        if (slot >= 0) sm_counts[slot] = sm_count;
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
    static bool attr_set[kMaxDevices][2] = {};
    if (slot < 0 || !attr_set[slot][use_tma ? 1 : 0]) {
        cudaError_t e = use_tma ? cudaFuncSetAttribute(k_diffuse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)kDiffSmem)
                                : cudaFuncSetAttribute(k_diffuse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                       (int)kDiffSmem);
        int rc = check_cuda(e, "cudaFuncSetAttribute(k_diffuse)");
        if (rc) return rc;
        if (slot >= 0) attr_set[slot][use_tma ? 1 : 0] = true;
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
    // Chains of up to max_fused_substeps substeps per launch.  The number of launches keeps the parity of `steps`, so
    // the result lands in the buffer the substep-by-substep chain would leave it in.
    const int64_t launches = abm_diffusion_launches(steps, max_fused_substeps);
    const int64_t per = steps / launches, longer = steps % launches;    // the first `longer` launches take per+1
    int64_t k = 0;
    for (int64_t l = 0; l < launches; ++l) {
        const int len = (int)(per + (l < longer ? 1 : 0));
        const int from = (int)(l & 1), to = from ^ 1;
        if (len == 1) {
            double dt = (k == steps - 1) ? last_dt : diffuse_dt;
            P.a = dt * diffuse_const / spat_res2;
            P.b = 1 - 4 * P.a;
            P.clamp_in = (k == 0);
            P.degrade_out = (k == steps - 1);
            if (use_tma)
                k_diffuse<true><<<grid, kDiffThreads, kDiffSmem, st>>>(maps[from], bufs[from], bufs[to], P);
            else
                k_diffuse<false><<<grid, kDiffThreads, kDiffSmem, st>>>(maps[from], bufs[from], bufs[to], P);
            ABM_CHECK_LAUNCH("k_diffuse");
        } else {
            FuseParams F;
            F.gx = gx;
            F.gy = gy;
            F.tiles_j = P.tiles_j;
            F.cmax = P.cmax;
            F.keep = P.keep;
            for (int s = 0; s < kMaxFuse; ++s) {
                double dt = (k + s == steps - 1) ? last_dt : diffuse_dt;
                F.a[s] = dt * diffuse_const / spat_res2;
                F.b[s] = 1 - 4 * F.a[s];
            }
            F.clamp_in = (k == 0);
            F.degrade_out = (k + len == steps);
            int rc = 0;
            switch (len) {
                case 2: rc = launch_fused<2>(bufs[from], bufs[to], gx, gy, use_tma, F, ntiles, st); break;
                case 3: rc = launch_fused<3>(bufs[from], bufs[to], gx, gy, use_tma, F, ntiles, st); break;
                case 4: rc = launch_fused<4>(bufs[from], bufs[to], gx, gy, use_tma, F, ntiles, st); break;
                default: rc = launch_fused<5>(bufs[from], bufs[to], gx, gy, use_tma, F, ntiles, st); break;
            }
            if (rc) return rc;
        }
        k += len;
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

ABM_EXPORT int64_t abm_diffusion_launches(int64_t steps, int max_fused_substeps) {
    if (steps < 1) return 0;
    const int64_t fuse = max_fused_substeps < 1 ? kMaxFuse : (max_fused_substeps > kMaxFuse ? kMaxFuse : max_fused_substeps);
    int64_t launches = (steps + fuse - 1) / fuse;
    if ((launches ^ steps) & 1) ++launches;
    return launches;
}

ABM_EXPORT int abm_sample_concentrations(const double *locations, int64_t n, const double *field,
                                         const int64_t gsize_host[3], double spat_res, int64_t row_lo,
                                         int64_t row_hi, double *sampled, const int32_t *n_device, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(locations && field && gsize_host && sampled, "null pointer");
    ABM_REQUIRE(spat_res > 0 && gsize_host[0] > 0 && gsize_host[1] > 0 && gsize_host[2] > 0, "bad grid");
    ABM_REQUIRE(row_lo >= 0 && row_hi <= gsize_host[0] && row_lo < row_hi, "bad row window");
    k_sample<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(locations, (int)n, field, gsize_host[0],
                                                              gsize_host[1], gsize_host[2], spat_res, row_lo,
                                                              row_hi, sampled, n_device);
    ABM_CHECK_LAUNCH("k_sample");
    return 0;
}

ABM_EXPORT int abm_deposit_morphogens(const double *locations, const double *amounts, double uniform_amount,
                                      int64_t n, double *field, const int64_t gsize_host[3], double spat_res,
                                      int64_t row_lo, int64_t row_hi, const int32_t *n_device, void *stream) {
    ABM_REQUIRE(n >= 0 && n < 0x7fffffff, "n out of range");
    if (n == 0) return 0;
    ABM_REQUIRE(locations && field && gsize_host, "null pointer");
    ABM_REQUIRE(spat_res > 0 && gsize_host[0] > 0 && gsize_host[1] > 0 && gsize_host[2] > 0, "bad grid");
    ABM_REQUIRE(row_lo >= 0 && row_hi <= gsize_host[0] && row_lo < row_hi, "bad row window");
    k_deposit<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(locations, amounts, uniform_amount, (int)n, field,
                                                               gsize_host[0],
                                                               gsize_host[1], gsize_host[2], spat_res, row_lo,
                                                               row_hi, n_device);
    ABM_CHECK_LAUNCH("k_deposit");
    return 0;
}
