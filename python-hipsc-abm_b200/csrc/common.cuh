// common.cuh -- shared device/host helpers for libabm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "abm_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libabm_b200 is written for sm_100a (B200) only"
#endif

#define ABM_EXPORT extern "C" __attribute__((visibility("default")))

namespace abm {

constexpr int kSMsB200 = 148;

void set_error(const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);   // 0 or -(1000+e), records the text

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

#define ABM_CHECK_LAUNCH(what)                                   \
    do {                                                         \
        int _rc = abm::check_cuda(cudaGetLastError(), what);     \
        if (_rc) return _rc;                                     \
    } while (0)

#define ABM_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            abm::set_error("%s: %s", __func__, msg);             \
            return -1;                                           \
        }                                                        \
    } while (0)

// ---- bin geometry --------------------------------------------------------------------------------
struct BinGeom {
    double distance;
    double inv_distance;
    double origin[3]; // lattice origin (0 for a whole colony; the slab's lower corner for a slab-decomposed one)
    int dims[3];      // ceil(extent/distance)+3 per axis (reference simulation.py:202)
};

// Exact floor(a/b) for b > 0: numpy's float floor_divide (npy_divmod) returns the mathematically
// exact floor, which is what bin_locations uses (reference simulation.py:210).  q = floor(a/b) can be
// one too large when a/b rounds up to an integer; the FMA remainder a - q*b is exact and fixes it.
__device__ __forceinline__ double exact_floor_div(double a, double b) {
    double q = floor(a / b);
    double r = fma(-q, b, a);
    if (r < 0.0) q -= 1.0;
    else if (r >= b) q += 1.0;
    return q;
}

// Same result without the division: the estimate floor(a * (1/b)) is off by at most one for the magnitudes that
// occur (|a/b| < 2^31), and the exact FMA remainder corrects it in either direction.
__device__ __forceinline__ double exact_floor_mul(double a, double b, double inv_b) {
    double q = floor(a * inv_b);
    double r = fma(-q, b, a);
    if (r < 0.0) q -= 1.0;
    else if (r >= b) q += 1.0;
    return q;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// bin coordinate of one axis, clamped into the allocated lattice
__device__ __forceinline__ int bin_coord(double x, double distance, int dim) {
    return clampi((int)exact_floor_div(x, distance) + 1, 0, dim - 1);
}
__device__ __forceinline__ int bin_coord_fast(double x, double distance, double inv_distance, int dim) {
    // |x/distance| beyond int range cannot occur for finite colonies; clamp first to keep the cast defined
    double q = exact_floor_mul(x, distance, inv_distance);
    q = fmin(fmax(q, -2.0), 2147483000.0);
    return clampi((int)q + 1, 0, dim - 1);
}

__device__ __forceinline__ int lin_bin(int bx, int by, int bz, const int dims[3]) {
    return (bz * dims[1] + by) * dims[0] + bx;
}

// ---- the distance predicate ------------------------------------------------------------------------
// Sum of squares without FMA contraction, in the oracle's order ((dx*dx + dy*dy) + dz*dz).
__device__ __forceinline__ double sumsq_nofma(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Thresholds so that "sqrt(s) <= d" (the reference predicate, backend.py:143) is decided without
// a square root except in a 2^-50-wide band around d*d.
struct Radius {
    double d, lo, hi;
    __host__ __device__ explicit Radius(double dist) : d(dist) {
        double d2 = dist * dist;
        lo = d2 * (1.0 - 0x1p-50);
        hi = d2 * (1.0 + 0x1p-50);
    }
    __device__ __forceinline__ bool within(double s) const {
        if (s <= lo) return true;
        if (s > hi) return false;
        return tie(s, d);
    }
    // out of line on purpose: inlined, the compiler if-converts it and every candidate pays for the square root
    static __device__ __noinline__ bool tie(double s, double dist) { return __dsqrt_rn(s) <= dist; }
};

// ---- programmatic dependent launch -------------------------------------------------------------------
// The kernels of the mechanics loop form a chain (scan -> placement -> fused substep -> scan ...), each consuming
// what its predecessor wrote.  Launched with the programmatic-serialization attribute, a kernel's CTAs may become
// resident while the predecessor's last wave drains: they run their prologue, then block in pdl_wait() until the
// predecessor has completed and its writes are visible.  pdl_launch_dependents() (first statement of every chain
// kernel) is what allows the successor to start being scheduled once every CTA of this grid has started.
// Rule: no global-memory access before pdl_wait().
#ifndef ABM_PDL
#define ABM_PDL 1
#endif
__device__ __forceinline__ void pdl_launch_dependents() {
#if ABM_PDL
    asm volatile("griddepcontrol.launch_dependents;");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if ABM_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

// chained: launch with the programmatic-serialization attribute (the caller's ABM_LOOP_CHAINED / ABM_SORT_CHAINED
// flag; there is no process-wide switch)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, bool chained,
                                  Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = (ABM_PDL && chained) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ---- per-device caches ------------------------------------------------------------------------------
// Function attributes (opt-in dynamic shared memory) and the SM count belong to a DEVICE, not to the process: a
// second engine on cuda:1 must not inherit what was set / read on cuda:0.  Slot of the current device, or -1.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}

// ---- device-side cell count ----------------------------------------------------------------------------
// Every per-cell kernel takes its cell count as a launch parameter n (grid size) and, optionally, as a pointer to a
// device int: the count is then min(n, *n_dev), n being only the capacity the launch was sized for.  A pipeline
// whose cell count is decided on the device (cells migrating between slabs) can so enqueue -- or replay as a CUDA
// graph -- a whole step without ever bringing the count to the host.
__device__ __forceinline__ int device_count(int n, const int *__restrict__ n_dev) {
    if (n_dev) {
        int m = *n_dev;
        n = m < n ? (m < 0 ? 0 : m) : n;
    }
    return n;
}

// Largest squared displacement of a cell in one substep, kept as the bit pattern of a non-negative float (ordered like
// an int) in max_len[1]: the slab decomposition's halo depth assumes a bound on it (slab.required_halo) and checks
// the assumption after the fact.  window (device, may be null = every cell): only cells with x < window[0] or
// x > window[1] are looked at -- a slab's ghost zones, the only place where a wrong cell can exist while the halo
// holds.  One REDUX and one atomic per warp.
__device__ __forceinline__ void record_move(int *__restrict__ max_len, const double *__restrict__ window, double x,
                                            float d2) {
    if (window && x >= __ldg(window) && x <= __ldg(window + 1)) d2 = 0.0f;
    const unsigned active = __activemask();
    const int bits = __reduce_max_sync(active, __float_as_int(d2 >= 0.0f ? d2 : 0.0f));
    if (bits && (int)(threadIdx.x & 31) == __ffs(active) - 1) atomicMax(max_len + 1, bits);
}

// ---- warp helpers ----------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ int warp_inclusive_scan(int v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= o) v += t;
    }
    return v;
}

// ascending insertion sort of a short int row
__device__ __forceinline__ void insertion_sort(int *a, int n) {
    for (int i = 1; i < n; ++i) {
        int v = a[i];
        int j = i - 1;
        while (j >= 0 && a[j] > v) {
            a[j + 1] = a[j];
            --j;
        }
        a[j + 1] = v;
    }
}

// ---- JKR contact model (reference cell_backend.py:79-98) ------------------------------------------
struct JkrConst {
    double k23;        // ((pi*adhesion)/e_hat)^(2/3)
    double pi_adh;     // pi*adhesion is applied as (f*pi)*adhesion in the reference; kept separate below
    double adhesion;
    // the radius-dependent factors for a colony whose radii all equal uni_r (0 = radii differ)
    double uni_r, uni_inv_scale, uni_force_scale;
};

__host__ JkrConst make_jkr_const(double poisson, double youngs, double adhesion, double uniform_radius);

// Returns true if the bond survives; fx,fy,fz = force on the cell at p from the cell at q.
// v = p - q, s = |v|^2 (no-FMA sum of squares), rp/rq radii in um.
__device__ __forceinline__ bool jkr_bond(double vx, double vy, double vz, double s, double rp, double rq,
                                         const JkrConst &c, double &fx, double &fy, double &fz) {
    const double kPi = 3.141592653589793;
    double mag = __dsqrt_rn(s);
    double overlap = (rp + rq - mag) / 1e6;
    double r_hat = 1.0 / (1e6 * ((1.0 / rp) + (1.0 / rq)));
    double overlap_ = c.k23 * cbrt(r_hat);
    double d = overlap / overlap_;
    if (d > -0.360562) {
        double f = (-0.0204 * (d * d * d)) + (0.4942 * (d * d)) + (1.0801 * d) - 1.324;
        double jkr = f * kPi * c.adhesion * r_hat;
        if (mag != 0.0) {
            fx = jkr * (vx / mag);
            fy = jkr * (vy / mag);
            fz = jkr * (vz / mag);
        } else {
            fx = fy = fz = 0.0;
        }
        return true;
    }
    return false;
}

// Stokes integration of one axis (reference cell_backend.py:155-170), operation order preserved.
__device__ __forceinline__ double stokes_move(double x, double f_mot, double f_jkr, double friction, double dt,
                                              double size_k) {
    double velocity = __ddiv_rn(__dadd_rn(f_mot, f_jkr), friction);
    double moved = __dadd_rn(x, __dmul_rn(dt, __dmul_rn(velocity, 1e6)));
    if (moved > size_k) return size_k;
    if (moved < 0.0) return 0.0;
    return moved;
}

__device__ __forceinline__ double stokes_friction(double radius, double stokes) {
    const double kPi = 3.141592653589793;
    return __dmul_rn(__dmul_rn(__dmul_rn(6.0, kPi), stokes), __ddiv_rn(radius, 1e6));
}

// block-wide exclusive scan utilities live in scan.cuh
}  // namespace abm
