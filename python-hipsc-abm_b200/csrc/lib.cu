// lib.cu -- library-level entry points: version, error text, device info, small host helpers.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace abm {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return -(1000 + (int)e);
}

JkrConst make_jkr_const(double poisson, double youngs, double adhesion, double uniform_radius) {
    // reference cell_backend.py:84,88 -- e_hat and the radius-independent factor of overlap_
    double e_hat = 1.0 / (((1 - poisson * poisson) / youngs) + ((1 - poisson * poisson) / youngs));
    JkrConst c;
    c.k23 = pow((M_PI * adhesion) / e_hat, 2.0 / 3.0);
    c.pi_adh = M_PI * adhesion;
    c.adhesion = adhesion;
    c.uni_r = c.uni_inv_scale = c.uni_force_scale = 0.0;
    if (uniform_radius > 0) {
        // reference cell_backend.py:85,88 for r1 == r2 == uniform_radius
        double r_hat = 1.0 / (1e6 * ((1.0 / uniform_radius) + (1.0 / uniform_radius)));
        double overlap_ = c.k23 * cbrt(r_hat);
        c.uni_r = uniform_radius;
        c.uni_inv_scale = 1.0 / (1e6 * overlap_);
        c.uni_force_scale = M_PI * adhesion * r_hat;
    }
    return c;
}

}  // namespace abm

ABM_EXPORT int abm_abi_version(void) { return ABM_B200_ABI_VERSION; }

ABM_EXPORT int abm_compiled_arch(void) { return 100; }

ABM_EXPORT const char *abm_last_error(void) { return abm::g_error; }

ABM_EXPORT int abm_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    int rc = abm::check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc) return rc;
    cudaDeviceProp p;
    rc = abm::check_cuda(cudaGetDeviceProperties(&p, dev), "cudaGetDeviceProperties");
    if (rc) return rc;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}

ABM_EXPORT int abm_read_i32_sync(const int32_t *device_value, int32_t *value_host, void *stream) {
    ABM_REQUIRE(device_value && value_host, "null pointer");
    int rc = abm::check_cuda(cudaMemcpyAsync(value_host, device_value, sizeof(int32_t), cudaMemcpyDeviceToHost,
                                             abm::as_stream(stream)),
                             "cudaMemcpyAsync(read_i32)");
    if (rc) return rc;
    return abm::check_cuda(cudaStreamSynchronize(abm::as_stream(stream)), "cudaStreamSynchronize");
}

ABM_EXPORT double abm_jkr_reach(double distance, double max_radius, double poisson, double youngs,
                                double adhesion_const) {
    // Largest centre distance at which a bond can survive: d > -0.360562 <=> mag < r1+r2+0.360562e6*overlap_
    // with overlap_ = k23*cbrt(r_hat) and r_hat <= max_radius/2e6 (reference cell_backend.py:81-94).
    abm::JkrConst c = abm::make_jkr_const(poisson, youngs, adhesion_const, 0.0);
    double brk = 2.0 * max_radius + 0.360562e6 * c.k23 * cbrt(max_radius / 2e6);
    double reach = distance > brk ? distance : brk;
    return reach * (1.0 + 1e-9);
}
