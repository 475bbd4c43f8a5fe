/*
 * abm_rng.h -- the random-number contract of the DEVICE-RESIDENT rule kernels.
 *
 * The reference's rule methods (cell_methods.py:11-340) draw from Python's `random` (one global Mersenne Twister,
 * consumed in cell-index order): inherently serial.  The device rules replace it by a COUNTER-BASED generator: every
 * draw is a pure function of (seed, simulation step, stream = which rule, the cell's persistent id, draw index), so a
 * cell's randomness does not depend on how many cells exist, on their order in memory, or on which GPU owns them.
 * This changes the random stream, hence the restated oracle (oracle/abm_oracle.c orc_rules_*) uses the same contract;
 * the seeded-parity path against the ORIGINAL stream stays on the host (oracle/model.py, tests/test_model_states.py).
 *
 * Everything here is plain C99 usable from gcc and nvcc (host and device), and uses only IEEE-exact operations
 * (+ - * / sqrt, integer ops) in a fixed order without FMA contraction, so the CPU restatement and the CUDA kernels
 * produce BIT-IDENTICAL draws: log and sin/cos are evaluated by the fixed polynomials below, not by libm.
 */
#ifndef ABM_RNG_H
#define ABM_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define ABM_RNG_FN static __host__ __device__ __forceinline__
#else
#include <math.h>
#define ABM_RNG_FN static inline
#endif

/* un-contracted arithmetic: nvcc fuses a*b+c into an FMA in device code unless told otherwise */
#if defined(__CUDA_ARCH__)
#define ABM_MUL(a, b) __dmul_rn((a), (b))
#define ABM_ADD(a, b) __dadd_rn((a), (b))
#define ABM_SUB(a, b) __dsub_rn((a), (b))
#define ABM_DIV(a, b) __ddiv_rn((a), (b))
#define ABM_SQRT(a) __dsqrt_rn((a))
#else
#define ABM_MUL(a, b) ((a) * (b))
#define ABM_ADD(a, b) ((a) + (b))
#define ABM_SUB(a, b) ((a) - (b))
#define ABM_DIV(a, b) ((a) / (b))
#define ABM_SQRT(a) sqrt((a))
#endif

/* streams: one per use of randomness in the rules */
enum {
    ABM_STREAM_DIV_COUNTER = 1,   /* cell_division: div_counters += randint(0,1)        (cell_methods.py:65)  */
    ABM_STREAM_DIV_VECTOR = 2,    /* cell_division: random_vector() of the displacement  (cell_methods.py:103) */
    ABM_STREAM_PATHWAY = 3,       /* cell_pathway: one gauss(0,1) per neighbour          (cell_methods.py:192) */
    ABM_STREAM_DIFF_COUNTER = 4,  /* cell_differentiate: diff_counters += randint(0,1)   (cell_methods.py:239) */
    ABM_STREAM_MOTILITY = 5       /* cell_motility: random_vector()                      (cell_methods.py:275..) */
};

ABM_RNG_FN uint64_t abm_mix64(uint64_t x) {           /* splitmix64 finalizer */
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

/* the k-th 64-bit draw of (seed, step, stream, cell id) */
ABM_RNG_FN uint64_t abm_draw(uint64_t seed, uint64_t step, uint64_t stream, uint64_t id, uint64_t k) {
    uint64_t x = abm_mix64(seed + 0x9e3779b97f4a7c15ull * (step + 1));
    x = abm_mix64(x ^ (stream * 0xd1342543de82ef95ull + id));
    return abm_mix64(x + 0x9e3779b97f4a7c15ull * (k + 1));
}

ABM_RNG_FN double abm_uniform(uint64_t bits) { return (double)(bits >> 11) * 0x1p-53; }   /* [0,1), 53 bits */

ABM_RNG_FN int abm_coin(uint64_t bits) { return (int)(bits >> 63); }                      /* randint(0,1) */

/* natural logarithm of a positive normal x by argument reduction to m in [sqrt(1/2), sqrt(2)) and the series
 * ln m = 2 t (1 + t^2/3 + t^4/5 + ...), t = (m-1)/(m+1), 13 terms (|t| <= 0.1716: truncation < 2e-20); error
 * <= 2 ulp.  Fixed operation order, no libm. */
ABM_RNG_FN double abm_log(double x) {
    union { double d; uint64_t u; } v;
    v.d = x;
    int e = (int)((v.u >> 52) & 0x7ff) - 1022;                       /* x = m 2^e, m in [0.5,1) */
    v.u = (v.u & 0x000fffffffffffffull) | 0x3fe0000000000000ull;
    double m = v.d;
    if (m < 0.70710678118654752440) {
        m = ABM_MUL(m, 2.0);
        e -= 1;
    }
    const double t = ABM_DIV(ABM_SUB(m, 1.0), ABM_ADD(m, 1.0));
    const double t2 = ABM_MUL(t, t);
    double p = 1.0 / 25.0;
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 23.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 21.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 19.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 17.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 15.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 13.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 11.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 9.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 7.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 5.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0 / 3.0);
    p = ABM_ADD(ABM_MUL(p, t2), 1.0);
    return ABM_ADD(ABM_MUL(ABM_MUL(2.0, t), p), ABM_MUL((double)e, 0.69314718055994530942));
}

/* sin and cos of 2 pi u, u in [0,1): quadrant reduction (exact) + Taylor polynomials on [0, pi/2) */
ABM_RNG_FN void abm_sincos_2pi(double u, double *sn, double *cs) {
    const double a = ABM_MUL(u, 4.0);
    int q = (int)a;                                                  /* 0..3 */
    const double f = ABM_SUB(a, (double)q);
    const double x = ABM_MUL(f, 1.57079632679489661923);
    const double x2 = ABM_MUL(x, x);
    /* sin x = x (1 - x2/3! + x2^2/5! - ...), cos x = 1 - x2/2! + x2^2/4! - ...  up to x^23 / x^24 */
    double s = -1.0 / 25852016738884976640000.0;                     /* -1/23! */
    s = ABM_ADD(ABM_MUL(s, x2), 1.0 / 51090942171709440000.0);       /* 1/21! */
    s = ABM_ADD(ABM_MUL(s, x2), -1.0 / 121645100408832000.0);        /* 1/19! */
    s = ABM_ADD(ABM_MUL(s, x2), 1.0 / 355687428096000.0);            /* 1/17! */
    s = ABM_ADD(ABM_MUL(s, x2), -1.0 / 1307674368000.0);             /* 1/15! */
    s = ABM_ADD(ABM_MUL(s, x2), 1.0 / 6227020800.0);                 /* 1/13! */
    s = ABM_ADD(ABM_MUL(s, x2), -1.0 / 39916800.0);                  /* 1/11! */
    s = ABM_ADD(ABM_MUL(s, x2), 1.0 / 362880.0);                     /* 1/9! */
    s = ABM_ADD(ABM_MUL(s, x2), -1.0 / 5040.0);                      /* 1/7! */
    s = ABM_ADD(ABM_MUL(s, x2), 1.0 / 120.0);                        /* 1/5! */
    s = ABM_ADD(ABM_MUL(s, x2), -1.0 / 6.0);                         /* 1/3! */
    s = ABM_ADD(ABM_MUL(s, x2), 1.0);
    s = ABM_MUL(s, x);
    double c = 1.0 / 620448401733239439360000.0;                     /* 1/24! */
    c = ABM_ADD(ABM_MUL(c, x2), -1.0 / 1124000727777607680000.0);    /* 1/22! */
    c = ABM_ADD(ABM_MUL(c, x2), 1.0 / 2432902008176640000.0);        /* 1/20! */
    c = ABM_ADD(ABM_MUL(c, x2), -1.0 / 6402373705728000.0);          /* 1/18! */
    c = ABM_ADD(ABM_MUL(c, x2), 1.0 / 20922789888000.0);             /* 1/16! */
    c = ABM_ADD(ABM_MUL(c, x2), -1.0 / 87178291200.0);               /* 1/14! */
    c = ABM_ADD(ABM_MUL(c, x2), 1.0 / 479001600.0);                  /* 1/12! */
    c = ABM_ADD(ABM_MUL(c, x2), -1.0 / 3628800.0);                   /* 1/10! */
    c = ABM_ADD(ABM_MUL(c, x2), 1.0 / 40320.0);                      /* 1/8! */
    c = ABM_ADD(ABM_MUL(c, x2), -1.0 / 720.0);                       /* 1/6! */
    c = ABM_ADD(ABM_MUL(c, x2), 1.0 / 24.0);                         /* 1/4! */
    c = ABM_ADD(ABM_MUL(c, x2), -0.5);
    c = ABM_ADD(ABM_MUL(c, x2), 1.0);
    q &= 3;
    if (q == 0) { *sn = s; *cs = c; }
    else if (q == 1) { *sn = c; *cs = -s; }
    else if (q == 2) { *sn = -s; *cs = -c; }
    else { *sn = -c; *cs = s; }
}

/* gauss(0,1): Marsaglia's polar method on the draws k0, k0+1, ... of the stream; *used = draws consumed (2 per
 * attempt, 1.27 attempts on average).  Only the first variate of each accepted pair is used. */
ABM_RNG_FN double abm_gauss(uint64_t seed, uint64_t step, uint64_t stream, uint64_t id, uint64_t k0, uint64_t *used) {
    uint64_t k = k0;
    for (;;) {
        const double v1 = ABM_SUB(ABM_MUL(2.0, abm_uniform(abm_draw(seed, step, stream, id, k))), 1.0);
        const double v2 = ABM_SUB(ABM_MUL(2.0, abm_uniform(abm_draw(seed, step, stream, id, k + 1))), 1.0);
        k += 2;
        const double s = ABM_ADD(ABM_MUL(v1, v1), ABM_MUL(v2, v2));
        if (s < 1.0 && s > 0.0) {
            *used = k - k0;
            return ABM_MUL(v1, ABM_SQRT(ABM_DIV(ABM_MUL(-2.0, abm_log(s)), s)));
        }
    }
}

/* Simulation.random_vector() (reference simulation.py:473-487): a direction from one (2-D) or two (3-D) uniform
 * angles -- the reference's formula, including its non-uniform 3-D distribution. */
ABM_RNG_FN void abm_random_vector(uint64_t seed, uint64_t step, uint64_t stream, uint64_t id, int three_d,
                                  double out[3]) {
    double st, ct;
    abm_sincos_2pi(abm_uniform(abm_draw(seed, step, stream, id, 0)), &st, &ct);
    if (!three_d) {
        out[0] = ct;
        out[1] = st;
        out[2] = 0.0;
        return;
    }
    double sp, cp;
    abm_sincos_2pi(abm_uniform(abm_draw(seed, step, stream, id, 1)), &sp, &cp);
    out[0] = ABM_MUL(cp, ct);
    out[1] = ABM_MUL(cp, st);
    out[2] = sp;
}

#endif /* ABM_RNG_H */
