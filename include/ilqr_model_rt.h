/*
 * ilqr_model_rt.h -- runtime support for GENERATED model functions.
 *
 * The reference builds its model callables with Symbolics.jl
 * (/root/reference/src/dynamics.jl:16-34, src/costs.jl:17-44,
 * src/constraints.jl:17-43): in-place functions fn(out, x, u, w) writing
 * dense column-major arrays.  This engine emits the same callables as C
 * (iterativelqr.jl_b200/codegen.py) and compiles the SAME source for the
 * device (nvcc, sm_100a) and for the host (gcc, CPU oracle).
 *
 * Arithmetic contract (DESIGN.md "Arithmetic contract"): generated code is
 * compiled with implicit FMA contraction OFF on both sides (nvcc -fmad=false,
 * gcc -ffp-contract=off) and only uses IEEE-754 correctly rounded operations
 * (+ - * / sqrt, explicit fma) plus the trigonometric functions defined HERE
 * in portable arithmetic, so a model evaluates to identical bits on the GPU and
 * on the host.  That is what lets the parity tests demand identical iteration
 * counts from a chaotic problem (acrobot swing-up).
 */
#ifndef ILQR_MODEL_RT_H
#define ILQR_MODEL_RT_H

#include <math.h>

#if defined(__CUDACC__)
#define ILQR_HD __host__ __device__ __forceinline__
#define ILQR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ILQR_HD static inline __attribute__((always_inline))
#define ILQR_HD_NOINLINE static __attribute__((noinline))
#endif

/* fused multiply-add, correctly rounded on both sides (DFMA / vfmadd) */
ILQR_HD double ilqr_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

/* ---- sin / cos ---------------------------------------------------------
 * Cody-Waite reduction by pi/2 in three FMA steps (valid for |x| < 105615,
 * the same range CUDA's own fast path covers) followed by the classic
 * degree-13 / degree-14 minimax kernels on [-pi/4, pi/4].  Max observed error
 * < 1.5 ulp against mpmath (tests/test_model_rt.py; CUDA documents 2 ulp). Outside
 * the range (or for NaN/Inf) the platform sin/cos is used: correct, but not
 * bit-reproducible across host and device; flagged in DESIGN.md.
 */
#define ILQR_PIO2_HI 1.5707963267948966e+00   /* 0x1.921fb54442d18p+0  */
#define ILQR_PIO2_MID 6.123233995736766e-17   /* 0x1.1a62633145c07p-54 */
#define ILQR_PIO2_LO (-1.4973849048591698e-33) /* -0x1.f1976b7ed8fbcp-110 */
#define ILQR_TWO_OVER_PI 0.6366197723675814
#define ILQR_RND_MAGIC 6755399441055744.0     /* 1.5 * 2^52 */
#define ILQR_TRIG_MAX 105615.0

ILQR_HD double ilqr_ksin(double r) {
    const double z = r * r;
    double p = 1.58969099521155010221e-10;
    p = ilqr_fma(p, z, -2.50507602534068634195e-08);
    p = ilqr_fma(p, z, 2.75573137070700676789e-06);
    p = ilqr_fma(p, z, -1.98412698298579493134e-04);
    p = ilqr_fma(p, z, 8.33333333332248946124e-03);
    p = ilqr_fma(p, z, -1.66666666666666324348e-01);
    return ilqr_fma(z * r, p, r);
}

ILQR_HD double ilqr_kcos(double r) {
    const double z = r * r;
    double p = -1.13596475577881948265e-11;
    p = ilqr_fma(p, z, 2.08757232129817482790e-09);
    p = ilqr_fma(p, z, -2.75573143513906633035e-07);
    p = ilqr_fma(p, z, 2.48015872894767294178e-05);
    p = ilqr_fma(p, z, -1.38888888888741095749e-03);
    p = ilqr_fma(p, z, 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + z * (z * p));
}

/* reduce x to r in [-pi/4, pi/4], return quadrant (0..3) */
ILQR_HD int ilqr_rem_pio2(double x, double* r_out) {
    const double t = ilqr_fma(x, ILQR_TWO_OVER_PI, ILQR_RND_MAGIC);
    const double n = t - ILQR_RND_MAGIC;
    double r = ilqr_fma(-n, ILQR_PIO2_HI, x);
    r = ilqr_fma(-n, ILQR_PIO2_MID, r);
    r = ilqr_fma(-n, ILQR_PIO2_LO, r);
    *r_out = r;
    return ((int)n) & 3;
}

ILQR_HD void ilqr_sincos(double x, double* s_out, double* c_out) {
    if (!(fabs(x) < ILQR_TRIG_MAX)) { /* also catches NaN / Inf */
        *s_out = sin(x);
        *c_out = cos(x);
        return;
    }
    double r;
    const int q = ilqr_rem_pio2(x, &r);
    const double s = ilqr_ksin(r);
    const double c = ilqr_kcos(r);
    const double ss = (q & 1) ? c : s;
    const double cc = (q & 1) ? s : c;
    *s_out = (q & 2) ? -ss : ss;
    *c_out = ((q + 1) & 2) ? -cc : cc;
}

ILQR_HD double ilqr_sin(double x) {
    double s, c;
    ilqr_sincos(x, &s, &c);
    return s;
}

ILQR_HD double ilqr_cos(double x) {
    double s, c;
    ilqr_sincos(x, &s, &c);
    return c;
}

#endif /* ILQR_MODEL_RT_H */
