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
 * < 1.5 ulp against mpmath (tests/test_model_rt.py; CUDA documents 2 ulp). Larger
 * arguments take an integer Payne-Hanek reduction (below), NaN/Inf give NaN: every input
 * evaluates to the same bits on host and device.
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

/* ---- huge arguments: Payne-Hanek reduction in integer arithmetic -------------------
 * x * 2/pi is formed exactly enough from a 1152-bit table of 2/pi: the 64-bit mantissa of
 * |x| times the four table words that can influence (x * 2/pi) mod 4 down to 2^-126.
 * Integer operations only until the final double-double product, hence bit-reproducible
 * on host and device.  Rare path (a rollout that has already diverged), kept out of line. */
ILQR_HD unsigned long long ilqr_d2bits(double x) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(x);
#else
    union { double d; unsigned long long u; } v; v.d = x; return v.u;
#endif
}
ILQR_HD double ilqr_bits2d(unsigned long long u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    union { double d; unsigned long long u; } v; v.u = u; return v.d;
#endif
}
ILQR_HD int ilqr_clz64(unsigned long long v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
ILQR_HD double ilqr_pow2(int e) { return ilqr_bits2d((unsigned long long)(1023 + e) << 52); } /* normal range only */

ILQR_HD_NOINLINE int ilqr_rem_pio2_large(double ax, double* r_out) { /* ax finite, >= ILQR_TRIG_MAX */
    const unsigned long long T[18] = {
        0xa2f9836e4e441529ULL, 0xfc2757d1f534ddc0ULL, 0xdb6295993c439041ULL, 0xfe5163abdebbc561ULL,
        0xb7246e3a424dd2e0ULL, 0x06492eea09d1921cULL, 0xfe1deb1cb129a73eULL, 0xe88235f52ebb4484ULL,
        0xe99c7026b45f7e41ULL, 0x3991d639835339f4ULL, 0x9c845f8bbdf9283bULL, 0x1ff897ffde05980fULL,
        0xef2f118b5a0a6d1fULL, 0x6d367ecf27cb09b7ULL, 0x4f463f669e5fea2dULL, 0x7527bac7ebe5f17bULL,
        0x3d0739f78a5292eaULL, 0x6bfb5fb11f8d5d08ULL};
    const unsigned long long bits = ilqr_d2bits(ax);
    const int ex = (int)((bits >> 52) & 0x7ff) - 1023;
    const unsigned long long ia = ((bits & 0x000fffffffffffffULL) | 0x0010000000000000ULL) << 11;
    int j0 = (ex - 129) >= 0 ? (ex - 129) / 64 + 1 : 0;
    const int s0 = ex - 127 - 64 * j0;
    unsigned long long w[7] = {0, 0, 0, 0, 0, 0, 0}; /* w[0] least significant; 320-bit product + 2 guard words */
    for (int k = 3; k >= 0; --k) {
        const int idx = 3 - k;
        const unsigned __int128 p = (unsigned __int128)ia * T[j0 + k];
        unsigned __int128 acc = (unsigned __int128)w[idx] + (unsigned long long)p;
        w[idx] = (unsigned long long)acc;
        acc = (unsigned __int128)w[idx + 1] + (unsigned long long)(p >> 64) + (unsigned long long)(acc >> 64);
        w[idx + 1] = (unsigned long long)acc;
        unsigned long long carry = (unsigned long long)(acc >> 64);
        for (int q = idx + 2; q < 5 && carry; ++q) {
            const unsigned __int128 a2 = (unsigned __int128)w[q] + carry;
            w[q] = (unsigned long long)a2;
            carry = (unsigned long long)(a2 >> 64);
        }
    }
    /* weight-1 bit sits at p0 = 192 - s0 (190..303); shift left so the two quadrant bits top w[4] */
    const int sh = 318 - (192 - s0); /* 15..128 */
    const int ws = sh >> 6, bs = sh & 63;
    unsigned long long v[3]; /* the top three words after the shift */
    for (int i = 0; i < 3; ++i) {
        const int src = 4 - i - ws; /* word that lands in position 4 - i */
        const unsigned long long hi = (src >= 0) ? w[src] : 0ULL;
        const unsigned long long lo = (src - 1 >= 0) ? w[src - 1] : 0ULL;
        v[i] = bs ? ((hi << bs) | (lo >> (64 - bs))) : hi;
    }
    int n = (int)(v[0] >> 62);
    unsigned long long fh = (v[0] << 2) | (v[1] >> 62); /* fraction, top 64 bits */
    unsigned long long fl = (v[1] << 2) | (v[2] >> 62); /* next 64 bits */
    double sign = 1.0;
    if (fh >> 63) { /* fraction >= 1/2: round the quadrant up, fraction becomes negative */
        n = (n + 1) & 3;
        fl = ~fl + 1ULL;
        fh = ~fh + (fl == 0ULL ? 1ULL : 0ULL);
        sign = -1.0;
    }
    if (fh == 0ULL && fl == 0ULL) { *r_out = 0.0; return n; }
    int lz;
    unsigned long long hi64, lo64;
    if (fh != 0ULL) {
        lz = ilqr_clz64(fh);
        hi64 = lz ? ((fh << lz) | (fl >> (64 - lz))) : fh;
        lo64 = lz ? (fl << lz) : fl;
    } else {
        const int l2 = ilqr_clz64(fl);
        lz = 64 + l2;
        hi64 = fl << l2;
        lo64 = 0ULL;
    }
    const double a = (double)(hi64 >> 11);                                  /* 53 bits, exact */
    const double b = (double)(((hi64 & 0x7ffULL) << 42) | (lo64 >> 22));    /* next 53 bits, exact */
    const double f1 = a * ilqr_pow2(-53 - lz);
    const double f2 = b * ilqr_pow2(-106 - lz);
    const double p = f1 * ILQR_PIO2_HI;
    const double e = ilqr_fma(f1, ILQR_PIO2_HI, -p);
    const double r = p + (e + ilqr_fma(f1, ILQR_PIO2_MID, f2 * ILQR_PIO2_HI));
    *r_out = sign * r;
    return n;
}

/* quadrant fix-up shared by every path: (q, r) -> sin, cos */
ILQR_HD void ilqr_sincos_finish(int q, double r, double* s_out, double* c_out) {
    const double s = ilqr_ksin(r);
    const double c = ilqr_kcos(r);
    const double ss = (q & 1) ? c : s;
    const double cc = (q & 1) ? s : c;
    *s_out = (q & 2) ? -ss : ss;
    *c_out = ((q + 1) & 2) ? -cc : cc;
}

/* |x| < ILQR_TRIG_MAX only: the branch-free body of ilqr_sincos.  Generated model code calls it for a whole
 * group of independent arguments behind ONE range test (codegen.py: _emit_trig_group), so that the compiler can
 * interleave the dependent chains of the group; each argument sees exactly the operations ilqr_sincos applies. */
ILQR_HD void ilqr_sincos_small(double x, double* s_out, double* c_out) {
    double r;
    const int q = ilqr_rem_pio2(x, &r);
    ilqr_sincos_finish(q, r, s_out, c_out);
}

ILQR_HD int ilqr_trig_is_small(double x) { return fabs(x) < ILQR_TRIG_MAX; }

ILQR_HD void ilqr_sincos(double x, double* s_out, double* c_out) {
    double r;
    int q;
    const double ax = fabs(x);
    if (ax < ILQR_TRIG_MAX) {
        q = ilqr_rem_pio2(x, &r);
    } else if (ax <= 1.7976931348623157e308) {
        q = ilqr_rem_pio2_large(ax, &r);
        if (x < 0.0) { r = -r; q = (4 - q) & 3; }
    } else { /* NaN / Inf */
        *s_out = x - x;
        *c_out = x - x;
        return;
    }
    ilqr_sincos_finish(q, r, s_out, c_out);
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
