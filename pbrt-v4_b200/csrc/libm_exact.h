/* libm_exact.h -- bit-exact restatement of the float logf / sinf / cosf the reference CPU path calls (glibc 2.39 on x86-64 with
 * FMA: the __logf_fma / __sinf_fma / __cosf_fma ifunc variants, i.e. ARM optimized-routines' algorithms with the a*b+c terms
 * contracted; sysdeps/ieee754/flt-32/e_logf.c, s_sinf.c, s_cosf.c, sincosf.h; constants as found in libm.so.6's rodata).
 *
 * Why: CUDA's float functions differ from glibc's in 15-33 % of arguments and "evaluate in double, round once" still differs
 * wherever glibc itself is not correctly rounded (0.06-1.3 % for these three).  A 1-ulp difference is invisible in a smooth
 * integrand but decorrelates a path wherever a float is hashed into an RNG seed (LayeredBxDF, transmittance tracking), so exact
 * agreement needs the host's algorithm.  Everything here is IEEE double +, *, fma and float<->double conversion: the device
 * (compiled with -fmad=false) produces the same bits as the host.  Verified exhaustively against the host libm by
 * tests/tools/libm_exact_check.c (all positive floats for logf, all |x| < 120 for sinf / cosf).
 *
 * C and C++ (host + device).  Arguments outside the verified domain fall back to the caller's double-precision evaluation. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define BPT_LM_HD __host__ __device__ __forceinline__
#else
#define BPT_LM_HD static inline
#endif

BPT_LM_HD uint32_t bpt_lm_asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
BPT_LM_HD float bpt_lm_asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ------------------------------------------------------------------ logf (LOGF_TABLE_BITS = 4, LOGF_POLY_ORDER = 4) */
BPT_LM_HD double bpt_lm_logf_invc(int i) {
    switch (i) {
    case 0: return 0x1.661ec79f8f3bep+0; case 1: return 0x1.571ed4aaf883dp+0; case 2: return 0x1.49539f0f010b0p+0; case 3: return 0x1.3c995b0b80385p+0;
    case 4: return 0x1.30d190c8864a5p+0; case 5: return 0x1.25e227b0b8ea0p+0; case 6: return 0x1.1bb4a4a1a343fp+0; case 7: return 0x1.12358f08ae5bap+0;
    case 8: return 0x1.0953f419900a7p+0; case 9: return 0x1.0000000000000p+0; case 10: return 0x1.e608cfd9a47acp-1; case 11: return 0x1.ca4b31f026aa0p-1;
    case 12: return 0x1.b2036576afce6p-1; case 13: return 0x1.9c2d163a1aa2dp-1; case 14: return 0x1.886e6037841edp-1; default: return 0x1.767dcf5534862p-1;
    }
}
BPT_LM_HD double bpt_lm_logf_logc(int i) {
    switch (i) {
    case 0: return -0x1.57bf7808caadep-2; case 1: return -0x1.2bef0a7c06ddbp-2; case 2: return -0x1.01eae7f513a67p-2; case 3: return -0x1.b31d8a68224e9p-3;
    case 4: return -0x1.6574f0ac07758p-3; case 5: return -0x1.1aa2bc79c8100p-3; case 6: return -0x1.a4e76ce8c0e5ep-4; case 7: return -0x1.1973c5a611cccp-4;
    case 8: return -0x1.252f438e10c1ep-5; case 9: return 0x0.0p+0; case 10: return 0x1.aa5aa5df25984p-5; case 11: return 0x1.c5e53aa362eb4p-4;
    case 12: return 0x1.526e57720db08p-3; case 13: return 0x1.bc2860d224770p-3; case 14: return 0x1.1058bc8a07ee1p-2; default: return 0x1.4043057b6ee09p-2;
    }
}
BPT_LM_HD float bpt_exact_logf(float x) {
    uint32_t ix = bpt_lm_asuint(x);
    if (ix == 0x3f800000u) return 0;
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        /* x < 0x1p-126 or inf or nan */
        if (ix * 2 == 0) return -INFINITY;
        if (ix == 0x7f800000u) return x;
        if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return NAN;
        ix = bpt_lm_asuint(x * 0x1p23f);   /* subnormal: normalise */
        ix -= 23u << 23;
    }
    uint32_t tmp = ix - 0x3f330000u;
    int i = (int)((tmp >> (23 - 4)) % 16);
    int k = (int32_t)tmp >> 23;
    uint32_t iz = ix - (tmp & 0xff800000u);
    double invc = bpt_lm_logf_invc(i), logc = bpt_lm_logf_logc(i);
    double z = (double)bpt_lm_asfloat(iz);
    const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
    double r = fma(z, invc, -1.0);
    double y0 = fma((double)k, Ln2, logc);
    double r2 = r * r;
    double y = fma(A1, r, A2);
    y = fma(A0, r2, y);
    y = fma(y, r2, y0 + r);
    return (float)y;
}

/* ------------------------------------------------------------------ sinf / cosf (|x| < 120: the fast reduction path) */
BPT_LM_HD float bpt_lm_sinf_poly(double x, double x2, int neg, int n) {
    /* __sincosf_table[neg]: c0..c4, s1..s3 (table 1 negates the cosine coefficients) */
    const double c0 = neg ? -0x1p0 : 0x1p0, c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2, c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5,
                 c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10, c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        double x3 = x * x2;
        double t = fma(x2, s3, s2);
        double x7 = x3 * x2;
        double s = fma(x3, s1, x);
        return (float)fma(x7, t, s);
    } else {
        double x4 = x2 * x2;
        double q2 = fma(x2, c4, c3);
        double q1 = fma(x2, c1, c0);
        double x6 = x4 * x2;
        double c = fma(x4, c2, q1);
        return (float)fma(x6, q2, c);
    }
}
BPT_LM_HD uint32_t bpt_lm_abstop12(float x) { return (bpt_lm_asuint(x) >> 20) & 0x7ff; }
/* returns 0 and sets *ok = 0 when |y| >= 120 (or non-finite): the caller uses its double-precision fallback */
BPT_LM_HD float bpt_lm_sincosf(float y, int wantCos, int *ok) {
    double x = y;
    *ok = 1;
    if (bpt_lm_abstop12(y) < bpt_lm_abstop12(0x1.921FB6p-1f)) {   /* |y| < pi/4 */
        double s = x * x;
        if (bpt_lm_abstop12(y) < bpt_lm_abstop12(0x1p-12f)) return wantCos ? 1.0f : y;
        return bpt_lm_sinf_poly(x, s, 0, wantCos);
    }
    if (bpt_lm_abstop12(y) < bpt_lm_abstop12(120.0f)) {
        /* reduce_fast without TOINT_INTRINSICS: hpi_inv is prescaled by 2^24 */
        const double hpi_inv = 0x1.45f306dc9c883p+23, hpi = 0x1.921fb54442d18p+0;
        double r = x * hpi_inv;
        int n = ((int32_t)r + 0x800000) >> 24;
        x = fma(-(double)n, hpi, x);
        double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;   /* sign[4] = {1, -1, -1, 1} */
        int neg = (n & 2) ? 1 : 0;
        return bpt_lm_sinf_poly(x * sgn, x * x, neg, wantCos ? (n ^ 1) : n);
    }
    *ok = 0;
    return 0;
}
BPT_LM_HD float bpt_exact_sinf(float y) { int ok; float r = bpt_lm_sincosf(y, 0, &ok); return ok ? r : (float)sin((double)y); }
BPT_LM_HD float bpt_exact_cosf(float y) { int ok; float r = bpt_lm_sincosf(y, 1, &ok); return ok ? r : (float)cos((double)y); }
