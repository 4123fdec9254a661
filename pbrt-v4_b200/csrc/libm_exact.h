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
 * asinf / acosf / atanf / atan2f are glibc 2.39's fdlibm-derived float routines (sysdeps/ieee754/flt-32/e_asinf.c, e_acosf.c,
 * s_atanf.c, e_atan2f.c): plain float +, -, *, /, sqrtf -- IEEE operations that the baseline x86-64 build of libm cannot fuse,
 * so the device (-fmad=false, IEEE sqrt / div) reproduces them bit for bit.  Also verified exhaustively (asinf / acosf over
 * [-1, 1], atanf over every float, atan2f over 2e9 random pairs).
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

/* ------------------------------------------------------------------ asinf / acosf / atanf / atan2f (float arithmetic) */
BPT_LM_HD float bpt_lm_setlow(float f, uint32_t mask) { return bpt_lm_asfloat(bpt_lm_asuint(f) & mask); }
BPT_LM_HD float bpt_exact_asinf(float x) {
    const float one = 1.0f, huge = 1e30f, pio2_hi = 1.57079637050628662109375f, pio2_lo = -4.37113900018624283e-8f,
                pio4_hi = 0.785398185253143310546875f, p0 = 1.666675248e-1f, p1 = 7.495297643e-2f, p2 = 4.547037598e-2f,
                p3 = 2.417951451e-2f, p4 = 4.216630880e-2f;
    float t, w, p, q, c, r, s;
    int32_t hx = (int32_t)bpt_lm_asuint(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) return x * pio2_hi + x * pio2_lo;
    else if (ix > 0x3f800000) return (x - x) / (x - x);
    else if (ix < 0x3f000000) {
        if (ix < 0x32000000) { if (huge + x > one) return x; }
        else { t = x * x; w = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4)))); return x + x * w; }
    }
    w = one - fabsf(x);
    t = w * 0.5f;
    p = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
    s = sqrtf(t);
    if (ix >= 0x3F79999A) t = pio2_hi - (2.0f * (s + s * p) - pio2_lo);
    else {
        w = bpt_lm_setlow(s, 0xfffff000u);
        c = (t - w * w) / (s + w);
        r = p;
        p = 2.0f * s * r - (pio2_lo - 2.0f * c);
        q = pio4_hi - 2.0f * w;
        t = pio4_hi - (p - q);
    }
    return hx > 0 ? t : -t;
}
BPT_LM_HD float bpt_exact_acosf(float x) {
    const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f, pS0 = 1.6666667163e-01f,
                pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f, pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f,
                qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
    float z, p, q, r, w, s, c, df;
    int32_t hx = (int32_t)bpt_lm_asuint(x), ix = hx & 0x7fffffff;
    if (ix == 0x3f800000) { if (hx > 0) return 0.0f; else return pi + 2.0f * pio2_lo; }
    else if (ix > 0x3f800000) return (x - x) / (x - x);
    if (ix < 0x3f000000) {
        if (ix <= 0x23000000) return pio2_hi + pio2_lo;
        z = x * x;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        return pio2_hi - (x - (pio2_lo - x * r));
    } else if (hx < 0) {
        z = (one + x) * 0.5f;
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        s = sqrtf(z);
        r = p / q;
        w = r * s - pio2_lo;
        return pi - 2.0f * (s + w);
    } else {
        z = (one - x) * 0.5f;
        s = sqrtf(z);
        df = bpt_lm_setlow(s, 0xfffff000u);
        c = (z - df * df) / (s + df);
        p = z * (pS0 + z * (pS1 + z * (pS2 + z * (pS3 + z * (pS4 + z * pS5)))));
        q = one + z * (qS1 + z * (qS2 + z * (qS3 + z * qS4)));
        r = p / q;
        w = r * s + c;
        return 2.0f * (df + w);
    }
}
BPT_LM_HD float bpt_exact_atanf(float x) {
    const float one = 1.0f, huge = 1.0e30f;
    const float hi0 = 4.6364760399e-01f, hi1 = 7.8539812565e-01f, hi2 = 9.8279368877e-01f, hi3 = 1.5707962513e+00f;
    const float lo0 = 5.0121582440e-09f, lo1 = 3.7748947079e-08f, lo2 = 3.4473217170e-08f, lo3 = 7.5497894159e-08f;
    const float a0 = 3.3333334327e-01f, a1 = -2.0000000298e-01f, a2 = 1.4285714924e-01f, a3 = -1.1111110449e-01f, a4 = 9.0908870101e-02f,
                a5 = -7.6918758452e-02f, a6 = 6.6610731184e-02f, a7 = -5.8335702866e-02f, a8 = 4.9768779427e-02f, a9 = -3.6531571299e-02f,
                a10 = 1.6285819933e-02f;
    float w, s1, s2, z;
    int32_t hx = (int32_t)bpt_lm_asuint(x), ix = hx & 0x7fffffff, id;
    if (ix >= 0x4c000000) {   /* |x| >= 2^25 */
        if (ix > 0x7f800000) return x + x;
        return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    }
    if (ix < 0x3ee00000) {
        if (ix < 0x31000000) { if (huge + x > one) return x; }
        id = -1;
    } else {
        x = fabsf(x);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) { id = 0; x = (2.0f * x - one) / (2.0f + x); }
            else { id = 1; x = (x - one) / (x + one); }
        } else {
            if (ix < 0x401c0000) { id = 2; x = (x - 1.5f) / (one + 1.5f * x); }
            else { id = 3; x = -1.0f / x; }
        }
    }
    z = x * x;
    w = z * z;
    s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
    s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
    if (id < 0) return x - x * (s1 + s2);
    float hi = id == 0 ? hi0 : id == 1 ? hi1 : id == 2 ? hi2 : hi3, lo = id == 0 ? lo0 : id == 1 ? lo1 : id == 2 ? lo2 : lo3;
    z = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -z : z;
}
BPT_LM_HD float bpt_exact_atan2f(float y, float x) {
    const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
    float z;
    int32_t hx = (int32_t)bpt_lm_asuint(x), ix = hx & 0x7fffffff, hy = (int32_t)bpt_lm_asuint(y), iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;
    if (hx == 0x3f800000) return bpt_exact_atanf(y);
    int32_t m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0) { if (m < 2) return y; return m == 2 ? pi + tiny : -pi - tiny; }
    if (ix == 0) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) return m == 0 ? pi_o_4 + tiny : m == 1 ? -pi_o_4 - tiny : m == 2 ? 3.0f * pi_o_4 + tiny : -3.0f * pi_o_4 - tiny;
        return m == 0 ? 0.0f : m == 1 ? -0.0f : m == 2 ? pi + tiny : -pi - tiny;
    }
    if (iy == 0x7f800000) return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    int32_t k = (iy - ix) >> 23;
    if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60) z = 0.0f;
    else z = bpt_exact_atanf(fabsf(y / x));
    switch (m) {
    case 0: return z;
    case 1: return bpt_lm_asfloat(bpt_lm_asuint(z) ^ 0x80000000u);
    case 2: return pi - (z - pi_lo);
    default: return (z - pi_lo) - pi;
    }
}

/* ------------------------------------------------------------------ expf (EXP2F_TABLE_BITS = 5; |x| < 88: the common path) */
BPT_LM_HD uint64_t bpt_lm_exp2f_tab(unsigned i) {
    const uint64_t T[32] = {
        0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
        0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
        0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
        0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
        0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
        0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
        0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
        0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
    };
    return T[i];
}
/* *ok = 0 outside |x| < 88 (overflow / underflow / non-finite handling of the library is not restated) */
BPT_LM_HD float bpt_lm_expf(float x, int *ok) {
    *ok = 1;
    if (((bpt_lm_asuint(x) >> 20) & 0x7ff) >= ((bpt_lm_asuint(88.0f) >> 20) & 0x7ff)) { *ok = 0; return 0; }
    const double InvLn2N = 0x1.71547652b82fep+5, SHIFT = 0x1.8p+52, C0 = 0x1.c6af84b912394p-20, C1 = 0x1.ebfce50fac4f3p-13, C2 = 0x1.62e42ff0c52d6p-6;
    double xd = x;
    double z = InvLn2N * xd;
    double kd = z + SHIFT;
    uint64_t ki; memcpy(&ki, &kd, 8);
    kd -= SHIFT;
    double r = fma(InvLn2N, xd, -kd);        /* the FMA build fuses z's product into the subtraction */
    uint64_t t = bpt_lm_exp2f_tab((unsigned)(ki % 32)) + (ki << (52 - 5));
    double sc; memcpy(&sc, &t, 8);
    z = fma(C0, r, C1);
    double r2 = r * r;
    double y = fma(C2, r, 1.0);
    y = fma(z, r2, y);
    y = y * sc;
    return (float)y;
}
BPT_LM_HD float bpt_exact_expf(float x) { int ok; float r = bpt_lm_expf(x, &ok); return ok ? r : (float)exp((double)x); }

/* ------------------------------------------------------------------ log1pf, atanhf, coshf (fdlibm-derived float routines) */
BPT_LM_HD float bpt_exact_log1pf(float x) {
    const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, two25 = 3.355443200e+07f, Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f,
                Lp3 = 2.8571429849e-01f, Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f, Lp6 = 1.5313838422e-01f, Lp7 = 1.4798198640e-01f, zero = 0.0f;
    float hfsq, f = 0, c = 0, s, z, R, u;
    int32_t k = 1, hx = (int32_t)bpt_lm_asuint(x), hu = 0, ax = hx & 0x7fffffff;
    if (hx < 0x3ed413d7) {
        if (ax >= 0x3f800000) { if (x == -1.0f) return -two25 / zero; else return (x - x) / (x - x); }
        if (ax < 0x31000000) { if (ax < 0x24800000) return x; else return x - x * x * 0.5f; }
        if (hx > 0 || hx <= ((int32_t)0xbe95f61f)) { k = 0; f = x; hu = 1; }
    }
    if (hx >= 0x7f800000) return x + x;
    if (k != 0) {
        if (hx < 0x5a000000) {
            u = 1.0f + x;
            hu = (int32_t)bpt_lm_asuint(u);
            k = (hu >> 23) - 127;
            c = (k > 0) ? 1.0f - (u - x) : x - (u - 1.0f);
            c /= u;
        } else {
            u = x;
            hu = (int32_t)bpt_lm_asuint(u);
            k = (hu >> 23) - 127;
            c = 0;
        }
        hu &= 0x007fffff;
        if (hu < 0x3504f7) u = bpt_lm_asfloat((uint32_t)hu | 0x3f800000u);
        else { k += 1; u = bpt_lm_asfloat((uint32_t)hu | 0x3f000000u); hu = (0x00800000 - hu) >> 2; }
        f = u - 1.0f;
    }
    hfsq = 0.5f * f * f;
    if (hu == 0) {
        if (f == zero) { if (k == 0) return zero; c += k * ln2_lo; return k * ln2_hi + c; }
        R = hfsq * (1.0f - 0.66666666666666666f * f);
        if (k == 0) return f - R;
        return k * ln2_hi - ((R - (k * ln2_lo + c)) - f);
    }
    s = f / (2.0f + f);
    z = s * s;
    R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
    if (k == 0) return f - (hfsq - s * (hfsq + R));
    return k * ln2_hi - ((hfsq - (s * (hfsq + R) + (k * ln2_lo + c))) - f);
}
BPT_LM_HD float bpt_exact_atanhf(float x) {
    float xa = fabsf(x), t;
    if (xa < 0.5f) {
        if (xa < 0x1.0p-28f) return x;
        t = xa + xa;
        t = 0.5f * bpt_exact_log1pf(t + t * xa / (1.0f - xa));
    } else if (xa < 1.0f) t = 0.5f * bpt_exact_log1pf((xa + xa) / (1.0f - xa));
    else { if (xa > 1.0f) return (x - x) / (x - x); return copysignf(INFINITY, x); }   /* x / 0.0f */
    return copysignf(t, x);
}
BPT_LM_HD float bpt_lm_expm1f_small(float x) {   /* __expm1f for |x| <= 0.5 ln2 (the only range coshf uses) */
    const float one = 1.0f, huge = 1.0e30f, Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f,
                Q5 = -2.0109921195e-07f;
    uint32_t hx = bpt_lm_asuint(x) & 0x7fffffffu;
    float t, e, hxs, hfx, r1;
    if (hx < 0x33000000u) { t = huge + x; return x - (t - (huge + x)); }
    hfx = 0.5f * x;
    hxs = x * hfx;
    r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
    t = 3.0f - r1 * hfx;
    e = hxs * ((r1 - t) / (6.0f - x * t));
    return x - (x * e - hxs);
}
BPT_LM_HD float bpt_exact_coshf(float x) {
    const float one = 1.0f, half = 0.5f;
    int32_t ix = (int32_t)(bpt_lm_asuint(x) & 0x7fffffffu);
    int ok = 1;
    float t, w, r = 0;
    if (ix < 0x41b00000) {
        if (ix < 0x3eb17218) {
            if (ix < 0x24000000) return one;
            t = bpt_lm_expm1f_small(fabsf(x));
            w = one + t;
            return one + (t * t) / (w + w);
        }
        t = bpt_lm_expf(fabsf(x), &ok);
        r = half * t + half / t;
    } else if (ix < 0x42b17180) {
        t = bpt_lm_expf(fabsf(x), &ok);
        r = half * t;
    } else ok = 0;
    return ok ? r : (float)cosh((double)x);
}
