/* Test tool: glibc 2.39's float logf / sinf / cosf restated in plain double arithmetic (the algorithms of its
 * sysdeps/ieee754/flt-32/{e_logf,s_sinf,s_cosf}.c, i.e. ARM optimized-routines, constants as found in libm.so.6's rodata),
 * compared EXHAUSTIVELY with the host libm: every positive float for logf, every float with |x| < 120 for sinf / cosf; likewise
 * its fdlibm-derived float asinf / acosf (all of [-1, 1]), atanf (every float), atan2f (2e9 pseudo-random pairs), log1pf, atanhf
 * (all of (-1, 1)), and expf / coshf (all |x| < 88).
 * The restatement is what pmath.h uses on the device when BPT_EXACT_LIBM is set (csrc/libm_exact.h).
 *   gcc -O2 -ffp-contract=off -mfma -fopenmp tests/tools/libm_exact_check.c -Ipbrt-v4_b200/csrc -o /tmp/libmcheck -lm
 *   /tmp/libmcheck [stride]      (stride > 1 samples every stride-th float: the CPU test suite uses a stride)          */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libm_exact.h"

static inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main(int argc, char **argv) {
    uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 1;
    long badLog = 0, badSin = 0, badCos = 0, nLog = 0, nTrig = 0;
#pragma omp parallel for reduction(+ : badLog, nLog) schedule(static, 1 << 16)
    for (int64_t i = 0x00000001; i < 0x7f800000; i += stride) {
        float x = asfloat((uint32_t)i);
        float a = logf(x), b = bpt_exact_logf(x);
        if (asuint(a) != asuint(b)) ++badLog;
        ++nLog;
    }
    const uint32_t lim = asuint(120.0f);
#pragma omp parallel for reduction(+ : badSin, badCos, nTrig) schedule(static, 1 << 16)
    for (int64_t i = 0; i < (int64_t)lim; i += stride) {
        for (int sgn = 0; sgn < 2; ++sgn) {
            float x = asfloat((uint32_t)i | (sgn ? 0x80000000u : 0u));
            if (asuint(sinf(x)) != asuint(bpt_exact_sinf(x))) ++badSin;
            if (asuint(cosf(x)) != asuint(bpt_exact_cosf(x))) ++badCos;
            ++nTrig;
        }
    }
    printf("logf mismatches %ld / %ld\nsinf mismatches %ld / %ld\ncosf mismatches %ld / %ld\n", badLog, nLog, badSin, nTrig, badCos, nTrig);
    long badAsin = 0, badAcos = 0, badAtan = 0, badAtan2 = 0, nA = 0, nT = 0, n2 = 0;
#pragma omp parallel for reduction(+ : badAsin, badAcos, nA) schedule(static, 1 << 16)
    for (int64_t i = 0; i <= 0x3f800000; i += stride)
        for (int sgn = 0; sgn < 2; ++sgn) {
            float x = asfloat((uint32_t)i | (sgn ? 0x80000000u : 0u));
            if (asuint(asinf(x)) != asuint(bpt_exact_asinf(x))) ++badAsin;
            if (asuint(acosf(x)) != asuint(bpt_exact_acosf(x))) ++badAcos;
            ++nA;
        }
#pragma omp parallel for reduction(+ : badAtan, nT) schedule(static, 1 << 16)
    for (int64_t i = 0; i < 0x7f800000; i += stride)
        for (int sgn = 0; sgn < 2; ++sgn) {
            float x = asfloat((uint32_t)i | (sgn ? 0x80000000u : 0u));
            if (asuint(atanf(x)) != asuint(bpt_exact_atanf(x))) ++badAtan;
            ++nT;
        }
    const int64_t nPairs = 2000000000LL / stride;
#pragma omp parallel for reduction(+ : badAtan2, n2) schedule(static, 1 << 16)
    for (int64_t i = 0; i < nPairs; ++i) {
        uint64_t h = (uint64_t)i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        uint32_t a = (uint32_t)h, b = (uint32_t)(h >> 32) * 2654435761u ^ (uint32_t)i;
        float y = asfloat(a), x = asfloat(b);
        if (i & 1) {   /* moderate magnitudes: exponents 2^-15 .. 2^16 */
            y = asfloat((a & 0x807fffffu) | ((0x70u + ((a >> 23) & 0x1f)) << 23));
            x = asfloat((b & 0x807fffffu) | ((0x70u + ((b >> 23) & 0x1f)) << 23));
        }
        float r1 = atan2f(y, x), r2 = bpt_exact_atan2f(y, x);
        if (asuint(r1) != asuint(r2) && !(r1 != r1 && r2 != r2)) ++badAtan2;
        ++n2;
    }
    printf("asinf mismatches %ld / %ld\nacosf mismatches %ld / %ld\natanf mismatches %ld / %ld\natan2f mismatches %ld / %ld\n", badAsin, nA, badAcos, nA,
           badAtan, nT, badAtan2, n2);
    long badExp = 0, badCosh = 0, badL1p = 0, badAtanh = 0, nE = 0, nL = 0, nH = 0;
    const uint32_t lim88 = asuint(88.0f) & 0xfff00000u;
#pragma omp parallel for reduction(+ : badExp, badCosh, nE) schedule(static, 1 << 16)
    for (int64_t i = 0; i < (int64_t)lim88; i += stride)
        for (int sgn = 0; sgn < 2; ++sgn) {
            float x = asfloat((uint32_t)i | (sgn ? 0x80000000u : 0u));
            if (asuint(expf(x)) != asuint(bpt_exact_expf(x))) ++badExp;
            if (asuint(coshf(x)) != asuint(bpt_exact_coshf(x))) ++badCosh;
            ++nE;
        }
#pragma omp parallel for reduction(+ : badL1p, nL) schedule(static, 1 << 16)
    for (int64_t i = 0; i < 0x7f800000; i += stride)
        for (int sgn = 0; sgn < 2; ++sgn) {
            if (sgn && i > 0x3f800000) continue;
            float x = asfloat((uint32_t)i | (sgn ? 0x80000000u : 0u));
            float a = log1pf(x), b = bpt_exact_log1pf(x);
            if (asuint(a) != asuint(b) && !(a != a && b != b)) ++badL1p;
            ++nL;
        }
#pragma omp parallel for reduction(+ : badAtanh, nH) schedule(static, 1 << 16)
    for (int64_t i = 0; i < 0x3f800000; i += stride)
        for (int sgn = 0; sgn < 2; ++sgn) {
            float x = asfloat((uint32_t)i | (sgn ? 0x80000000u : 0u));
            if (asuint(atanhf(x)) != asuint(bpt_exact_atanhf(x))) ++badAtanh;
            ++nH;
        }
    printf("expf mismatches %ld / %ld\ncoshf mismatches %ld / %ld\nlog1pf mismatches %ld / %ld\natanhf mismatches %ld / %ld\n", badExp, nE, badCosh, nE,
           badL1p, nL, badAtanh, nH);
    return (badLog || badSin || badCos || badAsin || badAcos || badAtan || badAtan2 || badExp || badCosh || badL1p || badAtanh) ? 2 : 0;
}
