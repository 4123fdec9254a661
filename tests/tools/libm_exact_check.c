/* Test tool: glibc 2.39's float logf / sinf / cosf restated in plain double arithmetic (the algorithms of its
 * sysdeps/ieee754/flt-32/{e_logf,s_sinf,s_cosf}.c, i.e. ARM optimized-routines, constants as found in libm.so.6's rodata),
 * compared EXHAUSTIVELY with the host libm: every positive float for logf, every float with |x| < 120 for sinf / cosf.
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
    return (badLog || badSin || badCos) ? 2 : 0;
}
