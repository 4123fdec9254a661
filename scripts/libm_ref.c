/* glibc float transcendental values for the inputs in a binary file: usage libm_ref <fn> <in.bin> <out.bin>; in = n x (x,y) floats */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
int main(int argc, char **argv) {
    int fn = atoi(argv[1]);
    FILE *f = fopen(argv[2], "rb"); fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    int n = sz / 8; float *xy = malloc(sz); fread(xy, 1, sz, f); fclose(f);
    float *o = malloc(4 * n);
    for (int i = 0; i < n; ++i) {
        float a = xy[2 * i], b = xy[2 * i + 1], r = 0;
        switch (fn % 100) {
        case 0: r = sinf(a); break; case 1: r = cosf(a); break; case 2: r = asinf(a); break; case 3: r = acosf(a); break;
        case 4: r = atan2f(a, b); break; case 5: r = atanhf(a); break; case 6: r = coshf(a); break; case 7: r = expf(a); break;
        case 8: r = logf(a); break; case 9: r = tanf(a); break; case 19: r = powf(a, b); break; case 21: r = sqrtf(a); break; case 22: r = a / b; break;
        }
        o[i] = r;
    }
    f = fopen(argv[3], "wb"); fwrite(o, 4, n, f); fclose(f);
    return 0;
}
