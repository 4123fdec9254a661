// Diagnostic (tests only): evaluates the product's __host__ __device__ infinite-light code on the CPU against the
// reference's golden vectors (tests/golden/ref_vectors.bin: env_*).  Build:
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/envmap_host_check.cpp -o /tmp/envcheck
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include "lights.h"
#include "tables.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }

static void BuildPC1D(const float *f, int n, float mn, float mx, float *func, float *cdf, float *funcInt) {
    for (int i = 0; i < n; ++i) func[i] = std::abs(f[i]);
    cdf[0] = 0;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] * (mx - mn) / n;
    *funcInt = cdf[n];
    if (*funcInt == 0) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= *funcInt;
}
struct PC2D { std::vector<float> func, cdf, funcInt, mfunc, mcdf; float mfuncInt; PC2DView v; };
static void BuildPC2D(const std::vector<float> &f, int nu, int nv, PC2D *t) {
    t->func.resize((size_t)nu * nv); t->cdf.resize((size_t)(nu + 1) * nv); t->funcInt.resize(nv);
    for (int v = 0; v < nv; ++v) BuildPC1D(f.data() + (size_t)v * nu, nu, 0, 1, t->func.data() + (size_t)v * nu, t->cdf.data() + (size_t)v * (nu + 1), &t->funcInt[v]);
    t->mfunc.resize(nv); t->mcdf.resize(nv + 1);
    std::vector<float> marg(t->funcInt);
    BuildPC1D(marg.data(), nv, 0, 1, t->mfunc.data(), t->mcdf.data(), &t->mfuncInt);
    t->v = PC2DView{t->func.data(), t->cdf.data(), t->funcInt.data(), t->mfunc.data(), t->mcdf.data(), t->mfuncInt, nu, nv, 0, 1, 0, 1};
}
int main() {
    Tables T, G; std::string err;
    if (!T.Load("pbrt-v4_b200/data/pbrt_tables.bin", &err)) { printf("%s\n", err.c_str()); return 1; }
    // golden blob has the same container format
    FILE *f = fopen("tests/golden/ref_vectors.bin", "rb"); char magic[8]; fread(magic, 1, 8, f);
    std::map<std::string, std::vector<float>> g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; fread(&dt, 4, 1, f); fread(&n, 8, 1, f); std::vector<float> v(n); fread(v.data(), 4, n, f); g[nm] = v; }
    fclose(f);
    const int res = 32;
    std::vector<float> &pix = g["env_pixels"], &xf = g["env_xform"], &in = g["env_in"], &out = g["env_out"];
    std::vector<float> d((size_t)res * res);
    for (size_t i = 0; i < d.size(); ++i) { float s = 0; s += pix[3 * i]; s += pix[3 * i + 1]; s += pix[3 * i + 2]; d[i] = (s / 3) * 1.f; }
    PC2D dist, comp;
    BuildPC2D(d, res, res, &dist);
    double acc = 0; for (float v : d) acc += v;
    float average = (float)(acc / d.size());
    for (float &v : d) v = stdmax(v - average, 0.f);
    BuildPC2D(d, res, res, &comp);
    DeviceImageLight im{}; im.rgb = pix.data(); im.res = res; im.distrib = dist.v; im.compensatedDistrib = comp.v;
    SceneView sv{}; sv.imageLights = &im; sv.illumD65 = T.F("srgb_illuminant").data();
    sv.rgb2specScale = T.F("rgb2spec_srgb_scale").data(); sv.rgb2specCoeffs = T.F("rgb2spec_srgb_coeffs").data();
    DeviceLight L{}; L.type = LIGHT_IMAGE_INFINITE; L.image = 0;
    float photometric = 0; for (int i = 0; i < kNLambda; ++i) photometric += T.F("cie_y")[i] * T.F("srgb_illuminant")[i];
    L.scale = 0.37f;   // the golden light is constructed with scale 0.37 directly (no photometric normalisation)
    memcpy(L.renderFromLight.m, xf.data(), 64); memcpy(L.lightFromRender.m, xf.data() + 16, 64);
    // Preprocess(Bounds3f((-1,-2,-3),(2,3,5)))
    Pt3 bmn(-1, -2, -3), bmx(2, 3, 5); Pt3 c = (bmn + bmx) / 2; L.sceneRadius = Distance(c, bmx);
    int bad[6] = {0, 0, 0, 0, 0, 0}; int n = (int)in.size() / 12;
    for (int i = 0; i < n; ++i) {
        const float *a = &in[12 * i], *o = &out[22 * i];
        Wavelengths wl; for (int k = 0; k < 4; ++k) { wl.lambda[k] = a[5 + k]; wl.pdf[k] = 1; }
        Vec3 dir(a[0], a[1], a[2]);
        Spec Le = InfiniteLightLe(sv, L, dir, wl);
        for (int k = 0; k < 4; ++k) if (f2b(Le[k]) != f2b(o[k])) { if (bad[0]++ < 3) printf("Le mismatch %d: %g vs %g\n", i, Le[k], o[k]); break; }
        LightSampleContext ctx; ctx.pi = Pt3i(Pt3(a[9], a[10], a[11])); ctx.n = ctx.ns = Nrm3(0, 0, 0);
        float pdf = InfiniteLightPDF_Li(sv, L, ctx, dir);
        if (f2b(pdf) != f2b(o[4])) { if (bad[1]++ < 3) printf("pdf mismatch %d: %g vs %g\n", i, pdf, o[4]); }
        Vec2 sq = EqualAreaSphereToSquare(dir);
        if (f2b(sq.x) != f2b(o[5]) || f2b(sq.y) != f2b(o[6])) { if (bad[2]++ < 3) printf("sphere->square mismatch %d\n", i); }
        Vec3 back = EqualAreaSquareToSphere(Vec2(a[3], a[4]));
        if (f2b(back.x) != f2b(o[7]) || f2b(back.y) != f2b(o[8]) || f2b(back.z) != f2b(o[9])) { if (bad[3]++ < 3) printf("square->sphere mismatch %d: %g %g %g vs %g %g %g\n", i, back.x, back.y, back.z, o[7], o[8], o[9]); }
        LightLiSample ls = InfiniteLightSampleLi(sv, L, ctx, Vec2(a[3], a[4]), wl);
        if (ls.valid != (o[10] == 1)) bad[4]++;
        else if (ls.valid) {
            bool ok = f2b(ls.pdf) == f2b(o[18]) && f2b(ls.wi.x) == f2b(o[15]) && f2b(ls.wi.y) == f2b(o[16]) && f2b(ls.wi.z) == f2b(o[17]);
            for (int k = 0; k < 4; ++k) ok = ok && f2b(ls.L[k]) == f2b(o[11 + k]);
            Pt3 pl = ls.pLight.p();
            ok = ok && f2b(pl.x) == f2b(o[19]) && f2b(pl.y) == f2b(o[20]) && f2b(pl.z) == f2b(o[21]);
            if (!ok) { if (bad[5]++ < 3) printf("SampleLi mismatch %d: pdf %g vs %g, wi %g %g %g vs %g %g %g, L0 %g vs %g\n", i, ls.pdf, o[18], ls.wi.x, ls.wi.y, ls.wi.z, o[15], o[16], o[17], ls.L[0], o[11]); }
        }
    }
    printf("n=%d mismatches: Le %d pdf %d s2sq %d sq2s %d valid %d SampleLi %d\n", n, bad[0], bad[1], bad[2], bad[3], bad[4], bad[5]);
    return (bad[0] + bad[1] + bad[2] + bad[3] + bad[4] + bad[5]) ? 2 : 0;
}
