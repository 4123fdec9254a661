// Test tool: camera sampling of the product (csrc/camera.h: GetCameraSample, GenerateCameraRay, CameraRayDifferentials, Transform
// application with error bounds; csrc/sampling.h: ZSobol dimensions 0..5, PiecewiseConstant2D filter sampling, wavelength
// sampling) compiled for the CPU against golden vectors written by the reference itself (oracle/ref_dump.cpp "cam0_*" pinhole,
// "cam1_*" thin lens).  The FilterSampler tables are rebuilt here exactly as engine.cu's commit builds them.
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/camera_host_check.cpp -o /tmp/camcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "camera.h"
#include "spectrum.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b) || (a != a && b != b); }
static std::map<std::string, std::vector<float>> readBlob(const char *path, std::vector<uint32_t> *sobol = nullptr) {
    std::map<std::string, std::vector<float>> g;
    FILE *f = fopen(path, "rb"); char magic[8];
    if (!f || fread(magic, 1, 8, f) != 8) return g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; if (fread(&dt, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) break;
        std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) break;
        if (sobol && !strcmp(nm, "sobol32_2x52")) { sobol->resize(n); memcpy(sobol->data(), v.data(), 4 * n); }
        g[nm] = v; }
    fclose(f);
    return g;
}
// FilterSampler tables (filters.cpp FilterSampler ctor, util/sampling.h PiecewiseConstant1D/2D): same code as engine.cu
static void BuildPC1D(const float *f, int n, float mn, float mx, float *func, float *cdf, float *funcInt) {
    for (int i = 0; i < n; ++i) func[i] = std::abs(f[i]);
    cdf[0] = 0;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] * (mx - mn) / n;
    *funcInt = cdf[n];
    if (*funcInt == 0) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= *funcInt;
}
int main() {
    std::vector<uint32_t> sobol;
    auto g = readBlob("tests/golden/ref_vectors.bin");
    readBlob("pbrt-v4_b200/data/pbrt_tables.bin", &sobol);
    if (sobol.size() != 104) { printf("no sobol table\n"); return 3; }
    memcpy(h_sobol32, sobol.data(), sizeof(h_sobol32));
    int total = 0;
    for (int cfg = 0; cfg < 2; ++cfg) {
        const std::vector<float> &P = g[cfg ? "cam1_params" : "cam0_params"], &in = g[cfg ? "cam1_in" : "cam0_in"], &out = g[cfg ? "cam1_out" : "cam0_out"];
        if (P.empty()) { printf("missing golden\n"); return 3; }
        int xres = (int)P[0], yres = (int)P[1], spp = (int)P[2];
        DeviceCamera cam; std::memset(&cam, 0, sizeof(cam));
        cam.lensRadius = P[3]; cam.focalDistance = P[4];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { cam.cameraFromRaster.m[i][j] = P[5 + 4 * i + j]; cam.renderFromCamera.m[i][j] = P[21 + 4 * i + j]; }
        cam.dxCamera = Vec3(P[37], P[38], P[39]); cam.dyCamera = Vec3(P[40], P[41], P[42]);
        cam.rayDiffScale = stdmax(.125f, 1 / std::sqrt((float)spp));
        SamplerParams sp; std::memset(&sp, 0, sizeof(sp));
        sp.type = 0; sp.randomize = RANDOMIZE_FAST_OWEN; sp.seed = 0; sp.samplesPerPixel = spp;
        auto log2i = [](unsigned v) { int l = 0; while (v >>= 1) ++l; return l; };
        auto rup = [](int v) { int r = 1; while (r < v) r <<= 1; return r; };
        sp.log2SamplesPerPixel = log2i((unsigned)spp);
        sp.nBase4Digits = log2i((unsigned)rup(std::max(xres, yres))) + (sp.log2SamplesPerPixel + 1) / 2;
        // Gaussian filter, radius 1.5, sigma 0.5: tabulated at 32 samples per unit radius (engine.cu commit)
        float rx = 1.5f, ry = 1.5f, sigma = 0.5f;
        float expX = Gaussian(rx, 0, sigma), expY = Gaussian(ry, 0, sigma);
        int nx = int(32 * rx), ny = int(32 * ry);
        std::vector<float> f((size_t)nx * ny), func((size_t)nx * ny), cdf((size_t)(nx + 1) * ny), funcInt(ny), mfunc(ny), mcdf(ny + 1);
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) {
                float px = Lerp((x + 0.5f) / nx, -rx, rx), py = Lerp((y + 0.5f) / ny, -ry, ry);
                f[(size_t)y * nx + x] = stdmax(0.f, Gaussian(px, 0, sigma) - expX) * stdmax(0.f, Gaussian(py, 0, sigma) - expY);
            }
        for (int v = 0; v < ny; ++v) BuildPC1D(f.data() + (size_t)v * nx, nx, -rx, rx, func.data() + (size_t)v * nx, cdf.data() + (size_t)v * (nx + 1), &funcInt[v]);
        float mfuncInt;
        std::vector<float> marg(funcInt);
        BuildPC1D(marg.data(), ny, -ry, ry, mfunc.data(), mcdf.data(), &mfuncInt);
        DeviceFilter filt; std::memset(&filt, 0, sizeof(filt));
        filt.sampled = 1; filt.nx = nx; filt.ny = ny; filt.radiusX = rx; filt.radiusY = ry; filt.f = f.data();
        filt.distrib.func = func.data(); filt.distrib.cdf = cdf.data(); filt.distrib.funcInt = funcInt.data(); filt.distrib.mfunc = mfunc.data();
        filt.distrib.mcdf = mcdf.data(); filt.distrib.mfuncInt = mfuncInt; filt.distrib.nu = nx; filt.distrib.nv = ny;
        filt.distrib.min0 = -rx; filt.distrib.max0 = rx; filt.distrib.min1 = -ry; filt.distrib.max1 = ry;
        int n = (int)in.size() / 3, badS = 0, badR = 0, badD = 0, badW = 0;
        for (int q = 0; q < n; ++q) {
            const float *a = &in[3 * q], *o = &out[28 * q];
            Sampler sampler;
            CamSample cs = GetCameraSample(sp, filt, sampler, (int)a[0], (int)a[1], (int)a[2]);
            bool okS = eq(cs.lu, o[0]) && eq(cs.pFilm.x, o[1]) && eq(cs.pFilm.y, o[2]) && eq(cs.fweight, o[3]) && eq(cs.time, o[4]) && eq(cs.pLens.x, o[5]) && eq(cs.pLens.y, o[6]);
            if (!okS) { if (badS++ < 3) printf("cfg %d query %d: sample lu %g pFilm %g %g w %g  vs  %g %g %g %g\n", cfg, q, cs.lu, cs.pFilm.x, cs.pFilm.y, cs.fweight, o[0], o[1], o[2], o[3]); continue; }
            Wavelengths wl = Wavelengths::SampleVisible(cs.lu);
            if (!(eq(wl.lambda[0], o[25]) && eq(wl.lambda[3], o[26]) && eq(wl.pdf[0], o[27]))) { if (badW++ < 3) printf("cfg %d query %d: lambda %g %g pdf %g vs %g %g %g\n", cfg, q, wl.lambda[0], wl.lambda[3], wl.pdf[0], o[25], o[26], o[27]); }
            Pt3 ro; Vec3 rd;
            GenerateCameraRay(cam, cs.pFilm, cs.pLens, &ro, &rd);
            bool okR = eq(ro.x, o[7]) && eq(ro.y, o[8]) && eq(ro.z, o[9]) && eq(rd.x, o[10]) && eq(rd.y, o[11]) && eq(rd.z, o[12]);
            if (!okR) { if (badR++ < 3) printf("cfg %d query %d: ray o %g %g %g d %g %g %g  vs  %g %g %g  %g %g %g\n", cfg, q, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, o[7], o[8], o[9], o[10], o[11], o[12]); continue; }
            RayDiffs df = CameraRayDifferentials(cam, cs, ro, rd);
            const float mine[12] = {df.rxO.x, df.rxO.y, df.rxO.z, df.rxD.x, df.rxD.y, df.rxD.z, df.ryO.x, df.ryO.y, df.ryO.z, df.ryD.x, df.ryD.y, df.ryD.z};
            bool okD = true;
            for (int k = 0; k < 12; ++k) okD = okD && eq(mine[k], o[13 + k]);
            if (!okD) { if (badD++ < 3) printf("cfg %d query %d: differentials rxD %g %g %g vs %g %g %g\n", cfg, q, df.rxD.x, df.rxD.y, df.rxD.z, o[16], o[17], o[18]); }
        }
        printf("camera %d (%s): samples %d, wavelengths %d, rays %d, differentials %d mismatches of %d\n", cfg, cfg ? "thin lens" : "pinhole", badS, badW, badR, badD, n);
        total += badS + badW + badR + badD;
    }
    return total ? 2 : 0;
}
