// Test tool: the product's RGBGridMedium (csrc/media.h: RGBGridLookup, MediumSamplePoint / MediumSampleRay for MEDIUM_RGBGRID; the voxel
// conversion and majorant grid exactly as engine.cu's commit builds them) compiled for the CPU against golden vectors written by the
// reference itself (oracle/ref_dump.cpp "rgbmed_*": majorant grid and SampleT_maj event traces with sigma_a, sigma_s, Le).
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/rgbmedium_host_check.cpp -o /tmp/rgbmedcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "media.h"
#include "lights.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b) || (a != a && b != b); }
static std::map<std::string, std::vector<float>> readBlob(const char *path) {
    std::map<std::string, std::vector<float>> g;
    FILE *f = fopen(path, "rb"); char magic[8];
    if (!f || fread(magic, 1, 8, f) != 8) return g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; if (fread(&dt, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) break;
        std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) break; g[nm] = v; }
    fclose(f);
    return g;
}
int main() {
    auto g = readBlob("tests/golden/ref_vectors.bin");
    auto tb = readBlob("pbrt-v4_b200/data/pbrt_tables.bin");
    if (!g.count("rgbmed_params")) { printf("missing golden\n"); return 3; }
    const std::vector<float> &P = g["rgbmed_params"];
    const int nx = (int)P[0], ny = (int)P[1], nz = (int)P[2];
    const size_t nv = (size_t)nx * ny * nz;
    SceneView sv; std::memset(&sv, 0, sizeof(sv));
    sv.rgb2specScale = tb["rgb2spec_srgb_scale"].data(); sv.rgb2specCoeffs = tb["rgb2spec_srgb_coeffs"].data();
    auto voxels = [&](const std::vector<float> &rgb, std::vector<float4> *out, std::vector<float> *maxv) {
        out->resize(nv); if (maxv) maxv->resize(nv);
        for (size_t v = 0; v < nv; ++v) {
            const float *c = &rgb[3 * v];
            float m = stdmax(stdmax(c[0], c[1]), c[2]), scale = 2 * m, q[3] = {0, 0, 0}, co[3];
            if (scale) { q[0] = c[0] / scale; q[1] = c[1] / scale; q[2] = c[2] / scale; }
            RGBToCoeffsDev(sv, q, co);
            (*out)[v] = make_float4(scale, co[0], co[1], co[2]);
            if (maxv) {
                float result = stdmax(SigmoidPoly(co[0], co[1], co[2], 360.f), SigmoidPoly(co[0], co[1], co[2], 830.f));
                float lambda = -co[1] / (2 * co[0]);
                if (lambda >= 360 && lambda <= 830) result = stdmax(result, SigmoidPoly(co[0], co[1], co[2], lambda));
                (*maxv)[v] = scale * result;
            }
        }
    };
    std::vector<float4> va, vs, vl; std::vector<float> ma, ms;
    voxels(g["rgbmed_sigma_a"], &va, &ma); voxels(g["rgbmed_sigma_s"], &vs, &ms); voxels(g["rgbmed_Le"], &vl, nullptr);
    const int R = 16;
    std::vector<float> maj((size_t)R * R * R), majA(maj.size()), majS(maj.size());
    ComputeMajorantGrid(ma.data(), nx, ny, nz, R, majA.data()); ComputeMajorantGrid(ms.data(), nx, ny, nz, R, majS.data());
    int badMaj = 0;
    for (size_t c = 0; c < maj.size(); ++c) { float maxSigma_t = majA[c] + majS[c]; maj[c] = P[9] * maxSigma_t; if (!eq(maj[c], g["rgbmed_majorant"][c])) ++badMaj; }
    printf("majorant grid mismatches %d / %zu\n", badMaj, maj.size());
    std::vector<float> pool = tb["srgb_illuminant"];
    DeviceMedium med; std::memset(&med, 0, sizeof(med));
    med.type = MEDIUM_RGBGRID; med.g = 0.2f; med.sigmaScale = P[9]; med.rgbLeScale = P[10]; med.Le = -1; med.illum = 0;
    for (int k = 0; k < 3; ++k) { med.bmin[k] = P[3 + k]; med.bmax[k] = P[6 + k]; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) med.mInv.m[i][j] = P[11 + 4 * i + j];
    med.nx = nx; med.ny = ny; med.nz = nz; med.rgbA = va.data(); med.rgbS = vs.data(); med.rgbLe = vl.data();
    med.majorant = maj.data(); med.mres[0] = med.mres[1] = med.mres[2] = R;
    const int kEvents = 5, rec = 1 + 12 * kEvents + 5;
    const std::vector<float> &in = g["rgbmed_in"], &out = g["rgbmed_out"];
    int n = (int)in.size() / 14, nbad = 0, nEvents = 0;
    for (int i = 0; i < n; ++i) {
        const float *a = &in[14 * i], *o = &out[(size_t)rec * i];
        Pt3 ro(a[0], a[1], a[2]); Vec3 rd(a[3], a[4], a[5]);
        Wavelengths wl; for (int k = 0; k < 4; ++k) { wl.lambda[k] = a[10 + k]; wl.pdf[k] = 1; }
        RNG rng((uint64_t)a[8], (uint64_t)a[9]);
        std::vector<float> ev; int nEv = 0;
        Spec T = SampleT_maj(med, pool.data(), ro, rd, a[6], a[7], rng, wl,
            [&](Pt3 p, const MediumProps &mp, Spec sm, Spec Tm) -> bool {
                float r[12] = {p.x, p.y, p.z, mp.sigma_a[0], mp.sigma_a[3], mp.sigma_s[0], mp.sigma_s[2], mp.Le[0], mp.Le[1], sm[0], Tm[0], Tm[3]};
                ev.insert(ev.end(), r, r + 12);
                return ++nEv < kEvents;
            });
        ev.resize(12 * kEvents, 0.f);
        bool ok = nEv == (int)o[0];
        int firstBad = -1;
        for (int k = 0; ok && k < 12 * kEvents; ++k) if (!eq(ev[k], o[1 + k])) { ok = false; firstBad = k; }
        for (int k = 0; ok && k < 4; ++k) ok = eq(T[k], o[1 + 12 * kEvents + k]);
        if (ok) ok = eq(rng.UniformFloat(), o[1 + 12 * kEvents + 4]);
        nEvents += nEv;
        if (!ok && nbad++ < 4) printf("ray %d: events %d vs %g; first bad field %d: %.9g vs %.9g\n", i, nEv, o[0], firstBad, firstBad >= 0 ? ev[firstBad] : 0.f, firstBad >= 0 ? o[1 + firstBad] : 0.f);
    }
    printf("RGB grid SampleT_maj mismatches %d / %d rays (%d events)\n", nbad, n, nEvents);
    return (badMaj + nbad) ? 2 : 0;
}
