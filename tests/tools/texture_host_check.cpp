// Test tool: image textures and the differentials texture filtering reads, of the product (csrc/texture_host.h: pyramid generation;
// csrc/texture.h: MIPMap::Filter<RGB> / <Float> with every filter and wrap mode; csrc/camera.h: FindMinimumDifferentials;
// csrc/differentials.h: Approximate_dp_dxy, ComputeDifferentials, SpawnRay's specular differentials) compiled for the CPU against
// golden vectors written by the reference itself (oracle/ref_dump.cpp "tex*", "diff*").
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/texture_host_check.cpp -o /tmp/texcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "differentials.h"
#include "texture_host.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b) || (a != a && b != b); }
static std::map<std::string, std::vector<float>> readBlob(const char *path) {
    std::map<std::string, std::vector<float>> g;
    FILE *f = fopen(path, "rb"); char magic[8];
    if (!f || fread(magic, 1, 8, f) != 8) return g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; if (fread(&dt, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) break;
        std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) break;
        g[nm] = v; }
    fclose(f);
    return g;
}
int main() {
    auto g = readBlob("tests/golden/ref_vectors.bin");
    auto tb = readBlob("pbrt-v4_b200/data/pbrt_tables.bin");
    if (!tb.count("srgb8_to_linear") || !tb.count("mip_filter_lut") || !g.count("tex_queries")) { printf("missing tables / golden\n"); return 3; }
    int total = 0;
    // ---- textures
    struct Img { int w, h, nc; bool is8; };
    const Img imgs[3] = {{13, 9, 3, true}, {16, 8, 1, false}, {8, 8, 4, true}};
    const std::vector<float> &q = g["tex_queries"];
    const int nQ = (int)q.size() / 6;
    texhost::Pyramid pyr{tb["srgb8_to_linear"].data()};
    for (int k = 0; k < 3; ++k) {
        char nm[32];
        snprintf(nm, sizeof(nm), "tex%d_pixels", k); const std::vector<float> &pix = g[nm];
        snprintf(nm, sizeof(nm), "tex%d_out", k); const std::vector<float> &out = g[nm];
        HostImage im; im.xres = imgs[k].w; im.yres = imgs[k].h; im.nc = imgs[k].nc; im.is8 = imgs[k].is8; im.srgb = true;
        if (im.is8) { im.p8.resize(pix.size()); for (size_t i = 0; i < pix.size(); ++i) im.p8[i] = (uint8_t)pix[i]; } else im.p32 = pix;
        int bad = 0, n = 0;
        for (int wrap = 0; wrap < 3; ++wrap)
            for (int filt = 0; filt < 4; ++filt) {
                const float *o = &out[(size_t)(wrap * 4 + filt) * nQ * 4];
                if (wrap == 0 && k == 0) continue;
                std::vector<HostMIPLevel> py = pyr.Generate(im, wrap);
                std::vector<DeviceMIPLevel> lv;
                for (auto &l : py) lv.push_back(DeviceMIPLevel{l.texels.data(), l.xres, l.yres});
                DeviceTexture t{}; t.firstLevel = 0; t.nLevels = (int)lv.size(); t.nc = im.nc; t.wrap = wrap; t.filter = filt; t.maxAnisotropy = 8.f;
                for (int i = 0; i < nQ; ++i, ++n) {
                    const float *a = &q[6 * i];
                    TexVal<true> c = MipFilter<true>(t, lv.data(), tb["mip_filter_lut"].data(), Vec2(a[0], a[1]), Vec2(a[2], a[3]), Vec2(a[4], a[5]));
                    TexVal<false> f = MipFilter<false>(t, lv.data(), tb["mip_filter_lut"].data(), Vec2(a[0], a[1]), Vec2(a[2], a[3]), Vec2(a[4], a[5]));
                    bool ok = eq(c.v[0], o[4 * i]) && eq(c.v[1], o[4 * i + 1]) && eq(c.v[2], o[4 * i + 2]) && eq(f.v[0], o[4 * i + 3]);
                    if (!ok && bad++ < 4)
                        printf("image %d wrap %d filter %d query %d (st %g %g d %g %g %g %g): %.9g %.9g %.9g | %.9g   vs   %.9g %.9g %.9g | %.9g\n", k, wrap, filt, i, a[0], a[1], a[2],
                               a[3], a[4], a[5], c.v[0], c.v[1], c.v[2], f.v[0], o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                }
            }
        printf("texture image %d (%dx%d, %d channels, %s): %d mismatches of %d lookups\n", k, im.xres, im.yres, im.nc, im.is8 ? "8-bit sRGB" : "float", bad, n);
        total += bad;
    }
    // ---- differentials
    for (int cfg = 0; cfg < 2; ++cfg) {
        char nm[32];
        snprintf(nm, sizeof(nm), "diff%d_params", cfg); const std::vector<float> &P = g[nm];
        snprintf(nm, sizeof(nm), "diff%d_in", cfg); const std::vector<float> &in = g[nm];
        snprintf(nm, sizeof(nm), "diff%d_out", cfg); const std::vector<float> &out = g[nm];
        if (P.empty()) { printf("missing golden\n"); return 3; }
        int xres = (int)P[0], yres = (int)P[1], spp = (int)P[2];
        DeviceCamera cam; std::memset(&cam, 0, sizeof(cam));
        cam.lensRadius = P[3]; cam.focalDistance = P[4];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
            cam.cameraFromRaster.m[i][j] = P[5 + 4 * i + j]; cam.renderFromCamera.m[i][j] = P[21 + 4 * i + j]; cam.cameraFromRender.m[i][j] = P[37 + 4 * i + j]; }
        cam.dxCamera = Vec3(P[53], P[54], P[55]); cam.dyCamera = Vec3(P[56], P[57], P[58]);
        cam.rayDiffScale = stdmax(.125f, 1 / std::sqrt((float)spp));
        cam.sppScale = stdmax(.125f, 1 / std::sqrt((float)spp));
        FindMinimumDifferentials(&cam, xres, yres);
        const Vec3 *mine[4] = {&cam.minPosDifferentialX, &cam.minPosDifferentialY, &cam.minDirDifferentialX, &cam.minDirDifferentialY};
        int badMin = 0;
        for (int k = 0; k < 4; ++k)
            if (!(eq(mine[k]->x, P[59 + 3 * k]) && eq(mine[k]->y, P[60 + 3 * k]) && eq(mine[k]->z, P[61 + 3 * k]))) {
                ++badMin; printf("camera %d min differential %d: %.9g %.9g %.9g vs %.9g %.9g %.9g\n", cfg, k, mine[k]->x, mine[k]->y, mine[k]->z, P[59 + 3 * k], P[60 + 3 * k], P[61 + 3 * k]);
            }
        const int NI = 42, NO = 23;
        int n = (int)in.size() / NI, badD = 0, badS = 0;
        for (int qi = 0; qi < n; ++qi) {
            const float *a = &in[(size_t)NI * qi], *o = &out[(size_t)NO * qi];
            SurfaceIntr si; SurfaceExtras ex;
            si.pi = Pt3i(Pt3(a[0], a[1], a[2]));
            ex.dpdu = Vec3(a[3], a[4], a[5]); ex.dpdv = Vec3(a[6], a[7], a[8]);
            ex.dndus = Nrm3(a[9], a[10], a[11]); ex.dndvs = Nrm3(a[12], a[13], a[14]);
            Vec3 wo = Normalize(Vec3(a[15], a[16], a[17]));   // Interaction::wo = Normalize(wo) (interaction.h:32-33)
            RayDiffState ray; ray.has = a[18] != 0;
            ray.rxO = Pt3(a[19], a[20], a[21]); ray.ryO = Pt3(a[22], a[23], a[24]); ray.rxD = Vec3(a[25], a[26], a[27]); ray.ryD = Vec3(a[28], a[29], a[30]);
            Vec3 wi(a[31], a[32], a[33]);
            float eta = a[34]; int flags = (int)a[35];
            si.n = Nrm3(a[36], a[37], a[38]); si.ns = Nrm3(a[39], a[40], a[41]);
            UVDiffs d = ComputeDifferentials(cam, si, ex, ray);
            const float md[10] = {d.dpdx.x, d.dpdx.y, d.dpdx.z, d.dpdy.x, d.dpdy.y, d.dpdy.z, d.dudx, d.dudy, d.dvdx, d.dvdy};
            bool ok = true;
            for (int k = 0; k < 10; ++k) ok = ok && eq(md[k], o[k]);
            if (!ok) { if (badD++ < 4) printf("camera %d query %d (%s): dpdx %.9g %.9g %.9g dudx %.9g  vs  %.9g %.9g %.9g %.9g\n", cfg, qi, ray.has ? "ray differentials" : "approximate",
                                             md[0], md[1], md[2], md[6], o[0], o[1], o[2], o[6]); continue; }
            // the engine's flag values are the reference's (bxdfs.h BxDFFlags)
            RayDiffState rd = SpawnRayDifferentials(ray, si.p(), wo, wi, si.ns, ex.dndus, ex.dndvs, d, flags, eta);
            const float ms[13] = {rd.has ? 1.f : 0.f, rd.rxO.x, rd.rxO.y, rd.rxO.z, rd.ryO.x, rd.ryO.y, rd.ryO.z, rd.rxD.x, rd.rxD.y, rd.rxD.z, rd.ryD.x, rd.ryD.y, rd.ryD.z};
            ok = eq(ms[0], o[10]);
            if (rd.has) for (int k = 1; k < 13; ++k) ok = ok && eq(ms[k], o[10 + k]);
            if (!ok && badS++ < 4) printf("camera %d query %d flags %d: spawn has %g rxD %.9g %.9g %.9g  vs  %g %.9g %.9g %.9g\n", cfg, qi, flags, ms[0], ms[7], ms[8], ms[9], o[10], o[17], o[18], o[19]);
        }
        printf("camera %d: min differentials %d, ComputeDifferentials %d, SpawnRay differentials %d mismatches of %d\n", cfg, badMin, badD, badS, n);
        total += badMin + badD + badS;
    }
    return total ? 2 : 0;
}
