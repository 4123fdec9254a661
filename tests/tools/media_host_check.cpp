// Test tool: the product's media code (csrc/media.h) compiled for the CPU against golden vectors written by the reference itself
// (oracle/ref_dump.cpp "medium_*", "hg_*": GridMedium / HomogeneousMedium SampleT_maj event traces, majorant grid, HG sampling).
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/media_host_check.cpp -o /tmp/mediacheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "media.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b) || (a != a && b != b); }
int main() {
    FILE *f = fopen("tests/golden/ref_vectors.bin", "rb"); char magic[8]; if (!f || fread(magic, 1, 8, f) != 8) return 3;
    std::map<std::string, std::vector<float>> g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; if (fread(&dt, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) break; std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) break; g[nm] = v; }
    fclose(f);
    const std::vector<float> &P = g["medium_params"], &dens = g["medium_density"];
    int nx = (int)P[0], ny = (int)P[1], nz = (int)P[2];
    // spectrum pool: grid sigma_a, sigma_s, homogeneous sigma_a, sigma_s
    std::vector<float> pool;
    for (const char *k : {"medium_sigma_a", "medium_sigma_s", "medium_h_sigma_a", "medium_h_sigma_s"}) pool.insert(pool.end(), g[k].begin(), g[k].end());
    int bad = 0;
    std::vector<float> maj(16 * 16 * 16);
    ComputeMajorantGrid(dens.data(), nx, ny, nz, 16, maj.data());
    int badMaj = 0;
    for (size_t i = 0; i < maj.size(); ++i) if (!eq(maj[i], g["medium_majorant"][i])) ++badMaj;
    printf("majorant grid mismatches %d / %zu\n", badMaj, maj.size());
    bad += badMaj;
    DeviceMedium grid; std::memset(&grid, 0, sizeof(grid));
    grid.type = MEDIUM_GRID; grid.sigmaA = 0; grid.sigmaS = kNLambda; grid.g = 0.35f;
    for (int k = 0; k < 3; ++k) { grid.bmin[k] = P[3 + k]; grid.bmax[k] = P[6 + k]; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) grid.mInv.m[i][j] = P[9 + 16 + 4 * i + j];
    grid.density = dens.data(); grid.nx = nx; grid.ny = ny; grid.nz = nz;
    grid.majorant = maj.data(); grid.mres[0] = grid.mres[1] = grid.mres[2] = 16;
    DeviceMedium homog; std::memset(&homog, 0, sizeof(homog));
    homog.type = MEDIUM_HOMOGENEOUS; homog.sigmaA = 2 * kNLambda; homog.sigmaS = 3 * kNLambda; homog.g = -0.2f;
    const int kEvents = 6, rec = 1 + 12 * kEvents + 5;
    for (int which = 0; which < 2; ++which) {
        const std::vector<float> &in = g[which == 0 ? "medium_grid_in" : "medium_homog_in"], &out = g[which == 0 ? "medium_grid_out" : "medium_homog_out"];
        int n = (int)in.size() / 14, nbad = 0, nEvents = 0;
        for (int i = 0; i < n; ++i) {
            const float *a = &in[14 * i], *o = &out[(size_t)rec * i];
            Pt3 ro(a[0], a[1], a[2]); Vec3 rd(a[3], a[4], a[5]);
            Wavelengths wl; for (int k = 0; k < 4; ++k) { wl.lambda[k] = a[10 + k]; wl.pdf[k] = 1; }
            RNG rng((uint64_t)a[8], (uint64_t)a[9]);
            std::vector<float> ev; int nEv = 0;
            Spec T = SampleT_maj(which == 0 ? grid : homog, pool.data(), ro, rd, a[6], a[7], rng, wl,
                [&](Pt3 p, const MediumProps &mp, Spec sm, Spec Tm) -> bool {
                    float r[12] = {p.x, p.y, p.z, mp.sigma_a[0], mp.sigma_a[3], mp.sigma_s[0], mp.sigma_s[2], sm[0], sm[1], Tm[0], Tm[1], Tm[3]};
                    ev.insert(ev.end(), r, r + 12);
                    return ++nEv < kEvents;
                });
            ev.resize(12 * kEvents, 0.f);
            bool ok = nEv == (int)o[0];
            for (int k = 0; ok && k < 12 * kEvents; ++k) ok = eq(ev[k], o[1 + k]);
            for (int k = 0; ok && k < 4; ++k) ok = eq(T[k], o[1 + 12 * kEvents + k]);
            if (ok) ok = eq(rng.UniformFloat(), o[1 + 12 * kEvents + 4]);   // same number of RNG draws
            nEvents += nEv;
            if (!ok && nbad++ < 3) printf("%s ray %d: events %d vs %g; first p %g %g %g vs %g %g %g; T %g vs %g\n", which ? "homog" : "grid", i, nEv, o[0], ev[0], ev[1], ev[2], o[1], o[2], o[3], T[0], o[1 + 12 * kEvents]);
        }
        printf("%s SampleT_maj mismatches %d / %d rays (%d events)\n", which ? "homogeneous" : "grid", nbad, n, nEvents);
        bad += nbad;
    }
    const std::vector<float> &hin = g["hg_in"], &hout = g["hg_out"];
    int nh = (int)hin.size() / 9, badh = 0;
    for (int i = 0; i < nh; ++i) {
        const float *a = &hin[9 * i], *o = &hout[6 * i];
        float pdf; Vec3 wi = SampleHenyeyGreenstein(Vec3(a[0], a[1], a[2]), a[8], Vec2(a[6], a[7]), &pdf);
        float p2 = PhaseHG_p(a[8], Vec3(a[0], a[1], a[2]), Vec3(a[3], a[4], a[5]));
        if (!(eq(wi.x, o[0]) && eq(wi.y, o[1]) && eq(wi.z, o[2]) && eq(pdf, o[3]) && eq(pdf, o[4]) && eq(p2, o[5]))) { if (badh++ < 3) printf("hg %d: %g %g %g pdf %g vs %g %g %g pdf %g\n", i, wi.x, wi.y, wi.z, pdf, o[0], o[1], o[2], o[3]); }
    }
    printf("HG mismatches %d / %d\n", badh, nh);
    bad += badh;
    return bad ? 2 : 0;
}
