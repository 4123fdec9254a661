// Test tool: the Sobol' and Halton samplers of the product (csrc/sampling.h: SobolIntervalToIndex, SobolSample with every randomiser,
// GetPixel2D, RadicalInverse / ScrambledRadicalInverse / OwenScrambledRadicalInverse, the digit-permutation tables and the
// pixel-to-index constants) compiled for the CPU against golden vectors written by the reference itself (oracle/ref_dump.cpp "smp_*").
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/sampler_host_check.cpp -o /tmp/smpcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "sampling.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b); }
static std::map<std::string, std::vector<uint32_t>> readBlob(const char *path) {
    std::map<std::string, std::vector<uint32_t>> g;
    FILE *f = fopen(path, "rb"); char magic[8];
    if (!f || fread(magic, 1, 8, f) != 8) return g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; if (fread(&dt, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) break;
        std::vector<uint32_t> v(n); if (fread(v.data(), 4, n, f) != n) break;
        g[nm] = v; }
    fclose(f);
    return g;
}
int main() {
    auto g = readBlob("tests/golden/ref_vectors.bin");
    auto tb = readBlob("pbrt-v4_b200/data/pbrt_tables.bin");
    if (!g.count("smp_in") || !tb.count("sobol32_full") || !tb.count("vdc_sobol")) { printf("missing golden / tables\n"); return 3; }
    auto F = [](const std::vector<uint32_t> &v) { std::vector<float> o(v.size()); memcpy(o.data(), v.data(), 4 * v.size()); return o; };
    std::vector<float> in = F(g["smp_in"]), out = F(g["smp_out"]);
    auto pairs = [](const std::vector<uint32_t> &v) { std::vector<uint64_t> o(v.size() / 2); for (size_t i = 0; i < o.size(); ++i) o[i] = (uint64_t)v[2 * i] | ((uint64_t)v[2 * i + 1] << 32); return o; };
    std::vector<uint64_t> vdc = pairs(tb["vdc_sobol"]), vdcInv = pairs(tb["vdc_sobol_inv"]);
    struct Cfg { int kind, spp, seed, r; };
    const Cfg cfgs[7] = {{0, 8, 0, RANDOMIZE_FAST_OWEN}, {0, 5, 3, RANDOMIZE_NONE}, {0, 4, 1, RANDOMIZE_PERMUTE_DIGITS}, {0, 16, 2, RANDOMIZE_OWEN},
                         {1, 8, 0, RANDOMIZE_PERMUTE_DIGITS}, {1, 5, 7, RANDOMIZE_NONE}, {1, 4, 2, RANDOMIZE_OWEN}};
    const int xres = 200, yres = 150, NO = 21;
    std::vector<int> primes = FirstPrimes(1000);
    int n = (int)in.size() / 4, bad[7] = {0}, cnt[7] = {0};
    std::vector<uint16_t> perms; std::vector<uint32_t> offsets; int permSeed = -1;
    for (int q = 0; q < n; ++q) {
        const float *a = &in[4 * q], *o = &out[(size_t)NO * q];
        const Cfg &c = cfgs[(int)a[0]];
        SamplerParams sp; memset(&sp, 0, sizeof(sp));
        sp.type = c.kind == 0 ? 4 : 5; sp.randomize = c.r; sp.seed = c.seed; sp.samplesPerPixel = c.spp;
        sp.sobolFull = tb["sobol32_full"].data(); sp.vdc = vdc.data(); sp.vdcInv = vdcInv.data();
        int scale = 1; while (scale < std::max(xres, yres)) scale <<= 1;
        sp.scale = scale; sp.log2Scale = 0; while ((1 << sp.log2Scale) < scale) ++sp.log2Scale;
        if (c.kind == 1) {
            if (c.r == RANDOMIZE_PERMUTE_DIGITS && permSeed != c.seed) { BuildHaltonPermutations(primes, c.seed, &perms, &offsets); permSeed = c.seed; }
            sp.primes = primes.data(); sp.haltonPerms = perms.data(); sp.haltonPermOffset = offsets.data();
            HaltonPixelConstants(xres, yres, &sp);
        }
        Sampler s;
        std::vector<float> m;
        s.StartPixelSample(&sp, (int)a[1], (int)a[2], (int)a[3], 0);
        m.push_back(s.Get1D());
        Vec2 px = s.GetPixel2D(); m.push_back(px.x); m.push_back(px.y);
        m.push_back(s.Get1D());
        Vec2 l = s.Get2D(); m.push_back(l.x); m.push_back(l.y);
        for (int k = 0; k < 3; ++k) { m.push_back(s.Get1D()); Vec2 u = s.Get2D(); m.push_back(u.x); m.push_back(u.y); }
        s.StartPixelSample(&sp, (int)a[1], (int)a[2], (int)a[3], 997);
        for (int k = 0; k < 2; ++k) { m.push_back(s.Get1D()); Vec2 u = s.Get2D(); m.push_back(u.x); m.push_back(u.y); }
        bool ok = true;
        for (int k = 0; k < NO; ++k) ok = ok && eq(m[k], o[k]);
        ++cnt[(int)a[0]];
        if (!ok && bad[(int)a[0]]++ < 3) {
            printf("config %d pixel %d %d sample %d:", (int)a[0], (int)a[1], (int)a[2], (int)a[3]);
            for (int k = 0; k < NO; ++k) if (!eq(m[k], o[k])) printf(" [%d] %.9g vs %.9g", k, m[k], o[k]);
            printf("\n");
        }
    }
    int total = 0;
    // ---- pmj02bn
    if (!g.count("pmj_in") || !tb.count("pmj02bn_samples") || !tb.count("blue_noise_u16x2")) { printf("missing pmj golden / tables\n"); return 3; }
    {
        std::vector<float> pin = F(g["pmj_in"]), pout = F(g["pmj_out"]);
        const std::vector<uint32_t> &pm = tb["pmj02bn_samples"], &bn = tb["blue_noise_u16x2"];
        std::vector<uint16_t> bn16(2 * bn.size());
        for (size_t i = 0; i < bn.size(); ++i) { bn16[2 * i] = (uint16_t)(bn[i] & 0xffffu); bn16[2 * i + 1] = (uint16_t)(bn[i] >> 16); }
        const int spps[3] = {16, 4, 6}, seeds[3] = {0, 5, 1}, PO = 21;
        int pbad = 0, pn = (int)pin.size() / 4;
        std::vector<float> pix; int curCfg = -1, tile = 0;
        for (int q = 0; q < pn; ++q) {
            const float *a = &pin[4 * q], *o = &pout[(size_t)PO * q];
            int c = (int)a[0];
            SamplerParams sp; memset(&sp, 0, sizeof(sp));
            sp.type = 6; sp.seed = seeds[c]; sp.samplesPerPixel = spps[c];
            if (c != curCfg) { if (!BuildPmjPixelSamples(pm.data(), spps[c], &tile, &pix)) { printf("pixel sample table failed\n"); return 2; } curCfg = c; }
            sp.pmjSamples = pm.data(); sp.blueNoise = bn16.data(); sp.pmjPixelSamples = pix.data(); sp.pixelTileSize = tile;
            Sampler s; std::vector<float> m;
            s.StartPixelSample(&sp, (int)a[1], (int)a[2], (int)a[3], 0);
            m.push_back(s.Get1D());
            Vec2 px = s.GetPixel2D(); m.push_back(px.x); m.push_back(px.y);
            m.push_back(s.Get1D());
            Vec2 l = s.Get2D(); m.push_back(l.x); m.push_back(l.y);
            for (int k = 0; k < 5; ++k) { m.push_back(s.Get1D()); Vec2 u = s.Get2D(); m.push_back(u.x); m.push_back(u.y); }
            bool ok = true;
            for (int k = 0; k < PO; ++k) ok = ok && eq(m[k], o[k]);
            if (!ok && pbad++ < 3) { printf("pmj02bn config %d pixel %d %d sample %d:", c, (int)a[1], (int)a[2], (int)a[3]); for (int k = 0; k < PO; ++k) if (!eq(m[k], o[k])) printf(" [%d] %.9g vs %.9g", k, m[k], o[k]); printf("\n"); }
        }
        printf("pmj02bn (spp 16 / 4 / 6): %d mismatches of %d\n", pbad, pn);
        total += pbad;
    }
    for (int c = 0; c < 7; ++c) { printf("%s spp %d seed %d randomize %d: %d mismatches of %d\n", cfgs[c].kind ? "halton" : "sobol", cfgs[c].spp, cfgs[c].seed, cfgs[c].r, bad[c], cnt[c]); total += bad[c]; }
    return total ? 2 : 0;
}
