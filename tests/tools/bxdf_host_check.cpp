// Diagnostic (tests only): the product's BxDF code compiled for the CPU vs the reference's golden vectors (bxdf_*).
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/bxdf_host_check.cpp -o /tmp/bxdfcheck
#include <cstdio>
#include <map>
#include <string>
#include <vector>
#include "bxdfs.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b) || (a != a && b != b); }
int main() {
    FILE *f = fopen("tests/golden/ref_vectors.bin", "rb"); char magic[8]; fread(magic, 1, 8, f);
    std::map<std::string, std::vector<float>> g;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; fread(&dt, 4, 1, f); fread(&n, 8, 1, f); std::vector<float> v(n); fread(v.data(), 4, n, f); g[nm] = v; }
    fclose(f);
    std::vector<float> &in = g["bxdf_in"], &out = g["bxdf_out"];
    struct Cfg { float ia, ca, thick, g, albedo; int maxDepth, nSamples; };
    Cfg cfgs[4] = {{0.f, 0.f, 0.01f, 0.f, 0.f, 10, 1}, {0.3f, 0.2f, 0.01f, 0.f, 0.f, 10, 1}, {0.25f, 0.3f, 0.05f, 0.3f, 0.6f, 10, 2}, {0.f, 0.4f, 0.02f, -0.2f, 0.f, 6, 1}};
    int n = (int)in.size() / 10;
    int bad[3][4] = {{0}};
    const char *names[3] = {"conductor", "dielectric", "layered"};
    for (int i = 0; i < n; ++i) {
        const float *a = &in[10 * i];
        int c = (int)a[0];
        Vec3 wo(a[1], a[2], a[3]), wi(a[4], a[5], a[6]); float uc = a[7]; Vec2 u(a[8], a[9]);
        Spec eta(0.2f + 0.1f * c), k(3.f + 0.3f * c); eta[1] = 0.9f; k[2] = 2.2f;
        DielectricBxDF top; top.eta = 1.5f; top.mfDistrib = TRDist(cfgs[c].ia, cfgs[c].ia);
        ConductorBxDF bottom; bottom.mfDistrib = TRDist(cfgs[c].ca, cfgs[c].ca * 0.5f); bottom.eta = eta; bottom.k = k;
        CoatedConductorBxDF lay; lay.top = top; lay.bottom = bottom; lay.thickness = stdmax(cfgs[c].thick, FLT_MIN); lay.g = cfgs[c].g;
        lay.albedo = Spec(cfgs[c].albedo); lay.maxDepth = cfgs[c].maxDepth; lay.nSamples = cfgs[c].nSamples; lay.seed = 0;
        DielectricBxDF rough; rough.eta = 1.3f + 0.1f * c; rough.mfDistrib = TRDist(cfgs[c].ca, cfgs[c].ca);
        for (int which = 0; which < 3; ++which) {
            const float *o = &out[(size_t)(3 * i + which) * 18];
            Spec fv; float pdf; BSDFSample bs; int flags;
            if (which == 0) { fv = bottom.f(wo, wi, 0); pdf = bottom.PDF(wo, wi, 0); bs = bottom.Sample_f(wo, uc, u, 0); flags = bottom.Flags(); }
            else if (which == 1) { fv = rough.f(wo, wi, 0); pdf = rough.PDF(wo, wi, 0); bs = rough.Sample_f(wo, uc, u, 0); flags = rough.Flags(); }
            else { fv = lay.f(wo, wi, 0); pdf = lay.PDF(wo, wi, 0); bs = lay.Sample_f(wo, uc, u, 0); flags = lay.Flags(); }
            bool okf = eq(fv[0], o[0]) && eq(fv[1], o[1]) && eq(fv[2], o[2]) && eq(fv[3], o[3]);
            if (!okf) { if (bad[which][0]++ < 2) printf("%s f mismatch i=%d cfg=%d: %g %g vs %g %g\n", names[which], i, c, fv[0], fv[1], o[0], o[1]); }
            if (!eq(pdf, o[4])) { if (bad[which][1]++ < 2) printf("%s pdf mismatch i=%d cfg=%d: %g vs %g\n", names[which], i, c, pdf, o[4]); }
            if (flags != (int)o[5]) bad[which][3]++;
            bool oks = bs.valid == (o[6] == 1);
            if (oks && bs.valid) oks = eq(bs.f[0], o[7]) && eq(bs.f[1], o[8]) && eq(bs.f[2], o[9]) && eq(bs.f[3], o[10]) && eq(bs.wi.x, o[11]) && eq(bs.wi.y, o[12]) && eq(bs.wi.z, o[13]) &&
                                   eq(bs.pdf, o[14]) && bs.flags == (int)o[15] && eq(bs.eta, o[16]) && (bs.pdfIsProportional ? 1.f : 0.f) == o[17];
            if (!oks) { if (bad[which][2]++ < 2) printf("%s sample mismatch i=%d cfg=%d: valid %d/%g f0 %g/%g wi %g %g %g / %g %g %g pdf %g/%g flags %d/%g\n", names[which], i, c, bs.valid, o[6], bs.f[0], o[7], bs.wi.x, bs.wi.y, bs.wi.z, o[11], o[12], o[13], bs.pdf, o[14], bs.flags, o[15]); }
        }
    }
    int tot = 0;
    for (int w = 0; w < 3; ++w) { printf("%s: f %d pdf %d sample %d flags %d (of %d)\n", names[w], bad[w][0], bad[w][1], bad[w][2], bad[w][3], n); tot += bad[w][0] + bad[w][1] + bad[w][2] + bad[w][3]; }
    return tot ? 2 : 0;
}
