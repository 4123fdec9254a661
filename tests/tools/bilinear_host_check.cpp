// Test tool: csrc/bilinear.h (BilinearPatch: IsRectangle, area, bounds, NormalBounds, IntersectBilinearPatch + InteractionFromIntersection
// incl. (s,t) re-parameterisation and interpolated shading normals, Sample(ctx, u) with spherical-rectangle sampling, PDF(ctx, wi))
// compiled for the CPU against golden vectors written by the reference itself (oracle/ref_dump.cpp "blp_*": 16 patches -- rectangles,
// planar quads, twisted patches, patches collapsed to triangles, with and without uv / N, mirrored transforms -- x 300 rays x 96
// sampling queries).  Every float is compared bit for bit.
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/bilinear_host_check.cpp -o /tmp/blpcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "bilinear.h"
#include "xform.h"
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
    if (!g.count("blp_patches")) { printf("golden vectors lack the bilinear patch section\n"); return 3; }
    const std::vector<float> &S = g["blp_patches"], &R = g["blp_rays"], &H = g["blp_hits"], &SI = g["blp_samp_in"], &SO = g["blp_samp_out"];
    const int per = 5 + 4 * 8 + 10;
    const int nP = (int)S.size() / per;
    const int nRays = (int)R.size() / 7 / nP, nSamp = (int)SI.size() / 11 / nP;
    int bad = 0, nHit = 0, nChecked = 0, nSampValid = 0, nPdfPos = 0, nRect = 0;
    for (int k = 0; k < nP; ++k) {
        const float *s = &S[(size_t)k * per];
        PatchData pd;
        pd.hasUV = s[0] != 0; pd.hasN = s[1] != 0; pd.flipNormal = s[2] != 0;
        Pt3 P[4]; Vec2 UV[4]; Nrm3 N[4];
        for (int i = 0; i < 4; ++i) {
            const float *v = s + 5 + 8 * i;
            P[i] = Pt3(v[0], v[1], v[2]); UV[i] = Vec2(v[3], v[4]); N[i] = Nrm3(v[5], v[6], v[7]);
        }
        pd.p00 = P[0]; pd.p10 = P[1]; pd.p01 = P[2]; pd.p11 = P[3];
        pd.uv00 = UV[0]; pd.uv10 = UV[1]; pd.uv01 = UV[2]; pd.uv11 = UV[3];
        pd.n00 = N[0]; pd.n10 = N[1]; pd.n01 = N[2]; pd.n11 = N[3];
        pd.isRectangle = PatchIsRectangle(pd.p00, pd.p10, pd.p01, pd.p11);
        pd.area = PatchArea(pd.p00, pd.p10, pd.p01, pd.p11);
        nRect += pd.isRectangle;
        const float *ex = s + 5 + 32;
        Bounds3 bb = Union(Bounds3(pd.p00, pd.p01), Bounds3(pd.p10, pd.p11));
        Vec3 w; float cosTheta;
        PatchNormalBounds(pd, &w, &cosTheta);
        bool okB = (pd.isRectangle ? 1.f : 0.f) == s[3] && eq(pd.area, s[4]) && eq(bb.pMin.x, ex[0]) && eq(bb.pMin.y, ex[1]) && eq(bb.pMin.z, ex[2]) && eq(bb.pMax.x, ex[3]) &&
                   eq(bb.pMax.y, ex[4]) && eq(bb.pMax.z, ex[5]) && eq(w.x, ex[6]) && eq(w.y, ex[7]) && eq(w.z, ex[8]) && eq(cosTheta, ex[9]);
        if (!okB) { printf("patch %d: rectangle %d/%g area %.9g/%.9g bounds / normal cone (%.9g %.9g %.9g | %.9g vs %.9g %.9g %.9g | %.9g) differ\n", k, (int)pd.isRectangle, s[3], pd.area, s[4],
                           w.x, w.y, w.z, cosTheta, ex[6], ex[7], ex[8], ex[9]); ++bad; }
        for (int r = 0; r < nRays; ++r) {
            const float *ry = &R[((size_t)k * nRays + r) * 7];
            const float *o = &H[((size_t)k * nRays + r) * 40];
            Pt3 ro(ry[0], ry[1], ry[2]); Vec3 rd(ry[3], ry[4], ry[5]);
            float uu, vv, t;
            bool hit = IntersectBilinearPatch(ro, rd, ry[6], pd.p00, pd.p10, pd.p01, pd.p11, &uu, &vv, &t);
            ++nChecked;
            bool ok = hit == (o[0] != 0) && hit == (o[39] != 0);
            if (ok && hit) {
                ++nHit;
                SurfaceExtras e;
                SurfaceIntr si = BilinearInteraction(pd, uu, vv, &e);
                const float v[] = {t, si.pi.x.low, si.pi.x.high, si.pi.y.low, si.pi.y.high, si.pi.z.low, si.pi.z.high, si.n.x, si.n.y, si.n.z, si.ns.x, si.ns.y, si.ns.z,
                                   si.uv.x, si.uv.y, e.dpdu.x, e.dpdu.y, e.dpdu.z, e.dpdv.x, e.dpdv.y, e.dpdv.z, e.dndus.x, e.dndus.y, e.dndus.z, e.dndvs.x, e.dndvs.y, e.dndvs.z,
                                   si.dpdus.x, si.dpdus.y, si.dpdus.z, e.dpdvs.x, e.dpdvs.y, e.dpdvs.z};
                for (size_t q = 0; q < sizeof(v) / sizeof(v[0]); ++q)
                    if (!eq(v[q], o[1 + q])) { ok = false; if (bad < 12) printf("patch %d ray %d: member %zu %.9g vs %.9g\n", k, r, q, v[q], o[1 + q]); break; }
            } else if (!ok && bad < 12) printf("patch %d ray %d: hit %d vs %d\n", k, r, (int)hit, (int)o[0]);
            if (!ok) ++bad;
        }
        for (int r = 0; r < nSamp; ++r) {
            const float *in = &SI[((size_t)k * nSamp + r) * 11];
            const float *o = &SO[((size_t)k * nSamp + r) * 20];
            Pt3i pi; pi.x.low = in[0]; pi.x.high = in[1]; pi.y.low = in[2]; pi.y.high = in[3]; pi.z.low = in[4]; pi.z.high = in[5];
            Nrm3 n(in[6], in[7], in[8]);
            ShapeSample ss = PatchSampleSolidAngle(pd, pi.p(), n, Vec2(in[9], in[10]));
            bool ok = ss.valid == (o[0] != 0);
            if (ok && ss.valid) {
                ++nSampValid;
                const float v[] = {ss.pi.x.low, ss.pi.x.high, ss.pi.y.low, ss.pi.y.high, ss.pi.z.low, ss.pi.z.high, ss.n.x, ss.n.y, ss.n.z, ss.uv.x, ss.uv.y, ss.pdf};
                for (size_t q = 0; q < 12; ++q)
                    if (!eq(v[q], o[1 + q])) { ok = false; if (bad < 12) printf("patch %d sample %d: member %zu %.9g vs %.9g\n", k, r, q, v[q], o[1 + q]); break; }
            } else if (!ok && bad < 12) printf("patch %d sample %d: valid %d vs %d\n", k, r, (int)ss.valid, (int)o[0]);
            float pdf = PatchPDFSolidAngle(pd, pi, n, n, Vec3(o[13], o[14], o[15]));
            if (pdf > 0) ++nPdfPos;
            if (!eq(pdf, o[16])) { ok = false; if (bad < 12) printf("patch %d sample %d: PDF %.9g vs %.9g\n", k, r, pdf, o[16]); }
            if (!ok) ++bad;
        }
    }
    printf("bilinear patches: %d patches (%d rectangles), %d rays (%d hits), %d sampling queries (%d valid, %d positive PDFs), %d mismatches\n", nP, nRect, nChecked, nHit,
           nP * nSamp, nSampValid, nPdfPos, bad);
    return bad ? 1 : 0;
}
