// Test tool: csrc/quadrics.h (Sphere / Disk / Cylinder: constructors, bounds, BasicIntersect + InteractionFromIntersection with the
// reference's interval arithmetic, Sample(ctx, u), PDF(ctx, wi)) compiled for the CPU against golden vectors written by the reference
// itself (oracle/ref_dump.cpp "quad_*": 18 shapes x 400 rays x 96 sampling queries, identity / rigid / scaled + mirrored transforms,
// clipped zmin / zmax / phimax, reference points inside and outside).  Every float is compared bit for bit.
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/quadric_host_check.cpp -o /tmp/quadcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "quadrics.h"
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
    if (!g.count("quad_shapes")) { printf("golden vectors lack the quadric section\n"); return 3; }
    const std::vector<float> &S = g["quad_shapes"], &R = g["quad_rays"], &H = g["quad_hits"], &SI = g["quad_samp_in"], &SO = g["quad_samp_out"];
    const int perShape = 8 + 32 + 12;
    const int nShapes = (int)S.size() / perShape;
    const int nRays = (int)R.size() / 7 / nShapes, nSamp = (int)SI.size() / 11 / nShapes;
    int bad = 0, nHit = 0, nChecked = 0, nSampValid = 0, nPdfPos = 0;
    for (int k = 0; k < nShapes; ++k) {
        const float *s = &S[(size_t)k * perShape];
        Mat4 rfo, ofr;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { rfo.m[i][j] = s[8 + 4 * i + j]; ofr.m[i][j] = s[24 + 4 * i + j]; }
        M4 m; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = rfo.m[i][j];
        Xform X(m);
        const float *ex = s + 40;
        if ((X.SwapsHandedness() ? 1.f : 0.f) != ex[11]) { printf("shape %d: SwapsHandedness differs\n", k); ++bad; }
        Pt3 omin, omax;
        DeviceShape sh = MakeQuadric((int)s[0], s[1], s[2], s[3], s[4], s[5], s[6], s[7] != 0, ex[11] != 0, rfo, ofr, &omin, &omax);
        Bounds3 bb;
        for (int c = 0; c < 8; ++c) bb = Union(bb, XformPoint(rfo, Pt3((c & 1) ? omax.x : omin.x, (c & 2) ? omax.y : omin.y, (c & 4) ? omax.z : omin.z)));
        bool okB = eq(bb.pMin.x, ex[0]) && eq(bb.pMin.y, ex[1]) && eq(bb.pMin.z, ex[2]) && eq(bb.pMax.x, ex[3]) && eq(bb.pMax.y, ex[4]) && eq(bb.pMax.z, ex[5]) &&
                   eq(QuadricArea(sh), ex[6]);
        if (sh.type == SHAPE_DISK) {
            Nrm3 n = XformNormal(ofr, Nrm3(0, 0, 1));
            if (sh.reverseOrientation) n = -n;
            Vec3 w = Normalize(V(n));
            okB = okB && eq(w.x, ex[7]) && eq(w.y, ex[8]) && eq(w.z, ex[9]) && ex[10] == 1.f;
        } else okB = okB && ex[10] == -1.f;
        if (!okB) { printf("shape %d (type %d): bounds / area / normal bounds differ\n", k, sh.type); ++bad; }
        for (int r = 0; r < nRays; ++r) {
            const float *ry = &R[((size_t)k * nRays + r) * 7];
            const float *o = &H[((size_t)k * nRays + r) * 40];
            Pt3 ro(ry[0], ry[1], ry[2]); Vec3 rd(ry[3], ry[4], ry[5]);
            QuadricHit qh;
            bool hit = QuadricBasicIntersect(sh, ro, rd, ry[6], &qh);
            ++nChecked;
            bool ok = hit == (o[0] != 0) && hit == (o[39] != 0);
            if (ok && hit) {
                ++nHit;
                SurfaceExtras e;
                SurfaceIntr si = QuadricInteraction(sh, qh.pObj, &e);
                Vec3 wo = QuadricWo(sh, rd);
                const float v[] = {qh.t, si.pi.x.low, si.pi.x.high, si.pi.y.low, si.pi.y.high, si.pi.z.low, si.pi.z.high, si.n.x, si.n.y, si.n.z, si.ns.x, si.ns.y, si.ns.z,
                                   si.uv.x, si.uv.y, e.dpdu.x, e.dpdu.y, e.dpdu.z, e.dpdv.x, e.dpdv.y, e.dpdv.z, e.dndus.x, e.dndus.y, e.dndus.z, e.dndvs.x, e.dndvs.y, e.dndvs.z,
                                   wo.x, wo.y, wo.z, si.dpdus.x, si.dpdus.y, si.dpdus.z};
                for (size_t q = 0; q < sizeof(v) / sizeof(v[0]); ++q)
                    if (!eq(v[q], o[1 + q])) { ok = false; if (bad < 12) printf("shape %d (type %d) ray %d: member %zu %.9g vs %.9g\n", k, sh.type, r, q, v[q], o[1 + q]); break; }
            } else if (!ok && bad < 12) printf("shape %d (type %d) ray %d: hit %d vs %d\n", k, sh.type, r, (int)hit, (int)o[0]);
            if (!ok) ++bad;
        }
        for (int r = 0; r < nSamp; ++r) {
            const float *in = &SI[((size_t)k * nSamp + r) * 11];
            const float *o = &SO[((size_t)k * nSamp + r) * 20];
            Pt3i pi; pi.x.low = in[0]; pi.x.high = in[1]; pi.y.low = in[2]; pi.y.high = in[3]; pi.z.low = in[4]; pi.z.high = in[5];
            Nrm3 n(in[6], in[7], in[8]);
            ShapeSample ss = QuadricSampleSolidAngle(sh, pi, n, Vec2(in[9], in[10]));
            bool ok = ss.valid == (o[0] != 0);
            if (ok && ss.valid) {
                ++nSampValid;
                const float v[] = {ss.pi.x.low, ss.pi.x.high, ss.pi.y.low, ss.pi.y.high, ss.pi.z.low, ss.pi.z.high, ss.n.x, ss.n.y, ss.n.z, ss.uv.x, ss.uv.y, ss.pdf};
                for (size_t q = 0; q < 12; ++q)
                    if (!eq(v[q], o[1 + q])) { ok = false; if (bad < 12) printf("shape %d (type %d) sample %d: member %zu %.9g vs %.9g\n", k, sh.type, r, q, v[q], o[1 + q]); break; }
            } else if (!ok && bad < 12) printf("shape %d (type %d) sample %d: valid %d vs %d\n", k, sh.type, r, (int)ss.valid, (int)o[0]);
            float pdf = QuadricPDFSolidAngle(sh, pi, n, Vec3(o[13], o[14], o[15]));
            if (pdf > 0) ++nPdfPos;
            if (!eq(pdf, o[16])) { ok = false; if (bad < 12) printf("shape %d (type %d) sample %d: PDF %.9g vs %.9g\n", k, sh.type, r, pdf, o[16]); }
            if (!ok) ++bad;
        }
    }
    printf("quadrics: %d shapes, %d rays (%d hits), %d sampling queries (%d valid, %d positive PDFs), %d mismatches\n", nShapes, nChecked, nHit, nShapes * nSamp, nSampValid, nPdfPos, bad);
    return bad ? 1 : 0;
}
