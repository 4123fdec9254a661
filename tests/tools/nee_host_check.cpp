// Test tool: many-light next-event estimation of the product (csrc/lightbvh.h: BVHLightSampler construction; csrc/lights.h:
// light-BVH Sample / PMF, DiffuseAreaLight SampleLi / PDF_Li; csrc/shapes.h: spherical-triangle sampling) compiled for the CPU
// against golden vectors written by the reference itself (oracle/ref_dump.cpp "nee_*": 200 one-triangle area lights, 2000
// queries with surface and medium-like contexts).
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc -Iinclude tests/tools/nee_host_check.cpp -o /tmp/neecheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/b200pt.h"
#include "lights.h"
#include "lightbvh.h"
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
    auto t = readBlob("pbrt-v4_b200/data/pbrt_tables.bin");
    const std::vector<float> &verts = g["nee_verts"], &lp = g["nee_lights"], &in = g["nee_in"], &out = g["nee_out"];
    const std::vector<float> &illum = t["srgb_illuminant"];
    if (verts.empty() || illum.size() != (size_t)kNLambda) { printf("missing golden / tables\n"); return 3; }
    const int nL = (int)lp.size() / 2;
    // scene records exactly as the engine's commit lays them out (engine.cu: triangle pool, one mesh without normals / uvs,
    // one DiffuseAreaLight per triangle, LightBounds per light: DiffuseAreaLight::Bounds + Triangle::NormalBounds)
    std::vector<float4> tri(3 * (size_t)nL);
    std::vector<DeviceLight> lights(nL);
    std::vector<LightBoundsH> lbounds;
    std::vector<float> specPool(illum);
    float LemitMax = 0; for (float v : illum) LemitMax = stdmax(LemitMax, v);
    DeviceMesh mesh; std::memset(&mesh, 0, sizeof(mesh));
    mesh.material = 0; mesh.areaLightBase = 0; mesh.flags = 0; mesh.mediumInside = mesh.mediumOutside = -1;
    for (int i = 0; i < nL; ++i) {
        Pt3 p0(verts[9 * i], verts[9 * i + 1], verts[9 * i + 2]), p1(verts[9 * i + 3], verts[9 * i + 4], verts[9 * i + 5]), p2(verts[9 * i + 6], verts[9 * i + 7], verts[9 * i + 8]);
        tri[3 * i] = make_float4(p0.x, p0.y, p0.z, b2f(0u));
        tri[3 * i + 1] = make_float4(p1.x, p1.y, p1.z, b2f((uint32_t)i));
        tri[3 * i + 2] = make_float4(p2.x, p2.y, p2.z, b2f(0x100u));
        DeviceLight dl; std::memset(&dl, 0, sizeof(dl));
        dl.type = LIGHT_DIFFUSE_AREA; dl.tri = i; dl.scale = lp[2 * i]; dl.twoSided = lp[2 * i + 1] != 0; dl.LemitTable = 0;
        lights[i] = dl;
        float area = 0.5f * Length(Cross(p1 - p0, p2 - p0));
        float phi = LemitMax;
        phi *= dl.scale * area * kPi;
        Nrm3 n = Normalize(Nrm3(Cross(p1 - p0, p2 - p0)));
        Bounds3 b = Union(Bounds3(p0, p1), p2);
        lbounds.push_back(LightBoundsH(b, Normalize(V(n)), phi, 1.f, std::cos(kPi / 2), dl.twoSided != 0));
    }
    LightBVHBuilder lb;
    lb.Build(lbounds);
    SceneView sv; std::memset(&sv, 0, sizeof(sv));
    sv.triPool = tri.data(); sv.meshes = &mesh; sv.lights = lights.data(); sv.specPool = specPool.data(); sv.nLights = nL;
    sv.ls.type = BPT_LIGHTSAMPLER_BVH; sv.ls.nLights = nL; sv.ls.nInfinite = (int)lb.infiniteLights.size(); sv.ls.nNodes = (int)lb.nodes.size();
    sv.ls.nodes = lb.nodes.data(); sv.ls.infiniteLights = lb.infiniteLights.data(); sv.ls.bitTrails = lb.bitTrails.data(); sv.ls.lightInBVH = lb.lightInBVH.data();
    sv.allbMin = lb.allLightBounds.pMin; sv.allbMax = lb.allLightBounds.pMax;
    sv.illumD65 = illum.data();
    const int nQ = (int)in.size() / 14;
    int badPick = 0, badPmf = 0, badOther = 0, badLi = 0, badPdf = 0, nLi = 0;
    for (int q = 0; q < nQ; ++q) {
        const float *a = &in[14 * q], *o = &out[20 * q];
        LightSampleContext ctx;
        ctx.pi = Pt3i(Pt3(a[0], a[1], a[2])); ctx.n = Nrm3(a[3], a[4], a[5]); ctx.ns = ctx.n;
        Wavelengths wl; for (int k = 0; k < 4; ++k) { wl.lambda[k] = a[9 + k]; wl.pdf[k] = 1; }
        float pmf = 0;
        int li = LightSamplerSample(sv, ctx, a[6], &pmf);
        if (li != (int)o[0]) { if (badPick++ < 3) printf("query %d: picked %d, reference %g\n", q, li, o[0]); continue; }
        if (li < 0) continue;
        if (!eq(pmf, o[1]) || !eq(LightSamplerPMF(sv, ctx, li), o[2])) { if (badPmf++ < 3) printf("query %d: pmf %g / %g vs %g / %g\n", q, pmf, LightSamplerPMF(sv, ctx, li), o[1], o[2]); }
        if (!eq(LightSamplerPMF(sv, ctx, (int)a[13]), o[3])) ++badOther;
        LightLiSample ls = AreaLightSampleLi<false>(sv, lights[li], ctx, Vec2(a[7], a[8]), wl);
        bool refValid = o[4] == 1;
        if (ls.valid != refValid) { if (badLi++ < 3) printf("query %d: SampleLi valid %d vs %d\n", q, (int)ls.valid, (int)refValid); continue; }
        if (!refValid) continue;
        ++nLi;
        Pt3 pl = ls.pLight.p();
        bool ok = eq(ls.L[0], o[5]) && eq(ls.L[1], o[6]) && eq(ls.L[3], o[7]) && eq(ls.wi.x, o[8]) && eq(ls.wi.y, o[9]) && eq(ls.wi.z, o[10]) && eq(ls.pdf, o[11]) &&
                  eq(pl.x, o[12]) && eq(pl.y, o[13]) && eq(pl.z, o[14]) && eq(ls.nLight.x, o[15]) && eq(ls.nLight.y, o[16]) && eq(ls.nLight.z, o[17]);
        if (!ok) { if (badLi++ < 3) printf("query %d: SampleLi L %g wi %g %g %g pdf %g p %g %g %g  vs  L %g wi %g %g %g pdf %g p %g %g %g\n", q, ls.L[0], ls.wi.x, ls.wi.y, ls.wi.z, ls.pdf, pl.x, pl.y, pl.z, o[5], o[8], o[9], o[10], o[11], o[12], o[13], o[14]); }
        float pdfLi = AreaLightPDF_Li<false>(sv, lights[li], ctx, ls.wi);
        if (!eq(pdfLi, o[18])) { if (badPdf++ < 3) printf("query %d: PDF_Li %g vs %g\n", q, pdfLi, o[18]); }
    }
    printf("light BVH: %zu nodes for %d lights\n", lb.nodes.size(), nL);
    printf("pick mismatches %d / %d, pmf %d, PMF(other) %d, SampleLi %d / %d, PDF_Li %d\n", badPick, nQ, badPmf, badOther, badLi, nLi, badPdf);
    return (badPick || badPmf || badOther || badLi || badPdf) ? 2 : 0;
}
