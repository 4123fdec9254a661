// Test tool: Material::GetBxDF parameter mapping of the product (csrc/materials.cuh: roughness remapping, spectral eta / k
// sampling from piecewise-linear named spectra, reflectance -> k, secondary-wavelength termination by a dispersive interface)
// compiled for the CPU against golden vectors written by the reference itself (oracle/ref_dump.cpp "mat_*").
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/material_host_check.cpp -o /tmp/matcheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "materials.cuh"
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
    const std::vector<float> &in = g["mat_in"], &out = g["mat_out"];
    if (in.empty()) { printf("missing golden\n"); return 3; }
    std::vector<float> pool;
    // named spectra as the engine stores them: [lambdas..., values...] (PiecewiseLinearSpectrum::FromInterleaved, engine.cu MakeSpectrumRec)
    auto named = [&](const char *name) {
        const std::vector<float> &v = t[std::string("pl:") + name];
        SpectrumRec r; std::memset(&r, 0, sizeof(r));
        r.kind = SPEC_PIECEWISE; r.n = (int)v.size() / 2; r.table = (int)pool.size();
        pool.insert(pool.end(), v.begin(), v.end());
        return r;
    };
    auto constant = [](float c) { SpectrumRec r; std::memset(&r, 0, sizeof(r)); r.kind = SPEC_CONSTANT; r.c0 = c; return r; };
    SpectrumRec cuEta = named("metal-Cu-eta"), cuK = named("metal-Cu-k"), bk7 = named("glass-BK7");
    SceneView sv; std::memset(&sv, 0, sizeof(sv));
    auto mk = [&]() { DeviceMaterial m; std::memset(&m, 0, sizeof(m)); return m; };
    DeviceMaterial condSpectral = mk(); condSpectral.eta = cuEta; condSpectral.k = cuK; condSpectral.uroughness = 0.3f; condSpectral.vroughness = 0.05f; condSpectral.flags = 1;
    DeviceMaterial condRefl = mk(); condRefl.reflectance = constant(0.7f); condRefl.useReflectance = 1; condRefl.uroughness = condRefl.vroughness = 0.2f; condRefl.flags = 0;
    DeviceMaterial dielConst = mk(); dielConst.eta = constant(1.5f); dielConst.flags = 1;
    DeviceMaterial dielDisp = mk(); dielDisp.eta = bk7; dielDisp.uroughness = 0.1f; dielDisp.vroughness = 0.25f; dielDisp.flags = 1;
    DeviceMaterial coated = mk(); coated.thickness = 0.01f; coated.iface_eta = constant(1.5f); coated.albedo = constant(0.f); coated.eta = cuEta; coated.k = cuK;
    coated.flags = 1; coated.maxDepth = 10; coated.nSamples = 1;
    DeviceMaterial coated2 = mk(); coated2.iface_uroughness = 0.2f; coated2.iface_vroughness = 0.1f; coated2.thickness = 0.05f; coated2.iface_eta = bk7; coated2.g = 0.3f;
    coated2.albedo = constant(0.7f); coated2.uroughness = 0.3f; coated2.vroughness = 0.15f; coated2.reflectance = constant(0.7f); coated2.useReflectance = 1;
    coated2.flags = 0; coated2.maxDepth = 6; coated2.nSamples = 2;
    const int rec = 8 + 8 + 5 + 5 + 16 + 16;
    int n = (int)in.size(), bad = 0;
    for (int q = 0; q < n; ++q) {
        const float *o = &out[(size_t)rec * q];
        sv.specPool = pool.data();
        std::vector<float> mine;
        bool ch;
        { Wavelengths wl = Wavelengths::SampleVisible(in[q]); ConductorBxDF b = ConductorMaterialEval::GetBxDF<false>(condSpectral, sv, wl, &ch, nullptr);
          mine.insert(mine.end(), {b.mfDistrib.alpha_x, b.mfDistrib.alpha_y, b.eta[0], b.eta[3], b.k[0], b.k[3], wl.pdf[0], wl.pdf[3]}); }
        { Wavelengths wl = Wavelengths::SampleVisible(in[q]); ConductorBxDF b = ConductorMaterialEval::GetBxDF<false>(condRefl, sv, wl, &ch, nullptr);
          mine.insert(mine.end(), {b.mfDistrib.alpha_x, b.mfDistrib.alpha_y, b.eta[0], b.eta[3], b.k[0], b.k[3], wl.pdf[0], wl.pdf[3]}); }
        { Wavelengths wl = Wavelengths::SampleVisible(in[q]); DielectricBxDF b = DielectricMaterialEval::GetBxDF<false>(dielConst, sv, wl, &ch, nullptr);
          mine.insert(mine.end(), {b.mfDistrib.alpha_x, b.mfDistrib.alpha_y, b.eta, wl.pdf[0], wl.pdf[3]}); }
        { Wavelengths wl = Wavelengths::SampleVisible(in[q]); DielectricBxDF b = DielectricMaterialEval::GetBxDF<false>(dielDisp, sv, wl, &ch, nullptr);
          mine.insert(mine.end(), {b.mfDistrib.alpha_x, b.mfDistrib.alpha_y, b.eta, wl.pdf[0], wl.pdf[3]}); }
        for (int c = 0; c < 2; ++c) {
            Wavelengths wl = Wavelengths::SampleVisible(in[q]); CoatedConductorBxDF b = CoatedConductorMaterialEval::GetBxDF<false>(c ? coated2 : coated, sv, wl, &ch, nullptr);
            mine.insert(mine.end(), {b.top.mfDistrib.alpha_x, b.top.mfDistrib.alpha_y, b.top.eta, b.bottom.mfDistrib.alpha_x, b.bottom.mfDistrib.alpha_y, b.bottom.eta[0],
                                     b.bottom.eta[3], b.bottom.k[0], b.bottom.k[3], b.thickness, b.g, b.albedo[0], (float)b.maxDepth, (float)b.nSamples, wl.pdf[0], wl.pdf[3]});
        }
        bool ok = (int)mine.size() == rec;
        for (int k = 0; ok && k < rec; ++k) ok = eq(mine[k], o[k]);
        if (!ok && bad++ < 3) { printf("query %d:", q); for (int k = 0; k < rec && k < (int)mine.size(); ++k) if (!eq(mine[k], o[k])) printf(" [%d] %g vs %g", k, mine[k], o[k]); printf("\n"); }
    }
    printf("material parameter mapping: %d mismatches of %d wavelength samples x 6 materials\n", bad, n);
    return bad ? 2 : 0;
}
