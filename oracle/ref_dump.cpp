// ref_dump.cpp -- TEST / DATA-GENERATION TOOL, linked against the UNMODIFIED reference (oracle/_ref/lib/libpbrt.a).
// Not part of the product; built by oracle/build_ref.py.
//
//   ref_dump tables <out.bin>       standard data tables the engine loads at start-up (CIE matching functions,
//                                   D65 illuminant, sRGB matrices, the sRGB RGB->spectrum coefficient table produced by
//                                   the reference's rgb2spec_opt generator, Sobol' generator matrices, named metal spectra)
//   ref_dump golden <out.bin>       golden vectors of path functions evaluated by the reference itself
//                                   (hash, ZSobol sampler, IntersectTriangle, sampling routines, BVH closest hits ...)
#include <algorithm>
#include <any>
#include <array>
#include <atomic>
#include <condition_variable>
#include <fstream>
#include <functional>
#include <future>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <set>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <variant>
#include <vector>
#include <memory_resource>
#include <cmath>
#include <limits>
#include <numeric>
#include <random>
#include <chrono>
#include <iomanip>
#include <typeindex>
#include <type_traits>
#include <utility>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#define private public
#define protected public
#include <pbrt/pbrt.h>
#include <pbrt/cpu/aggregates.h>
#include <pbrt/cpu/primitive.h>
#include <pbrt/bxdfs.h>
#include <pbrt/lightsamplers.h>
#include <pbrt/cameras.h>
#include <pbrt/film.h>
#include <pbrt/filters.h>
#include <pbrt/paramdict.h>
#include <pbrt/lights.h>
#include <pbrt/materials.h>
#include <pbrt/textures.h>
#include <pbrt/media.h>
#include <pbrt/options.h>
#include <pbrt/samplers.h>
#include <pbrt/shapes.h>
#include <pbrt/util/color.h>
#include <pbrt/util/colorspace.h>
#include <pbrt/util/hash.h>
#include <pbrt/util/lowdiscrepancy.h>
#include <pbrt/util/mesh.h>
#include <pbrt/util/rng.h>
#include <pbrt/util/sampling.h>
#include <pbrt/util/sobolmatrices.h>
#include <pbrt/util/pmj02tables.h>
#include <pbrt/util/bluenoise.h>
#include <pbrt/util/spectrum.h>
#undef private
#undef protected
// the EWA weight table is a file-scope static of the reference's mipmap.cpp (printed there with nine digits, not reproducible from
// its generating formula): compile that source file into this tool, where it lies, to dump the table
#include <pbrt/util/mipmap.cpp>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace pbrt;

namespace pbrt {
extern const float sRGBToSpectrumTable_Scale[64];
extern const RGBToSpectrumTable::CoefficientArray sRGBToSpectrumTable_Data;
}

struct Blob {
    FILE *f;
    explicit Blob(const char *path) { f = fopen(path, "wb"); if (!f) { perror(path); exit(1); } fwrite("BPTTBL1", 1, 8, f); }
    ~Blob() { fclose(f); }
    void put(const char *name, uint32_t dtype, const void *data, uint64_t count) {
        char nm[32]; memset(nm, 0, 32); strncpy(nm, name, 31);
        fwrite(nm, 1, 32, f); fwrite(&dtype, 4, 1, f); fwrite(&count, 8, 1, f); fwrite(data, 4, count, f);
    }
    void f32(const char *name, const std::vector<float> &v) { put(name, 0, v.data(), v.size()); }
    void u32(const char *name, const std::vector<uint32_t> &v) { put(name, 1, v.data(), v.size()); }
    void i32(const char *name, const std::vector<int> &v) { put(name, 2, v.data(), v.size()); }
};

static std::vector<float> Dense(Spectrum s) {
    std::vector<float> v;
    for (int l = 360; l <= 830; ++l) v.push_back(s(l));
    return v;
}

static void DumpTables(const char *path) {
    Blob b(path);
    b.f32("cie_x", Dense(&Spectra::X()));
    b.f32("cie_y", Dense(&Spectra::Y()));
    b.f32("cie_z", Dense(&Spectra::Z()));
    b.f32("cie_y_integral", {CIE_Y_integral});
    const RGBColorSpace *cs = RGBColorSpace::sRGB;
    b.f32("srgb_illuminant", Dense(&cs->illuminant));
    std::vector<float> m;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.push_back(cs->RGBFromXYZ[i][j]);
    b.f32("srgb_rgb_from_xyz", m);
    m.clear();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.push_back(cs->XYZFromRGB[i][j]);
    b.f32("srgb_xyz_from_rgb", m);
    b.f32("srgb_white", {cs->w.x, cs->w.y});
    b.f32("rgb2spec_srgb_scale", std::vector<float>(sRGBToSpectrumTable_Scale, sRGBToSpectrumTable_Scale + 64));
    const float *cd = &sRGBToSpectrumTable_Data[0][0][0][0][0];
    b.f32("rgb2spec_srgb_coeffs", std::vector<float>(cd, cd + 3 * 64 * 64 * 64 * 3));
    b.u32("sobol32_2x52", std::vector<uint32_t>(SobolMatrices32, SobolMatrices32 + 2 * SobolMatrixSize));
    b.u32("sobol32_full", std::vector<uint32_t>(SobolMatrices32, SobolMatrices32 + NSobolDimensions * SobolMatrixSize));
    {   // SobolIntervalToIndex's matrices, 25 rows of 52 64-bit columns each, as (low, high) 32-bit pairs
        std::vector<uint32_t> v, vi;
        for (int r = 0; r < 25; ++r)
            for (int c = 0; c < SobolMatrixSize; ++c) {
                v.push_back((uint32_t)VdCSobolMatrices[r][c]); v.push_back((uint32_t)(VdCSobolMatrices[r][c] >> 32));
                vi.push_back((uint32_t)VdCSobolMatricesInv[r][c]); vi.push_back((uint32_t)(VdCSobolMatricesInv[r][c] >> 32));
            }
        b.u32("vdc_sobol", v);
        b.u32("vdc_sobol_inv", vi);
    }
    {   // PMJ02BNSampler: the five pmj02bn sets (fixed point) and the blue-noise textures (two 16-bit values per word, [texture][x][y] order)
        const uint32_t *pm = &pmj02bnSamples[0][0][0];
        b.u32("pmj02bn_samples", std::vector<uint32_t>(pm, pm + (size_t)nPMJ02bnSets * nPMJ02bnSamples * 2));
        const uint16_t *bn = &BlueNoiseTextures[0][0][0];
        const size_t nb = (size_t)NumBlueNoiseTextures * BlueNoiseResolution * BlueNoiseResolution;
        std::vector<uint32_t> packed(nb / 2);
        for (size_t i = 0; i < nb / 2; ++i) packed[i] = (uint32_t)bn[2 * i] | ((uint32_t)bn[2 * i + 1] << 16);
        b.u32("blue_noise_u16x2", packed);
    }
    b.f32("mip_filter_lut", std::vector<float>(MIPFilterLUT, MIPFilterLUT + MIPFilterLUTSize));
    b.f32("srgb8_to_linear", std::vector<float>(SRGBToLinearLUT, SRGBToLinearLUT + 256));
    // named spectra (piecewise linear): lambdas then values
    const char *names[] = {"metal-Cu-eta", "metal-Cu-k", "metal-Au-eta", "metal-Au-k", "metal-Ag-eta", "metal-Ag-k",
                           "metal-Al-eta", "metal-Al-k", "glass-BK7", "glass-BAF10", "glass-F11", "stdillum-D65",
                           "illum-A", "illum-D50"};
    for (const char *nm : names) {
        Spectrum s = GetNamedSpectrum(nm);
        if (!s) { fprintf(stderr, "named spectrum %s missing\n", nm); continue; }
        if (s.Is<PiecewiseLinearSpectrum>()) {
            const PiecewiseLinearSpectrum *pl = s.Cast<PiecewiseLinearSpectrum>();
            std::vector<float> v(pl->lambdas.begin(), pl->lambdas.end());
            v.insert(v.end(), pl->values.begin(), pl->values.end());
            b.f32((std::string("pl:") + nm).c_str(), v);
        } else {
            b.f32((std::string("dense:") + nm).c_str(), Dense(s));
        }
    }
}

// ------------------------------------------------------------------------------------------------ golden vectors
static void DumpGolden(const char *path) {
    Blob b(path);
    RNG rng(1234);
    // hashes
    {
        std::vector<uint32_t> out;
        for (int i = 0; i < 64; ++i) {
            int a = rng.Uniform<uint32_t>() >> 8, c = rng.Uniform<uint32_t>() & 0xffff;
            uint64_t h = Hash(a, c);
            uint64_t mx = MixBits(((uint64_t)a << 20) ^ c);
            float fa = rng.Uniform<float>(), fb = rng.Uniform<float>(), fc = rng.Uniform<float>();
            uint64_t h3 = Hash(Point3f(fa, fb, fc));
            uint64_t h1 = Hash(fa);
            out.insert(out.end(), {(uint32_t)a, (uint32_t)c, (uint32_t)h, (uint32_t)(h >> 32), (uint32_t)mx, (uint32_t)(mx >> 32),
                                   FloatToBits(fa), FloatToBits(fb), FloatToBits(fc), (uint32_t)h3, (uint32_t)(h3 >> 32),
                                   (uint32_t)h1, (uint32_t)(h1 >> 32)});
        }
        b.u32("hash", out);
    }
    // ZSobol + Independent sampler values: (px, py, sampleIndex, dim) -> Get1D, Get2D
    for (int cfg = 0; cfg < 4; ++cfg) {
        int spp = cfg == 0 ? 64 : (cfg == 1 ? 32 : (cfg == 2 ? 256 : 16));
        Point2i res = cfg == 2 ? Point2i(1024, 1024) : Point2i(256, 200);
        int seed = cfg == 1 ? 7 : 0;
        RandomizeStrategy rs = cfg == 3 ? RandomizeStrategy::Owen : (cfg == 1 ? RandomizeStrategy::PermuteDigits : RandomizeStrategy::FastOwen);
        ZSobolSampler zs(spp, res, rs, seed);
        IndependentSampler is(spp, seed);
        std::vector<int> in;
        std::vector<float> out, outI;
        for (int i = 0; i < 512; ++i) {
            int px = rng.Uniform<uint32_t>() % res.x, py = rng.Uniform<uint32_t>() % res.y;
            int si = rng.Uniform<uint32_t>() % spp, dim = rng.Uniform<uint32_t>() % 60;
            in.insert(in.end(), {px, py, si, dim});
            zs.StartPixelSample({px, py}, si, dim);
            float a = zs.Get1D();
            zs.StartPixelSample({px, py}, si, dim);
            Point2f p = zs.Get2D();
            out.insert(out.end(), {a, p.x, p.y});
            is.StartPixelSample({px, py}, si, dim);
            float ia = is.Get1D();
            is.StartPixelSample({px, py}, si, dim);
            Point2f ip = is.Get2D();
            outI.insert(outI.end(), {ia, ip.x, ip.y});
        }
        b.i32(StringPrintf("sampler%d_params", cfg).c_str(), {spp, res.x, res.y, seed, (int)rs});
        b.i32(StringPrintf("sampler%d_in", cfg).c_str(), in);
        b.f32(StringPrintf("sampler%d_zsobol", cfg).c_str(), out);
        b.f32(StringPrintf("sampler%d_indep", cfg).c_str(), outI);
    }
    // IntersectTriangle: random rays vs random triangles (including edge/vertex grazing cases)
    {
        std::vector<float> in, out;
        for (int i = 0; i < 4096; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f p0(r(), r(), r()), p1(r(), r(), r()), p2(r(), r(), r());
            if (i % 4 == 1) { p1 = p0 + 0.01f * Vector3f(r(), r(), r()); p2 = p0 + 0.01f * Vector3f(r(), r(), r()); }
            Point3f o(3 * r(), 3 * r(), 3 * r());
            Float b0 = rng.Uniform<float>(), b1 = rng.Uniform<float>() * (1 - b0);
            Point3f target = b0 * p0 + b1 * p1 + (1 - b0 - b1) * p2;
            if (i % 8 == 3) target = p0;                           // through a vertex
            if (i % 8 == 5) target = 0.5f * p0 + 0.5f * p1;          // through an edge
            if (i % 3 == 2) target += Vector3f(r(), r(), r()) * 0.7f;  // likely miss
            Vector3f d = target - o;
            if (i % 2) d = Normalize(d);
            Float tMax = (i % 5 == 4) ? 0.9f * Length(target - o) / Length(d) : Infinity;
            pstd::optional<TriangleIntersection> ti = IntersectTriangle(Ray(o, d), tMax, p0, p1, p2);
            in.insert(in.end(), {o.x, o.y, o.z, d.x, d.y, d.z, tMax, p0.x, p0.y, p0.z, p1.x, p1.y, p1.z, p2.x, p2.y, p2.z});
            if (ti) out.insert(out.end(), {1.f, ti->b0, ti->b1, ti->b2, ti->t}); else out.insert(out.end(), {0.f, 0.f, 0.f, 0.f, 0.f});
        }
        b.f32("tri_in", in);
        b.f32("tri_out", out);
    }
    // the reference's own known-answer case, Triangle.BadCases (shapes_test.cpp:435): must miss
    {
        std::vector<float> in = {-1113.45459f, 79.049614f, -56.2431908f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        b.f32("tri_badcase_note", in);
    }
    // sampling routines
    {
        std::vector<float> in, out;
        for (int i = 0; i < 1024; ++i) {
            Point2f u(rng.Uniform<float>(), rng.Uniform<float>());
            Point2f d = SampleUniformDiskConcentric(u);
            Vector3f c = SampleCosineHemisphere(u);
            Float lam = SampleVisibleWavelengths(u.x), lpdf = VisibleWavelengthsPDF(lam);
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            pstd::array<Point3f, 3> v = {Point3f(r(), 2 + r(), r()), Point3f(r(), 2 + r(), r()), Point3f(r(), 2 + r(), r())};
            Point3f p(r(), r() - 1, r());
            Float pdf;
            pstd::array<Float, 3> bc = SampleSphericalTriangle(v, p, u, &pdf);
            Float w[4] = {rng.Uniform<float>(), rng.Uniform<float>(), rng.Uniform<float>(), rng.Uniform<float>()};
            Point2f bl = SampleBilinear(u, w);
            Float blp = BilinearPDF(bl, w);
            Float area = SphericalTriangleArea(Normalize(v[0] - p), Normalize(v[1] - p), Normalize(v[2] - p));
            Vector3f wdir = Normalize((bc[0] * v[0] + bc[1] * v[1] + bc[2] * v[2]) - p);
            Point2f inv = InvertSphericalTriangleSample(v, p, wdir);
            in.insert(in.end(), {u.x, u.y, v[0].x, v[0].y, v[0].z, v[1].x, v[1].y, v[1].z, v[2].x, v[2].y, v[2].z, p.x, p.y, p.z,
                                 w[0], w[1], w[2], w[3]});
            out.insert(out.end(), {d.x, d.y, c.x, c.y, c.z, lam, lpdf, bc[0], bc[1], bc[2], pdf, bl.x, bl.y, blp, area, inv.x, inv.y});
        }
        b.f32("sampling_in", in);
        b.f32("sampling_out", out);
    }
    // FastExp / Gaussian / RGB -> spectrum
    {
        std::vector<float> in, out;
        for (int i = 0; i < 512; ++i) {
            float x = -20 * rng.Uniform<float>();
            RGB rgb(rng.Uniform<float>(), rng.Uniform<float>(), rng.Uniform<float>());
            if (i % 7 == 0) rgb = RGB(rgb.r, rgb.r, rgb.r);
            RGBSigmoidPolynomial sp = RGBColorSpace::sRGB->ToRGBCoeffs(rgb);
            float lam = 360 + 470 * rng.Uniform<float>();
            RGBAlbedoSpectrum as(*RGBColorSpace::sRGB, rgb);
            RGBIlluminantSpectrum il(*RGBColorSpace::sRGB, RGB(17 * rgb.r, 12 * rgb.g, 4 * rgb.b));
            in.insert(in.end(), {x, rgb.r, rgb.g, rgb.b, lam});
            out.insert(out.end(), {FastExp(x), sp.c0, sp.c1, sp.c2, as(lam), il(lam), il.scale});
        }
        b.f32("spectrum_in", in);
        b.f32("spectrum_out", out);
    }
    // BxDFs in the local frame: conductor, dielectric, coated conductor (LayeredBxDF) -- f, Sample_f, PDF
    {
        struct Cfg { float ia, ca, thick, g, albedo; int maxDepth, nSamples; };
        Cfg cfgs[4] = {{0.f, 0.f, 0.01f, 0.f, 0.f, 10, 1}, {0.3f, 0.2f, 0.01f, 0.f, 0.f, 10, 1}, {0.25f, 0.3f, 0.05f, 0.3f, 0.6f, 10, 2}, {0.f, 0.4f, 0.02f, -0.2f, 0.f, 6, 1}};
        std::vector<float> in, out;
        for (int c = 0; c < 4; ++c) {
            SampledSpectrum eta(0.2f + 0.1f * c), k(3.f + 0.3f * c), alb(cfgs[c].albedo);
            eta[1] = 0.9f; k[2] = 2.2f;
            DielectricBxDF top(1.5f, TrowbridgeReitzDistribution(cfgs[c].ia, cfgs[c].ia));
            ConductorBxDF bottom(TrowbridgeReitzDistribution(cfgs[c].ca, cfgs[c].ca * 0.5f), eta, k);
            CoatedConductorBxDF layered(top, bottom, cfgs[c].thick, alb, cfgs[c].g, cfgs[c].maxDepth, cfgs[c].nSamples);
            for (int i = 0; i < 256; ++i) {
                auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
                Vector3f wo = Normalize(Vector3f(r(), r(), r())), wi = Normalize(Vector3f(r(), r(), r()));
                if (i % 3 == 0) wi.z = std::abs(wi.z) * (wo.z > 0 ? 1 : -1);
                Float uc = rng.Uniform<float>(); Point2f u(rng.Uniform<float>(), rng.Uniform<float>());
                in.insert(in.end(), {(float)c, wo.x, wo.y, wo.z, wi.x, wi.y, wi.z, uc, u.x, u.y});
                auto put = [&](SampledSpectrum f, Float pdf, pstd::optional<BSDFSample> bs, int flags) {
                    out.insert(out.end(), {f[0], f[1], f[2], f[3], pdf, (float)flags});
                    if (bs) out.insert(out.end(), {1.f, bs->f[0], bs->f[1], bs->f[2], bs->f[3], bs->wi.x, bs->wi.y, bs->wi.z, bs->pdf, (float)bs->flags, bs->eta, bs->pdfIsProportional ? 1.f : 0.f});
                    else out.insert(out.end(), {0.f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
                };
                put(bottom.f(wo, wi, TransportMode::Radiance), bottom.PDF(wo, wi, TransportMode::Radiance, BxDFReflTransFlags::All),
                    bottom.Sample_f(wo, uc, u, TransportMode::Radiance), (int)bottom.Flags());
                DielectricBxDF rough(1.3f + 0.1f * c, TrowbridgeReitzDistribution(cfgs[c].ca, cfgs[c].ca));
                put(rough.f(wo, wi, TransportMode::Radiance), rough.PDF(wo, wi, TransportMode::Radiance), rough.Sample_f(wo, uc, u, TransportMode::Radiance), (int)rough.Flags());
                put(layered.f(wo, wi, TransportMode::Radiance), layered.PDF(wo, wi, TransportMode::Radiance), layered.Sample_f(wo, uc, u, TransportMode::Radiance), (int)layered.Flags());
            }
        }
        b.f32("bxdf_in", in);
        b.f32("bxdf_out", out);
    }
    // ImageInfiniteLight: Le / SampleLi / PDF_Li on a seeded procedural 32x32 equal-area environment map
    {
        Allocator alloc;
        const int res = 32;
        Image img(PixelFormat::Float, {res, res}, {"R", "G", "B"}, nullptr, alloc);
        std::vector<float> pix;
        for (int y = 0; y < res; ++y)
            for (int x = 0; x < res; ++x)
                for (int c = 0; c < 3; ++c) {
                    float v = (rng.Uniform<float>() < 0.02f ? 200.f : 1.f) * rng.Uniform<float>() * (c + 1);
                    img.SetChannel({x, y}, c, v);
                    pix.push_back(v);
                }
        Transform rfl = RotateX(-90) * RotateZ(25);
        ImageInfiniteLight light(rfl, img, RGBColorSpace::sRGB, 0.37f, "golden", alloc);
        light.Preprocess(Bounds3f(Point3f(-1, -2, -3), Point3f(2, 3, 5)));
        std::vector<float> in, out;
        for (int i = 0; i < 512; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Vector3f d = Normalize(Vector3f(r(), r(), r()));
            Point2f u(rng.Uniform<float>(), rng.Uniform<float>());
            SampledWavelengths lambda = SampledWavelengths::SampleVisible(rng.Uniform<float>());
            SampledSpectrum Le = light.Le(Ray(Point3f(0, 0, 0), d), lambda);
            LightSampleContext ctx(Interaction(Point3f(r(), r(), r()), 0.f, Medium()));
            Float pdf = light.PDF_Li(ctx, d, true);
            pstd::optional<LightLiSample> ls = light.SampleLi(ctx, u, lambda, true);
            Point2f sq = EqualAreaSphereToSquare(d);
            Vector3f back = EqualAreaSquareToSphere(u);
            in.insert(in.end(), {d.x, d.y, d.z, u.x, u.y, lambda[0], lambda[1], lambda[2], lambda[3], ctx.p().x, ctx.p().y, ctx.p().z});
            out.insert(out.end(), {Le[0], Le[1], Le[2], Le[3], pdf, sq.x, sq.y, back.x, back.y, back.z});
            if (ls) out.insert(out.end(), {1.f, ls->L[0], ls->L[1], ls->L[2], ls->L[3], ls->wi.x, ls->wi.y, ls->wi.z, ls->pdf, ls->pLight.p().x, ls->pLight.p().y, ls->pLight.p().z});
            else out.insert(out.end(), {0.f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
        }
        std::vector<float> xf;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) xf.push_back(rfl.GetMatrix()[i][j]);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) xf.push_back(rfl.GetInverseMatrix()[i][j]);
        b.f32("env_pixels", pix);
        b.f32("env_xform", xf);
        b.f32("env_in", in);
        b.f32("env_out", out);
    }
    // BVHAggregate closest hits on a seeded random triangle soup (hit record = primitive, b0, b1, b2, t)
    {
        Allocator alloc;
        const int nTri = 2000;
        std::vector<int> idx; std::vector<Point3f> P;
        for (int i = 0; i < nTri; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f c(r(), r(), r());
            for (int k = 0; k < 3; ++k) { P.push_back(c + 0.15f * Vector3f(r(), r(), r())); idx.push_back(3 * i + k); }
        }
        TriangleMesh *mesh = alloc.new_object<TriangleMesh>(Transform(), false, idx, P, std::vector<Vector3f>(), std::vector<Normal3f>(),
                                                            std::vector<Point2f>(), std::vector<int>(), alloc);
        pstd::vector<Shape> tris = Triangle::CreateTriangles(mesh, alloc);
        std::vector<Primitive> prims;
        for (Shape s : tris) prims.push_back(new SimplePrimitive(s, nullptr));
        BVHAggregate bvh(prims, 4, BVHAggregate::SplitMethod::SAH);
        std::vector<float> verts, rays, hits;
        for (const Point3f &p : mesh->p ? std::vector<Point3f>(mesh->p, mesh->p + 3 * nTri) : P) { verts.push_back(p.x); verts.push_back(p.y); verts.push_back(p.z); }
        for (int i = 0; i < 20000; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f o(1.5f * r(), 1.5f * r(), 1.5f * r());
            Vector3f d = Normalize(Vector3f(r(), r(), r()));
            Float tMax = (i % 4 == 3) ? rng.Uniform<float>() : Infinity;
            pstd::optional<ShapeIntersection> si = bvh.Intersect(Ray(o, d), tMax);
            bool occluded = bvh.IntersectP(Ray(o, d), tMax);
            rays.insert(rays.end(), {o.x, o.y, o.z, tMax, d.x, d.y, d.z, 0.f});
            if (si) {
                // recover the primitive by re-testing (closest t is unique up to ties)
                int prim = -1; Float b0 = 0, b1 = 0, b2 = 0;
                for (int k = 0; k < nTri; ++k) {
                    auto ti = IntersectTriangle(Ray(o, d), tMax, mesh->p[3 * k], mesh->p[3 * k + 1], mesh->p[3 * k + 2]);
                    if (ti && ti->t == si->tHit) { prim = k; b0 = ti->b0; b1 = ti->b1; b2 = ti->b2; break; }
                }
                hits.insert(hits.end(), {(float)prim, b0, b1, b2, si->tHit, occluded ? 1.f : 0.f, si->intr.p().x, si->intr.p().y, si->intr.p().z,
                                         si->intr.n.x, si->intr.n.y, si->intr.n.z});
            } else
                hits.insert(hits.end(), {-1.f, 0, 0, 0, 0, occluded ? 1.f : 0.f, 0, 0, 0, 0, 0, 0});
        }
        b.f32("bvh_verts", verts);
        b.f32("bvh_rays", rays);
        b.f32("bvh_hits", hits);
    }
    // Media: GridMedium / HomogeneousMedium majorant tracking (SampleT_maj) and Henyey-Greenstein sampling.
    // (appended last so that the vectors above keep their values)
    {
        Allocator alloc;
        const int nx = 12, ny = 10, nz = 8;
        std::vector<Float> density(nx * ny * nz);
        for (Float &v : density) { Float r = rng.Uniform<float>(); v = r < 0.3f ? 0.f : 2.5f * rng.Uniform<float>(); }
        Transform rfm = Translate(Vector3f(0.3f, -0.2f, 0.5f)) * RotateY(30) * Scale(1.1f, 0.9f, 1.3f);
        RGBUnboundedSpectrum sa(*RGBColorSpace::sRGB, RGB(0.5f, 1.0f, 2.0f)), ss(*RGBColorSpace::sRGB, RGB(3.f, 2.f, 1.f));
        ConstantSpectrum zero(0.f);
        Bounds3f bounds(Point3f(-1, -0.5f, 0), Point3f(1, 1.5f, 2));
        GridMedium grid(bounds, rfm, &sa, &ss, 1.7f, 0.35f, SampledGrid<Float>(density, nx, ny, nz, alloc), {}, 1.f, 0.f, &zero,
                        SampledGrid<Float>({1.f}, 1, 1, 1, alloc), alloc);
        HomogeneousMedium homog(&sa, &ss, 0.6f, &zero, 1.f, -0.2f, alloc);
        std::vector<float> params = {(float)nx, (float)ny, (float)nz, bounds.pMin.x, bounds.pMin.y, bounds.pMin.z, bounds.pMax.x, bounds.pMax.y, bounds.pMax.z};
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) params.push_back(rfm.GetMatrix()[i][j]);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) params.push_back(rfm.GetInverseMatrix()[i][j]);
        b.f32("medium_params", params);
        b.f32("medium_density", std::vector<float>(density.begin(), density.end()));
        b.f32("medium_sigma_a", std::vector<float>(grid.sigma_a_spec.values.begin(), grid.sigma_a_spec.values.end()));
        b.f32("medium_sigma_s", std::vector<float>(grid.sigma_s_spec.values.begin(), grid.sigma_s_spec.values.end()));
        b.f32("medium_majorant", std::vector<float>(grid.majorantGrid.voxels.begin(), grid.majorantGrid.voxels.end()));
        b.f32("medium_h_sigma_a", std::vector<float>(homog.sigma_a_spec.values.begin(), homog.sigma_a_spec.values.end()));
        b.f32("medium_h_sigma_s", std::vector<float>(homog.sigma_s_spec.values.begin(), homog.sigma_s_spec.values.end()));
        const int kEvents = 6;
        for (int which = 0; which < 2; ++which) {
            std::vector<float> in, out;
            for (int i = 0; i < 768; ++i) {
                auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
                Point3f o(2 * r(), 2 * r(), 2 * r());
                if (i % 3 == 0) o = rfm(Point3f(bounds.Lerp(Point3f(rng.Uniform<float>(), rng.Uniform<float>(), rng.Uniform<float>()))));   // start inside
                Vector3f d = Normalize(Vector3f(r(), r(), r())) * (0.5f + 1.5f * rng.Uniform<float>());
                if (i % 17 == 0) d.y = 0;
                Float tMax = (i % 4 == 3) ? Infinity : 4 * rng.Uniform<float>();
                Float u = rng.Uniform<float>();
                uint64_t s0 = rng.Uniform<uint32_t>() >> 8, s1 = rng.Uniform<uint32_t>() >> 8;   // 24 bits: exactly representable as float
                SampledWavelengths lambda = SampledWavelengths::SampleVisible(rng.Uniform<float>());
                RNG trng(s0, s1);
                std::vector<float> ev;
                int nEv = 0;
                auto cb = [&](Point3f p, MediumProperties mp, SampledSpectrum sigma_maj, SampledSpectrum T_maj) {
                    ev.insert(ev.end(), {p.x, p.y, p.z, mp.sigma_a[0], mp.sigma_a[3], mp.sigma_s[0], mp.sigma_s[2], sigma_maj[0], sigma_maj[1],
                                         T_maj[0], T_maj[1], T_maj[3]});
                    return ++nEv < kEvents;
                };
                Ray ray(o, d, 0.f, which == 0 ? Medium(&grid) : Medium(&homog));
                SampledSpectrum T = which == 0 ? SampleT_maj<GridMedium>(ray, tMax, u, trng, lambda, cb)
                                               : SampleT_maj<HomogeneousMedium>(ray, tMax, u, trng, lambda, cb);
                ev.resize(12 * kEvents, 0.f);
                in.insert(in.end(), {o.x, o.y, o.z, d.x, d.y, d.z, tMax, u, (float)s0, (float)s1, lambda[0], lambda[1], lambda[2], lambda[3]});
                out.push_back((float)nEv);
                out.insert(out.end(), ev.begin(), ev.end());
                out.insert(out.end(), {T[0], T[1], T[2], T[3], trng.Uniform<float>()});
            }
            b.f32(which == 0 ? "medium_grid_in" : "medium_homog_in", in);
            b.f32(which == 0 ? "medium_grid_out" : "medium_homog_out", out);
        }
        std::vector<float> hin, hout;
        const float gs[4] = {-0.7f, 0.f, 0.3f, 0.9f};
        for (int i = 0; i < 512; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Vector3f wo = Normalize(Vector3f(r(), r(), r())), w2 = Normalize(Vector3f(r(), r(), r()));
            Point2f u(rng.Uniform<float>(), rng.Uniform<float>());
            HGPhaseFunction hg(gs[i % 4]);
            auto ps = hg.Sample_p(wo, u);
            hin.insert(hin.end(), {wo.x, wo.y, wo.z, w2.x, w2.y, w2.z, u.x, u.y, gs[i % 4]});
            hout.insert(hout.end(), {ps->wi.x, ps->wi.y, ps->wi.z, ps->pdf, ps->p, hg.p(wo, w2)});
        }
        b.f32("hg_in", hin);
        b.f32("hg_out", hout);
    }
    // Many-light next-event estimation: BVHLightSampler over 200 one-triangle DiffuseAreaLights (Sample / PMF with surface and
    // medium-like contexts) and DiffuseAreaLight::SampleLi / PDF_Li.  (appended last so that the vectors above keep their values)
    {
        Allocator alloc;
        const int nL = 200;
        std::vector<int> idx; std::vector<Point3f> P;
        std::vector<float> verts, lparams;
        for (int i = 0; i < nL; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f c(2 * r(), 2 * r(), 2 * r());
            for (int k = 0; k < 3; ++k) { P.push_back(c + 0.3f * Vector3f(r(), r(), r())); idx.push_back(3 * i + k); }
        }
        TriangleMesh *mesh = alloc.new_object<TriangleMesh>(Transform(), false, idx, P, std::vector<Vector3f>(), std::vector<Normal3f>(),
                                                            std::vector<Point2f>(), std::vector<int>(), alloc);
        for (int i = 0; i < 3 * nL; ++i) { verts.push_back(mesh->p[i].x); verts.push_back(mesh->p[i].y); verts.push_back(mesh->p[i].z); }
        pstd::vector<Shape> tris = Triangle::CreateTriangles(mesh, alloc);
        Spectrum Le = &RGBColorSpace::sRGB->illuminant;
        std::vector<Light> lights;
        std::map<const void *, int> lightIndex;
        for (int i = 0; i < nL; ++i) {
            float scale = 0.5f + 3 * rng.Uniform<float>();
            bool twoSided = (i % 3) == 0;
            lparams.push_back(scale); lparams.push_back(twoSided ? 1.f : 0.f);
            Light l = alloc.new_object<DiffuseAreaLight>(Transform(), MediumInterface(), Le, scale, tris[i], nullptr, Image(alloc), nullptr, twoSided);
            lightIndex[l.ptr()] = i;
            lights.push_back(l);
        }
        BVHLightSampler sampler(lights, alloc);
        std::vector<float> in, out;
        for (int q = 0; q < 2000; ++q) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f p(3 * r(), 3 * r(), 3 * r());
            Normal3f n(0, 0, 0);
            if (q % 4 != 0) n = Normal3f(Normalize(Vector3f(r(), r(), r())));
            LightSampleContext ctx(Point3fi(p), n, n);
            Float u = rng.Uniform<float>();
            Point2f u2(rng.Uniform<float>(), rng.Uniform<float>());
            SampledWavelengths lambda = SampledWavelengths::SampleVisible(rng.Uniform<float>());
            int other = std::min(nL - 1, int(rng.Uniform<float>() * nL));
            pstd::optional<SampledLight> sl = sampler.Sample(ctx, u);
            in.insert(in.end(), {p.x, p.y, p.z, n.x, n.y, n.z, u, u2.x, u2.y, lambda[0], lambda[1], lambda[2], lambda[3], (float)other});
            float rec[20] = {0};
            rec[0] = -1;
            if (sl) {
                int li = lightIndex[sl->light.ptr()];
                rec[0] = (float)li; rec[1] = sl->p; rec[2] = sampler.PMF(ctx, lights[li]);
                pstd::optional<LightLiSample> ls = lights[li].SampleLi(ctx, u2, lambda, true);
                if (ls) {
                    rec[4] = 1; rec[5] = ls->L[0]; rec[6] = ls->L[1]; rec[7] = ls->L[3]; rec[8] = ls->wi.x; rec[9] = ls->wi.y; rec[10] = ls->wi.z;
                    rec[11] = ls->pdf; rec[12] = ls->pLight.p().x; rec[13] = ls->pLight.p().y; rec[14] = ls->pLight.p().z;
                    rec[15] = ls->pLight.n.x; rec[16] = ls->pLight.n.y; rec[17] = ls->pLight.n.z;
                    rec[18] = lights[li].PDF_Li(ctx, ls->wi, true);
                }
            }
            rec[3] = sampler.PMF(ctx, lights[other]);
            out.insert(out.end(), rec, rec + 20);
        }
        b.f32("nee_verts", verts);
        b.f32("nee_lights", lparams);
        b.f32("nee_in", in);
        b.f32("nee_out", out);
    }
    // Camera: GetCameraSample (ZSobol dimensions 0..5 + Gaussian FilterSampler) and PerspectiveCamera::GenerateRayDifferential with
    // and without a lens, transformed to render space and scaled as the integrator does.  (appended last)
    for (int cfg = 0; cfg < 2; ++cfg) {
        Allocator alloc;
        const int xres = 96, yres = 64, spp = 16;
        Filter filt = alloc.new_object<GaussianFilter>(Vector2f(1.5f, 1.5f), 0.5f, alloc);
        PixelSensor *sensor = alloc.new_object<PixelSensor>(RGBColorSpace::sRGB, nullptr, 1.f, alloc);
        FilmBaseParameters fbp(Point2i(xres, yres), Bounds2i(Point2i(0, 0), Point2i(xres, yres)), filt, 35.f, sensor, "golden.pfm");
        RGBFilm *film = alloc.new_object<RGBFilm>(fbp, RGBColorSpace::sRGB, Infinity, false, alloc);
        Transform worldFromCamera = Inverse(LookAt(Point3f(0.3f, 1.1f, 3.2f), Point3f(0.1f, 0.6f, -0.2f), Vector3f(0.05f, 1, 0)));
        CameraTransform camXf{AnimatedTransform(worldFromCamera)};
        ParsedParameterVector params;
        auto addF = [&](const char *name, float v) { ParsedParameter *pp = new ParsedParameter(FileLoc()); pp->type = "float"; pp->name = name; pp->AddFloat(v); params.push_back(pp); };
        addF("fov", 38.f);
        if (cfg == 1) { addF("lensradius", 0.07f); addF("focaldistance", 3.1f); }
        ParameterDictionary dict(params, RGBColorSpace::sRGB);
        FileLoc loc;
        PerspectiveCamera *cam = PerspectiveCamera::Create(dict, camXf, Film(film), Medium(), &loc, alloc);
        ZSobolSampler sampler(spp, Point2i(xres, yres), RandomizeStrategy::FastOwen, 0);
        std::vector<float> cp;
        auto pushM = [&](const Transform &t) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) cp.push_back(t.GetMatrix()[i][j]); };
        cp.insert(cp.end(), {(float)xres, (float)yres, (float)spp, cam->lensRadius, cam->focalDistance});
        pushM(cam->cameraFromRaster);
        pushM(cam->cameraTransform.RenderFromCamera().startTransform);
        cp.insert(cp.end(), {cam->dxCamera.x, cam->dxCamera.y, cam->dxCamera.z, cam->dyCamera.x, cam->dyCamera.y, cam->dyCamera.z});
        std::vector<float> in, out;
        for (int q = 0; q < 1500; ++q) {
            Point2i pPixel(std::min(xres - 1, int(rng.Uniform<float>() * xres)), std::min(yres - 1, int(rng.Uniform<float>() * yres)));
            int sampleIndex = std::min(spp - 1, int(rng.Uniform<float>() * spp));
            sampler.StartPixelSample(pPixel, sampleIndex, 0);
            Float lu = sampler.Get1D();
            SampledWavelengths lambda = SampledWavelengths::SampleVisible(lu);
            CameraSample cs = GetCameraSample(Sampler(&sampler), pPixel, filt);
            pstd::optional<CameraRayDifferential> crd = cam->GenerateRayDifferential(cs, lambda);
            Float rayDiffScale = std::max<Float>(.125f, 1 / std::sqrt((Float)spp));
            crd->ray.ScaleDifferentials(rayDiffScale);
            const RayDifferential &r = crd->ray;
            in.insert(in.end(), {(float)pPixel.x, (float)pPixel.y, (float)sampleIndex});
            out.insert(out.end(), {lu, cs.pFilm.x, cs.pFilm.y, cs.filterWeight, cs.time, cs.pLens.x, cs.pLens.y, r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z,
                                   r.rxOrigin.x, r.rxOrigin.y, r.rxOrigin.z, r.rxDirection.x, r.rxDirection.y, r.rxDirection.z,
                                   r.ryOrigin.x, r.ryOrigin.y, r.ryOrigin.z, r.ryDirection.x, r.ryDirection.y, r.ryDirection.z,
                                   lambda[0], lambda[3], lambda.PDF()[0]});
        }
        b.f32(cfg == 0 ? "cam0_params" : "cam1_params", cp);
        b.f32(cfg == 0 ? "cam0_in" : "cam1_in", in);
        b.f32(cfg == 0 ? "cam0_out" : "cam1_out", out);
    }
    // Triangle::InteractionFromIntersection with per-vertex normals / uv (shading frame, error bounds, degenerate uv),
    // four mesh configurations: (normals, uv, reverse orientation) = (1,1,0) (1,0,1) (0,1,0) (0,0,1).  (appended last)
    for (int cfg = 0; cfg < 4; ++cfg) {
        Allocator alloc;
        const bool hasN = cfg < 2, hasUV = (cfg % 2) == 0, rev = (cfg % 2) == 1;
        const int nTri = 64;
        std::vector<int> idx; std::vector<Point3f> P; std::vector<Normal3f> N; std::vector<Point2f> UV;
        for (int i = 0; i < nTri; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f c(3 * r(), 3 * r(), 3 * r());
            Point3f q[3];
            for (int k = 0; k < 3; ++k) { q[k] = c + 0.8f * Vector3f(r(), r(), r()); P.push_back(q[k]); idx.push_back(3 * i + k); }
            Normal3f ng = Normal3f(Normalize(Cross(q[0] - q[2], q[1] - q[2])));
            for (int k = 0; k < 3; ++k) {
                Normal3f nv = (i % 5 == 4) ? Normal3f(Normalize(Vector3f(r(), r(), r()))) : Normal3f(Normalize(Vector3f(ng) + 0.4f * Vector3f(r(), r(), r())));
                N.push_back(nv);
                UV.push_back((i % 7 == 6) ? Point2f(0.25f, 0.5f) : Point2f(rng.Uniform<float>(), rng.Uniform<float>()));   // every 7th triangle: degenerate uv
            }
        }
        TriangleMesh *mesh = alloc.new_object<TriangleMesh>(Transform(), rev, idx, P, std::vector<Vector3f>(), hasN ? N : std::vector<Normal3f>(),
                                                            hasUV ? UV : std::vector<Point2f>(), std::vector<int>(), alloc);
        std::vector<float> md, in, out;
        md.insert(md.end(), {(float)nTri, hasN ? 1.f : 0.f, hasUV ? 1.f : 0.f, (rev ^ mesh->transformSwapsHandedness) ? 1.f : 0.f});
        for (int i = 0; i < 3 * nTri; ++i) {
            md.insert(md.end(), {mesh->p[i].x, mesh->p[i].y, mesh->p[i].z});
            if (hasN) md.insert(md.end(), {mesh->n[i].x, mesh->n[i].y, mesh->n[i].z}); else md.insert(md.end(), {0.f, 0.f, 0.f});
            if (hasUV) md.insert(md.end(), {mesh->uv[i].x, mesh->uv[i].y}); else md.insert(md.end(), {0.f, 0.f});
        }
        for (int q = 0; q < 256; ++q) {
            int tri = std::min(nTri - 1, int(rng.Uniform<float>() * nTri));
            float b0 = rng.Uniform<float>(), b1 = rng.Uniform<float>() * (1 - b0), b2 = 1 - b0 - b1;
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Vector3f wo = Normalize(Vector3f(r(), r(), r()));
            SurfaceInteraction si = Triangle::InteractionFromIntersection(mesh, tri, TriangleIntersection{b0, b1, b2, 1.f}, 0.f, wo);
            in.insert(in.end(), {(float)tri, b0, b1, b2});
            out.insert(out.end(), {si.pi.x.LowerBound(), si.pi.x.UpperBound(), si.pi.y.LowerBound(), si.pi.y.UpperBound(), si.pi.z.LowerBound(), si.pi.z.UpperBound(),
                                   si.n.x, si.n.y, si.n.z, si.shading.n.x, si.shading.n.y, si.shading.n.z, si.uv.x, si.uv.y,
                                   si.shading.dpdu.x, si.shading.dpdu.y, si.shading.dpdu.z});
        }
        b.f32(StringPrintf("triintr%d_mesh", cfg).c_str(), md);
        b.f32(StringPrintf("triintr%d_in", cfg).c_str(), in);
        b.f32(StringPrintf("triintr%d_out", cfg).c_str(), out);
    }
    // Material::GetBxDF parameter mapping (roughness remap, spectral eta / k sampling, reflectance -> k, termination of the
    // secondary wavelengths by a dispersive interface): the BxDF members the reference ends up with.  (appended last)
    {
        Allocator alloc;
        auto F = [&](float v) { return FloatTexture(alloc.new_object<FloatConstantTexture>(v)); };
        auto S = [&](Spectrum sp) { return SpectrumTexture(alloc.new_object<SpectrumConstantTexture>(sp)); };
        Spectrum cuEta = GetNamedSpectrum("metal-Cu-eta"), cuK = GetNamedSpectrum("metal-Cu-k"), bk7 = GetNamedSpectrum("glass-BK7");
        Spectrum c15 = alloc.new_object<ConstantSpectrum>(1.5f), c07 = alloc.new_object<ConstantSpectrum>(0.7f), c0 = alloc.new_object<ConstantSpectrum>(0.f);
        ConductorMaterial condSpectral(S(cuEta), S(cuK), nullptr, F(0.3f), F(0.05f), nullptr, nullptr, true);
        ConductorMaterial condRefl(nullptr, nullptr, S(c07), F(0.2f), F(0.2f), nullptr, nullptr, false);
        DielectricMaterial dielConst(F(0.f), F(0.f), c15, nullptr, nullptr, true);
        DielectricMaterial dielDisp(F(0.1f), F(0.25f), bk7, nullptr, nullptr, true);
        CoatedConductorMaterial coated(F(0.f), F(0.f), F(0.01f), c15, F(0.f), S(c0), F(0.f), F(0.f), S(cuEta), S(cuK), nullptr, nullptr, nullptr, true, 10, 1);
        CoatedConductorMaterial coated2(F(0.2f), F(0.1f), F(0.05f), bk7, F(0.3f), S(c07), F(0.3f), F(0.15f), nullptr, nullptr, S(c07), nullptr, nullptr, false, 6, 2);
        std::vector<float> in, out;
        for (int q = 0; q < 64; ++q) {
            Float u = rng.Uniform<float>();
            in.push_back(u);
            SurfaceInteraction si;
            MaterialEvalContext ctx(si);
            UniversalTextureEvaluator te;
            auto lam = [&]() { return SampledWavelengths::SampleVisible(u); };
            { SampledWavelengths l = lam(); ConductorBxDF bx = condSpectral.GetBxDF(te, ctx, l);
              out.insert(out.end(), {bx.mfDistrib.alpha_x, bx.mfDistrib.alpha_y, bx.eta[0], bx.eta[3], bx.k[0], bx.k[3], l.PDF()[0], l.PDF()[3]}); }
            { SampledWavelengths l = lam(); ConductorBxDF bx = condRefl.GetBxDF(te, ctx, l);
              out.insert(out.end(), {bx.mfDistrib.alpha_x, bx.mfDistrib.alpha_y, bx.eta[0], bx.eta[3], bx.k[0], bx.k[3], l.PDF()[0], l.PDF()[3]}); }
            { SampledWavelengths l = lam(); DielectricBxDF bx = dielConst.GetBxDF(te, ctx, l);
              out.insert(out.end(), {bx.mfDistrib.alpha_x, bx.mfDistrib.alpha_y, bx.eta, l.PDF()[0], l.PDF()[3]}); }
            { SampledWavelengths l = lam(); DielectricBxDF bx = dielDisp.GetBxDF(te, ctx, l);
              out.insert(out.end(), {bx.mfDistrib.alpha_x, bx.mfDistrib.alpha_y, bx.eta, l.PDF()[0], l.PDF()[3]}); }
            for (int c = 0; c < 2; ++c) {
                SampledWavelengths l = lam(); CoatedConductorBxDF bx = (c ? coated2 : coated).GetBxDF(te, ctx, l);
                out.insert(out.end(), {bx.top.mfDistrib.alpha_x, bx.top.mfDistrib.alpha_y, bx.top.eta, bx.bottom.mfDistrib.alpha_x, bx.bottom.mfDistrib.alpha_y,
                                       bx.bottom.eta[0], bx.bottom.eta[3], bx.bottom.k[0], bx.bottom.k[3], bx.thickness, bx.g, bx.albedo[0], (float)bx.maxDepth,
                                       (float)bx.nSamples, l.PDF()[0], l.PDF()[3]});
            }
        }
        b.f32("mat_in", in);
        b.f32("mat_out", out);
    }
    // Image textures: MIPMap::Filter<RGB> / <Float> (pyramid generation incl. resize of non-power-of-two images and 8-bit re-encoding
    // of the levels; point / bilinear / trilinear / EWA; repeat / clamp / black) on three synthetic images.  (appended last)
    {
        Allocator alloc;
        struct Img { int w, h, nc; bool is8; };
        const Img imgs[3] = {{13, 9, 3, true}, {16, 8, 1, false}, {8, 8, 4, true}};
        std::vector<float> q;
        const int nQ = 48;
        for (int i = 0; i < nQ; ++i) {
            float s = -0.5f + 2 * rng.Uniform<float>(), t = -0.5f + 2 * rng.Uniform<float>();
            auto mag = [&]() { return (rng.Uniform<float>() < 0.1f ? 0.f : std::pow(10.f, -3 + 3 * rng.Uniform<float>())) * (rng.Uniform<float>() < 0.5f ? -1.f : 1.f); };
            q.insert(q.end(), {s, t, mag(), mag(), mag(), mag()});
        }
        b.f32("tex_queries", q);
        for (int k = 0; k < 3; ++k) {
            const Img &im = imgs[k];
            std::vector<float> pix((size_t)im.w * im.h * im.nc);
            for (float &v : pix) v = im.is8 ? (float)(int)std::min(255.f, 256 * rng.Uniform<float>()) : 2 * rng.Uniform<float>();
            b.f32(StringPrintf("tex%d_pixels", k).c_str(), pix);
            std::vector<std::string> names = im.nc == 1 ? std::vector<std::string>{"Y"} : im.nc == 3 ? std::vector<std::string>{"R", "G", "B"}
                                                                                                 : std::vector<std::string>{"R", "G", "B", "A"};
            std::vector<float> out;
            for (int wrap = 0; wrap < 3; ++wrap)
                for (int filt = 0; filt < 4; ++filt) {
                    if (wrap == 0 && k == 0) {   // the reference aborts when it resizes a non-power-of-two image under WrapMode::Black (util/image.cpp:114)
                        for (int i = 0; i < 4 * nQ; ++i) out.push_back(-1.f);
                        continue;
                    }
                    Image image(im.is8 ? PixelFormat::U256 : PixelFormat::Float, Point2i(im.w, im.h), names, im.is8 ? ColorEncoding::sRGB : ColorEncoding());
                    for (size_t i = 0; i < pix.size(); ++i) {
                        if (im.is8) ((uint8_t *)image.RawPointer({0, 0}))[i] = (uint8_t)pix[i];
                        else ((float *)image.RawPointer({0, 0}))[i] = pix[i];
                    }
                    MIPMapFilterOptions opt;
                    opt.filter = (FilterFunction)filt; opt.maxAnisotropy = 8.f;
                    MIPMap mip(image, RGBColorSpace::sRGB, (WrapMode)wrap, alloc, opt);
                    for (int i = 0; i < nQ; ++i) {
                        const float *a = &q[6 * i];
                        RGB c = mip.Filter<RGB>(Point2f(a[0], a[1]), Vector2f(a[2], a[3]), Vector2f(a[4], a[5]));
                        Float f = mip.Filter<Float>(Point2f(a[0], a[1]), Vector2f(a[2], a[3]), Vector2f(a[4], a[5]));
                        out.insert(out.end(), {c.r, c.g, c.b, f});
                    }
                }
            b.f32(StringPrintf("tex%d_out", k).c_str(), out);
        }
    }
    // Differentials for texture filtering: CameraBase::FindMinimumDifferentials, Camera::Approximate_dp_dxy,
    // SurfaceInteraction::ComputeDifferentials (with and without ray differentials) and SpawnRay's specular differentials.  (appended last)
    for (int cfg = 0; cfg < 2; ++cfg) {
        Allocator alloc;
        const int xres = 96, yres = 64, spp = 16;
        Filter filt = alloc.new_object<GaussianFilter>(Vector2f(1.5f, 1.5f), 0.5f, alloc);
        PixelSensor *sensor = alloc.new_object<PixelSensor>(RGBColorSpace::sRGB, nullptr, 1.f, alloc);
        FilmBaseParameters fbp(Point2i(xres, yres), Bounds2i(Point2i(0, 0), Point2i(xres, yres)), filt, 35.f, sensor, "golden.pfm");
        RGBFilm *film = alloc.new_object<RGBFilm>(fbp, RGBColorSpace::sRGB, Infinity, false, alloc);
        Transform worldFromCamera = Inverse(LookAt(Point3f(0.3f, 1.1f, 3.2f), Point3f(0.1f, 0.6f, -0.2f), Vector3f(0.05f, 1, 0)));
        CameraTransform camXf{AnimatedTransform(worldFromCamera)};
        ParsedParameterVector params;
        auto addF = [&](const char *name, float v) { ParsedParameter *pp = new ParsedParameter(FileLoc()); pp->type = "float"; pp->name = name; pp->AddFloat(v); params.push_back(pp); };
        addF("fov", 38.f);
        if (cfg == 1) { addF("lensradius", 0.07f); addF("focaldistance", 3.1f); }
        ParameterDictionary dict(params, RGBColorSpace::sRGB);
        FileLoc loc;
        PerspectiveCamera *cam = PerspectiveCamera::Create(dict, camXf, Film(film), Medium(), &loc, alloc);
        std::vector<float> cp;
        auto pushM = [&](const SquareMatrix<4> &t) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) cp.push_back(t[i][j]); };
        cp.insert(cp.end(), {(float)xres, (float)yres, (float)spp, cam->lensRadius, cam->focalDistance});
        pushM(cam->cameraFromRaster.GetMatrix());
        pushM(cam->cameraTransform.RenderFromCamera().startTransform.GetMatrix());
        pushM(cam->cameraTransform.RenderFromCamera().startTransform.GetInverseMatrix());
        cp.insert(cp.end(), {cam->dxCamera.x, cam->dxCamera.y, cam->dxCamera.z, cam->dyCamera.x, cam->dyCamera.y, cam->dyCamera.z});
        for (Vector3f v : {cam->minPosDifferentialX, cam->minPosDifferentialY, cam->minDirDifferentialX, cam->minDirDifferentialY}) cp.insert(cp.end(), {v.x, v.y, v.z});
        b.f32(StringPrintf("diff%d_params", cfg).c_str(), cp);
        std::vector<float> in, out;
        for (int qi = 0; qi < 256; ++qi) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f p(1.5f * r(), 1 + r(), 1.5f * r());
            Vector3f dpdu(r(), r(), r()), dpdv(r(), r(), r());
            Normal3f dndu(0.3f * r(), 0.3f * r(), 0.3f * r()), dndv(0.3f * r(), 0.3f * r(), 0.3f * r());
            Vector3f wo = Normalize(Vector3f(r(), r(), r()));
            SurfaceInteraction si(Point3fi(p), Point2f(rng.Uniform<float>(), rng.Uniform<float>()), wo, dpdu, dpdv, dndu, dndv, 0.f, false);
            RayDifferential ray(Point3f(0.3f, 1.1f, 3.2f), -wo);
            const int mode = qi % 4;   // 0: no differentials (Approximate_dp_dxy)  1..3: ray differentials, then reflection / transmission / diffuse
            if (mode) {
                ray.hasDifferentials = true;
                ray.rxOrigin = ray.o + 0.01f * Vector3f(r(), r(), r()); ray.ryOrigin = ray.o + 0.01f * Vector3f(r(), r(), r());
                ray.rxDirection = ray.d + 0.02f * Vector3f(r(), r(), r()); ray.ryDirection = ray.d + 0.02f * Vector3f(r(), r(), r());
            }
            si.ComputeDifferentials(ray, Camera(cam), spp);
            Vector3f wi = Normalize(Vector3f(r(), r(), r()));
            Float eta = 1 + rng.Uniform<float>();
            int flags = mode == 1 ? (int)BxDFFlags::SpecularReflection : mode == 2 ? (int)BxDFFlags::SpecularTransmission : (int)BxDFFlags::DiffuseReflection;
            RayDifferential rd = si.SpawnRay(ray, BSDF(), wi, flags, eta);
            in.insert(in.end(), {p.x, p.y, p.z, dpdu.x, dpdu.y, dpdu.z, dpdv.x, dpdv.y, dpdv.z, dndu.x, dndu.y, dndu.z, dndv.x, dndv.y, dndv.z, wo.x, wo.y, wo.z,
                                 ray.hasDifferentials ? 1.f : 0.f, ray.rxOrigin.x, ray.rxOrigin.y, ray.rxOrigin.z, ray.ryOrigin.x, ray.ryOrigin.y, ray.ryOrigin.z,
                                 ray.rxDirection.x, ray.rxDirection.y, ray.rxDirection.z, ray.ryDirection.x, ray.ryDirection.y, ray.ryDirection.z,
                                 wi.x, wi.y, wi.z, eta, (float)flags, si.n.x, si.n.y, si.n.z, si.shading.n.x, si.shading.n.y, si.shading.n.z});
            out.insert(out.end(), {si.dpdx.x, si.dpdx.y, si.dpdx.z, si.dpdy.x, si.dpdy.y, si.dpdy.z, si.dudx, si.dudy, si.dvdx, si.dvdy,
                                   rd.hasDifferentials ? 1.f : 0.f, rd.rxOrigin.x, rd.rxOrigin.y, rd.rxOrigin.z, rd.ryOrigin.x, rd.ryOrigin.y, rd.ryOrigin.z,
                                   rd.rxDirection.x, rd.rxDirection.y, rd.rxDirection.z, rd.ryDirection.x, rd.ryDirection.y, rd.ryDirection.z});
        }
        b.f32(StringPrintf("diff%d_in", cfg).c_str(), in);
        b.f32(StringPrintf("diff%d_out", cfg).c_str(), out);
    }
    // SobolSampler and HaltonSampler: for random (pixel, sample index) the values a path draws first -- Get1D, GetPixel2D, Get1D, Get2D,
    // then 3 x (Get1D, Get2D) -- and a restart at dimension 997 (both samplers wrap back to dimension 2).  (appended last)
    {
        Allocator alloc;
        const Point2i res(200, 150);
        struct Cfg { int kind, spp, seed; RandomizeStrategy r; };
        const Cfg cfgs[7] = {{0, 8, 0, RandomizeStrategy::FastOwen}, {0, 5, 3, RandomizeStrategy::None}, {0, 4, 1, RandomizeStrategy::PermuteDigits}, {0, 16, 2, RandomizeStrategy::Owen},
                             {1, 8, 0, RandomizeStrategy::PermuteDigits}, {1, 5, 7, RandomizeStrategy::None}, {1, 4, 2, RandomizeStrategy::Owen}};
        std::vector<float> in, out;
        for (int c = 0; c < 7; ++c) {
            SobolSampler sob(cfgs[c].spp, res, cfgs[c].r, cfgs[c].seed);
            HaltonSampler hal(cfgs[c].spp, res, cfgs[c].r == RandomizeStrategy::FastOwen ? RandomizeStrategy::PermuteDigits : cfgs[c].r, cfgs[c].seed, alloc);
            Sampler smp = cfgs[c].kind == 0 ? Sampler(&sob) : Sampler(&hal);
            for (int q = 0; q < 96; ++q) {
                Point2i p(std::min(res.x - 1, int(rng.Uniform<float>() * res.x)), std::min(res.y - 1, int(rng.Uniform<float>() * res.y)));
                int si = std::min(cfgs[c].spp - 1, int(rng.Uniform<float>() * cfgs[c].spp));
                in.insert(in.end(), {(float)c, (float)p.x, (float)p.y, (float)si});
                smp.StartPixelSample(p, si, 0);
                out.push_back(smp.Get1D());
                Point2f px = smp.GetPixel2D(); out.push_back(px.x); out.push_back(px.y);
                out.push_back(smp.Get1D());
                Point2f l = smp.Get2D(); out.push_back(l.x); out.push_back(l.y);
                for (int k = 0; k < 3; ++k) { out.push_back(smp.Get1D()); Point2f u = smp.Get2D(); out.push_back(u.x); out.push_back(u.y); }
                smp.StartPixelSample(p, si, 997);
                for (int k = 0; k < 2; ++k) { out.push_back(smp.Get1D()); Point2f u = smp.Get2D(); out.push_back(u.x); out.push_back(u.y); }
            }
        }
        b.f32("smp_in", in);
        b.f32("smp_out", out);
    }
    // RGBGridMedium: the majorant grid and SampleT_maj event traces with sigma_a, sigma_s and Le at the events.  (appended last)
    {
        Allocator alloc;
        const int nx = 7, ny = 9, nz = 5;
        std::vector<float> rgbA, rgbS, rgbL;
        std::vector<RGBUnboundedSpectrum> va, vs;
        std::vector<RGBIlluminantSpectrum> vl;
        for (int i = 0; i < nx * ny * nz; ++i) {
            auto c = [&](float sc) { float r = rng.Uniform<float>(); return r < 0.25f ? 0.f : sc * rng.Uniform<float>(); };
            RGB a(c(1.5f), c(1.5f), c(1.5f)), sg(c(4.f), c(4.f), c(4.f)), le(c(2.f), c(1.f), c(.5f));
            if (i % 11 == 0) a = RGB(0.4f, 0.4f, 0.4f);   // grey voxels take the constant-coefficient branch of the RGB -> spectrum table
            rgbA.insert(rgbA.end(), {a.r, a.g, a.b}); rgbS.insert(rgbS.end(), {sg.r, sg.g, sg.b}); rgbL.insert(rgbL.end(), {le.r, le.g, le.b});
            va.push_back(RGBUnboundedSpectrum(*RGBColorSpace::sRGB, a)); vs.push_back(RGBUnboundedSpectrum(*RGBColorSpace::sRGB, sg));
            vl.push_back(RGBIlluminantSpectrum(*RGBColorSpace::sRGB, le));
        }
        Transform rfm = Translate(Vector3f(-0.2f, 0.1f, 0.3f)) * RotateX(20) * Scale(1.2f, 1.f, 0.8f);
        Bounds3f bounds(Point3f(-1, 0, -0.5f), Point3f(1, 1.5f, 1));
        RGBGridMedium med(bounds, rfm, 0.2f, SampledGrid<RGBUnboundedSpectrum>(va, nx, ny, nz, alloc), SampledGrid<RGBUnboundedSpectrum>(vs, nx, ny, nz, alloc), 1.3f,
                          SampledGrid<RGBIlluminantSpectrum>(vl, nx, ny, nz, alloc), 0.7f, alloc);
        std::vector<float> params = {(float)nx, (float)ny, (float)nz, bounds.pMin.x, bounds.pMin.y, bounds.pMin.z, bounds.pMax.x, bounds.pMax.y, bounds.pMax.z, 1.3f, 0.7f};
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) params.push_back(rfm.GetInverseMatrix()[i][j]);
        b.f32("rgbmed_params", params);
        b.f32("rgbmed_sigma_a", rgbA); b.f32("rgbmed_sigma_s", rgbS); b.f32("rgbmed_Le", rgbL);
        b.f32("rgbmed_majorant", std::vector<float>(med.majorantGrid.voxels.begin(), med.majorantGrid.voxels.end()));
        const int kEvents = 5;
        std::vector<float> in, out;
        for (int i = 0; i < 512; ++i) {
            auto r = [&]() { return 2 * rng.Uniform<float>() - 1; };
            Point3f o(2 * r(), 2 * r(), 2 * r());
            if (i % 3 == 0) o = rfm(Point3f(bounds.Lerp(Point3f(rng.Uniform<float>(), rng.Uniform<float>(), rng.Uniform<float>()))));
            Vector3f d = Normalize(Vector3f(r(), r(), r())) * (0.5f + 1.5f * rng.Uniform<float>());
            Float tMax = (i % 4 == 3) ? Infinity : 4 * rng.Uniform<float>();
            Float u = rng.Uniform<float>();
            uint64_t s0 = rng.Uniform<uint32_t>() >> 8, s1 = rng.Uniform<uint32_t>() >> 8;
            SampledWavelengths lambda = SampledWavelengths::SampleVisible(rng.Uniform<float>());
            RNG trng(s0, s1);
            std::vector<float> ev;
            int nEv = 0;
            auto cb = [&](Point3f p, MediumProperties mp, SampledSpectrum sigma_maj, SampledSpectrum T_maj) {
                ev.insert(ev.end(), {p.x, p.y, p.z, mp.sigma_a[0], mp.sigma_a[3], mp.sigma_s[0], mp.sigma_s[2], mp.Le[0], mp.Le[1], sigma_maj[0], T_maj[0], T_maj[3]});
                return ++nEv < kEvents;
            };
            Ray ray(o, d, 0.f, Medium(&med));
            SampledSpectrum T = SampleT_maj<RGBGridMedium>(ray, tMax, u, trng, lambda, cb);
            ev.resize(12 * kEvents, 0.f);
            in.insert(in.end(), {o.x, o.y, o.z, d.x, d.y, d.z, tMax, u, (float)s0, (float)s1, lambda[0], lambda[1], lambda[2], lambda[3]});
            out.push_back((float)nEv);
            out.insert(out.end(), ev.begin(), ev.end());
            out.insert(out.end(), {T[0], T[1], T[2], T[3], trng.Uniform<float>()});
        }
        b.f32("rgbmed_in", in);
        b.f32("rgbmed_out", out);
    }
    // PMJ02BNSampler: the same draw sequence as the Sobol' / Halton vectors above, three sample counts (one not a power of four).  (appended last)
    {
        Allocator alloc;
        const int spps[3] = {16, 4, 6};
        const int seeds[3] = {0, 5, 1};
        std::vector<float> in, out;
        for (int c = 0; c < 3; ++c) {
            PMJ02BNSampler pmj(spps[c], seeds[c], alloc);
            Sampler smp(&pmj);
            for (int q = 0; q < 96; ++q) {
                Point2i p(int(rng.Uniform<float>() * 300), int(rng.Uniform<float>() * 200));
                int si = std::min(spps[c] - 1, int(rng.Uniform<float>() * spps[c]));
                in.insert(in.end(), {(float)c, (float)p.x, (float)p.y, (float)si});
                smp.StartPixelSample(p, si, 0);
                out.push_back(smp.Get1D());
                Point2f px = smp.GetPixel2D(); out.push_back(px.x); out.push_back(px.y);
                out.push_back(smp.Get1D());
                Point2f l = smp.Get2D(); out.push_back(l.x); out.push_back(l.y);
                for (int k = 0; k < 5; ++k) { out.push_back(smp.Get1D()); Point2f u = smp.Get2D(); out.push_back(u.x); out.push_back(u.y); }
            }
        }
        b.f32("pmj_in", in);
        b.f32("pmj_out", out);
    }
    // quadrics (Sphere / Disk / Cylinder): constructor-derived members, bounds, BasicIntersect + InteractionFromIntersection,
    // Sample(ctx, u) and PDF(ctx, wi) -- for csrc/quadrics.h (tests/tools/quadric_host_check.cpp)
    {
        RNG qr(977);
        auto U = [&]() { return qr.Uniform<float>(); };
        std::vector<float> shapes, rays, hits, samp, sampOut;
        const int nShapes = 18, nRays = 400, nSamp = 96;
        for (int k = 0; k < nShapes; ++k) {
            int type = k % 3;   // 0 sphere, 1 disk, 2 cylinder
            Transform rfo = Translate(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1)) * Rotate(360 * U(), Normalize(Vector3f(U() - .5f, U() - .5f, U() - .3f)));
            if (k >= 6) rfo = rfo * Scale(.5f + U(), .5f + U(), (k % 2 ? -1.f : 1.f) * (.5f + U()));
            if (k < 3) rfo = Transform();
            // the scene's way of deriving the inverse (scene.h:475-477): rebuilt from the product matrix
            rfo = Transform(rfo.GetMatrix());
            Transform ofr = Inverse(rfo);
            bool reverse = (k / 3) % 2 == 1;
            float radius = .4f + U(), zmin = -radius * (k < 9 ? 2.f : U()), zmax = radius * (k < 9 ? 2.f : U()), phimax = k < 9 ? 360.f : 90.f + 250.f * U();
            if (type == 2) { zmin = -.3f - U(); zmax = .3f + U(); }
            float height = U() - .5f, inner = k < 9 ? 0.f : .3f * radius * U();
            Sphere sph(&rfo, &ofr, reverse, radius, zmin, zmax, phimax);
            Disk dsk(&rfo, &ofr, reverse, height, radius, inner, phimax);
            Cylinder cyl(&rfo, &ofr, reverse, radius, zmin, zmax, phimax);
            Shape shape = type == 0 ? Shape(&sph) : (type == 1 ? Shape(&dsk) : Shape(&cyl));
            shapes.insert(shapes.end(), {(float)type, radius, zmin, zmax, phimax, height, inner, reverse ? 1.f : 0.f});
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) shapes.push_back(rfo.GetMatrix()[i][j]);
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) shapes.push_back(ofr.GetMatrix()[i][j]);
            Bounds3f bb = shape.Bounds();
            DirectionCone nb = shape.NormalBounds();
            shapes.insert(shapes.end(), {bb.pMin.x, bb.pMin.y, bb.pMin.z, bb.pMax.x, bb.pMax.y, bb.pMax.z, shape.Area(), nb.w.x, nb.w.y, nb.w.z, nb.cosTheta,
                                         rfo.SwapsHandedness() ? 1.f : 0.f});
            Point3f c = rfo(Point3f(0, 0, type == 1 ? height : 0));
            for (int r = 0; r < nRays; ++r) {
                Point3f o = c + (r % 5 == 0 ? .3f : 3.f) * Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1);
                Point3f tgt = c + radius * Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1);
                Vector3f d = tgt - o;
                if (r % 3 == 0) d = Normalize(d);
                float tMax = r % 7 == 0 ? 2 * U() : Infinity;
                rays.insert(rays.end(), {o.x, o.y, o.z, d.x, d.y, d.z, tMax});
                Ray ray(o, d);
                pstd::optional<ShapeIntersection> si = shape.Intersect(ray, tMax);
                float out[40] = {0};
                out[0] = si ? 1.f : 0.f;
                out[39] = shape.IntersectP(ray, tMax) ? 1.f : 0.f;
                if (si) {
                    const SurfaceInteraction &in = si->intr;
                    float v[] = {si->tHit, in.pi.x.LowerBound(), in.pi.x.UpperBound(), in.pi.y.LowerBound(), in.pi.y.UpperBound(), in.pi.z.LowerBound(), in.pi.z.UpperBound(),
                                 in.n.x, in.n.y, in.n.z, in.shading.n.x, in.shading.n.y, in.shading.n.z, in.uv[0], in.uv[1], in.dpdu.x, in.dpdu.y, in.dpdu.z,
                                 in.dpdv.x, in.dpdv.y, in.dpdv.z, in.dndu.x, in.dndu.y, in.dndu.z, in.dndv.x, in.dndv.y, in.dndv.z, in.wo.x, in.wo.y, in.wo.z,
                                 in.shading.dpdu.x, in.shading.dpdu.y, in.shading.dpdu.z};
                    for (size_t q = 0; q < sizeof(v) / sizeof(v[0]); ++q) out[1 + q] = v[q];
                }
                hits.insert(hits.end(), out, out + 40);
            }
            for (int r = 0; r < nSamp; ++r) {
                // reference points outside, near and (spheres) inside the shape
                float dist = r % 4 == 0 ? .5f * radius : (r % 4 == 1 ? 1.05f * radius : 4.f + 40.f * U());
                Point3f p = c + dist * Normalize(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1));
                Vector3f err = r % 2 ? Vector3f(1e-5f * U(), 1e-5f * U(), 1e-5f * U()) : Vector3f(0, 0, 0);
                Normal3f n = r % 3 == 0 ? Normal3f(0, 0, 0) : Normal3f(Normalize(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1)));
                Point2f u(U(), U());
                ShapeSampleContext ctx(Point3fi(p, err), n, n, 0.f);
                samp.insert(samp.end(), {ctx.pi.x.LowerBound(), ctx.pi.x.UpperBound(), ctx.pi.y.LowerBound(), ctx.pi.y.UpperBound(), ctx.pi.z.LowerBound(), ctx.pi.z.UpperBound(),
                                         n.x, n.y, n.z, u[0], u[1]});
                pstd::optional<ShapeSample> ss = shape.Sample(ctx, u);
                float out[20] = {0};
                out[0] = ss ? 1.f : 0.f;
                Vector3f wi = Normalize(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1));
                if (ss) {
                    const Interaction &in = ss->intr;
                    float v[] = {in.pi.x.LowerBound(), in.pi.x.UpperBound(), in.pi.y.LowerBound(), in.pi.y.UpperBound(), in.pi.z.LowerBound(), in.pi.z.UpperBound(),
                                 in.n.x, in.n.y, in.n.z, in.uv[0], in.uv[1], ss->pdf};
                    for (size_t q = 0; q < sizeof(v) / sizeof(v[0]); ++q) out[1 + q] = v[q];
                    if (r % 2 == 0 && LengthSquared(in.p() - ctx.p()) > 0) wi = Normalize(in.p() - ctx.p());
                }
                out[13] = wi.x; out[14] = wi.y; out[15] = wi.z;
                out[16] = shape.PDF(ctx, wi);
                sampOut.insert(sampOut.end(), out, out + 20);
            }
        }
        b.f32("quad_shapes", shapes);
        b.f32("quad_rays", rays);
        b.f32("quad_hits", hits);
        b.f32("quad_samp_in", samp);
        b.f32("quad_samp_out", sampOut);
    }
    // bilinear patches: meshes of one patch each (flat rectangle, flat non-rectangle, twisted, degenerate-to-triangle; with and without
    // uv / N; mirrored transform) -- render-space vertices, area, IsRectangle, bounds, NormalBounds, Intersect + interaction,
    // Sample(ctx, u), PDF(ctx, wi) -- for csrc/bilinear.h (tests/tools/bilinear_host_check.cpp)
    {
        RNG qr(4242);
        auto U = [&]() { return qr.Uniform<float>(); };
        Allocator alloc;
        std::vector<float> patches, rays, hits, samp, sampOut;
        const int nPatches = 16, nRays = 300, nSamp = 96;
        for (int k = 0; k < nPatches; ++k) {
            int kind = k % 4;   // 0 rectangle, 1 planar quad, 2 twisted, 3 triangle (two coincident vertices)
            bool withUV = (k / 4) % 2 == 1, withN = k >= 8, reverse = k % 5 == 2;
            Transform rfo = Translate(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1)) * Rotate(360 * U(), Normalize(Vector3f(U() - .5f, U() - .5f, U() - .3f)));
            if (k >= 12) rfo = rfo * Scale(-1.f - U(), .5f + U(), .5f + U());
            rfo = Transform(rfo.GetMatrix());
            float w = .5f + 1.5f * U(), h = .5f + 1.5f * U();
            std::vector<Point3f> P = {Point3f(0, 0, 0), Point3f(w, 0, 0), Point3f(0, h, 0), Point3f(w, h, 0)};
            if (kind == 1) { P[3] = Point3f(w * (1.2f + U()), h * (.6f + U()), 0); P[1].y = .2f * U(); }
            if (kind == 2) { P[3].z = .3f + U(); P[0].z = -.4f * U(); }
            if (kind == 3) P[3] = P[1];
            std::vector<Point2f> UV;
            if (withUV) UV = {Point2f(.1f * U(), .1f * U()), Point2f(1 + U(), .2f * U()), Point2f(.2f * U(), 1 + U()), Point2f(1.5f + U(), 1.3f + U())};
            std::vector<Normal3f> N;
            if (withN) for (int i = 0; i < 4; ++i) N.push_back(Normal3f(.4f * (U() - .5f), .4f * (U() - .5f), (k % 2 ? -1.f : 1.f)));
            BilinearPatchMesh *mesh = alloc.new_object<BilinearPatchMesh>(rfo, reverse, std::vector<int>{0, 1, 2, 3}, P, N, UV, std::vector<int>(), nullptr, alloc);
            pstd::vector<Shape> shapes = BilinearPatch::CreatePatches(mesh, alloc);
            Shape shape = shapes[0];
            const BilinearPatch *blp = shape.Cast<BilinearPatch>();
            patches.insert(patches.end(), {withUV ? 1.f : 0.f, withN ? 1.f : 0.f, (reverse ^ mesh->transformSwapsHandedness) ? 1.f : 0.f, blp->IsRectangle(mesh) ? 1.f : 0.f, shape.Area()});
            for (int i = 0; i < 4; ++i) {
                patches.insert(patches.end(), {mesh->p[i].x, mesh->p[i].y, mesh->p[i].z});
                patches.insert(patches.end(), {withUV ? mesh->uv[i].x : 0.f, withUV ? mesh->uv[i].y : 0.f});
                patches.insert(patches.end(), {withN ? mesh->n[i].x : 0.f, withN ? mesh->n[i].y : 0.f, withN ? mesh->n[i].z : 0.f});
            }
            Bounds3f bb = shape.Bounds();
            DirectionCone nb = shape.NormalBounds();
            patches.insert(patches.end(), {bb.pMin.x, bb.pMin.y, bb.pMin.z, bb.pMax.x, bb.pMax.y, bb.pMax.z, nb.w.x, nb.w.y, nb.w.z, nb.cosTheta});
            Point3f c = Lerp(.5f, Lerp(.5f, mesh->p[0], mesh->p[2]), Lerp(.5f, mesh->p[1], mesh->p[3]));
            for (int r = 0; r < nRays; ++r) {
                Point3f o = c + (r % 5 == 0 ? .4f : 3.f) * Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1);
                Point3f tgt = Lerp(1.2f * U() - .1f, Lerp(1.2f * U() - .1f, mesh->p[0], mesh->p[2]), Lerp(U(), mesh->p[1], mesh->p[3]));
                Vector3f d = tgt - o;
                if (r % 3 == 0) d = Normalize(d);
                float tMax = r % 7 == 0 ? 2 * U() : Infinity;
                rays.insert(rays.end(), {o.x, o.y, o.z, d.x, d.y, d.z, tMax});
                Ray ray(o, d);
                pstd::optional<ShapeIntersection> si = shape.Intersect(ray, tMax);
                float out[40] = {0};
                out[0] = si ? 1.f : 0.f;
                out[39] = shape.IntersectP(ray, tMax) ? 1.f : 0.f;
                if (si) {
                    const SurfaceInteraction &in = si->intr;
                    float v[] = {si->tHit, in.pi.x.LowerBound(), in.pi.x.UpperBound(), in.pi.y.LowerBound(), in.pi.y.UpperBound(), in.pi.z.LowerBound(), in.pi.z.UpperBound(),
                                 in.n.x, in.n.y, in.n.z, in.shading.n.x, in.shading.n.y, in.shading.n.z, in.uv[0], in.uv[1], in.dpdu.x, in.dpdu.y, in.dpdu.z,
                                 in.dpdv.x, in.dpdv.y, in.dpdv.z, in.shading.dndu.x, in.shading.dndu.y, in.shading.dndu.z, in.shading.dndv.x, in.shading.dndv.y, in.shading.dndv.z,
                                 in.shading.dpdu.x, in.shading.dpdu.y, in.shading.dpdu.z, in.shading.dpdv.x, in.shading.dpdv.y, in.shading.dpdv.z};
                    for (size_t q = 0; q < sizeof(v) / sizeof(v[0]); ++q) out[1 + q] = v[q];
                }
                hits.insert(hits.end(), out, out + 40);
            }
            for (int r = 0; r < nSamp; ++r) {
                float dist = r % 4 == 0 ? .3f : (r % 4 == 1 ? 1.5f : 6.f + 60.f * U());
                Point3f p = c + dist * Normalize(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1));
                Vector3f err = r % 2 ? Vector3f(1e-5f * U(), 1e-5f * U(), 1e-5f * U()) : Vector3f(0, 0, 0);
                Normal3f n = r % 3 == 0 ? Normal3f(0, 0, 0) : Normal3f(Normalize(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1)));
                Point2f u(U(), U());
                ShapeSampleContext ctx(Point3fi(p, err), n, n, 0.f);
                samp.insert(samp.end(), {ctx.pi.x.LowerBound(), ctx.pi.x.UpperBound(), ctx.pi.y.LowerBound(), ctx.pi.y.UpperBound(), ctx.pi.z.LowerBound(), ctx.pi.z.UpperBound(),
                                         n.x, n.y, n.z, u[0], u[1]});
                pstd::optional<ShapeSample> ss = shape.Sample(ctx, u);
                float out[20] = {0};
                out[0] = ss ? 1.f : 0.f;
                Vector3f wi = Normalize(Vector3f(2 * U() - 1, 2 * U() - 1, 2 * U() - 1));
                if (ss) {
                    const Interaction &in = ss->intr;
                    float v[] = {in.pi.x.LowerBound(), in.pi.x.UpperBound(), in.pi.y.LowerBound(), in.pi.y.UpperBound(), in.pi.z.LowerBound(), in.pi.z.UpperBound(),
                                 in.n.x, in.n.y, in.n.z, in.uv[0], in.uv[1], ss->pdf};
                    for (size_t q = 0; q < sizeof(v) / sizeof(v[0]); ++q) out[1 + q] = v[q];
                    if (r % 8 != 7 && LengthSquared(in.p() - ctx.p()) > 0) wi = Normalize(in.p() - ctx.p());
                }
                out[13] = wi.x; out[14] = wi.y; out[15] = wi.z;
                out[16] = shape.PDF(ctx, wi);
                sampOut.insert(sampOut.end(), out, out + 20);
            }
        }
        b.f32("blp_patches", patches);
        b.f32("blp_rays", rays);
        b.f32("blp_hits", hits);
        b.f32("blp_samp_in", samp);
        b.f32("blp_samp_out", sampOut);
    }
}

int main(int argc, char **argv) {
    if (argc < 3) { fprintf(stderr, "usage: ref_dump tables|golden <out.bin>\n"); return 1; }
    PBRTOptions opt;
    opt.quiet = true;
    InitPBRT(opt);
    if (!strcmp(argv[1], "tables")) DumpTables(argv[2]);
    else if (!strcmp(argv[1], "golden")) DumpGolden(argv[2]);
    else return 1;
    return 0;
}
