// materials.cuh -- Material::GetBxDF for the per-material shading kernels (one k_shade instantiation per
// material tag, mirroring the reference's per-material MaterialEvalQueue kernels, wavefront/surfscatter.cpp:26-55).
//   DiffuseMaterial::GetBxDF          materials.h:465-469
//   ConductorMaterial::GetBxDF        materials.h:491-513
//   DielectricMaterial::GetBxDF       materials.h:182-203
//   CoatedConductorMaterial::GetBxDF  materials.cpp:346-398
// GetBxDF<false>: every parameter is a constant (SpectrumConstantTexture / FloatConstantTexture), i.e. what BasicTextureEvaluator
// handles (textures.h:1158-1213).  GetBxDF<true>: reflectance / transmittance / roughness parameters may be image textures evaluated
// at the TextureEvalContext of the hit (UniversalTextureEvaluator).
#pragma once
#include "scene.h"
#include "lights.h"
#include "bxdfs.h"
#include "texture.h"

namespace bpt {

template <bool TEX> BPT_HD Spec MatSpectrum(const SpectrumRec &rec, int tex, const SceneView &sv, const Wavelengths &wl, const TexEvalCtx *tc) {
    if (TEX && tex >= 0) return EvalSpectrumTexture<kTexMaxDepth>(sv, tex, *tc, wl, TEXSPEC_ALBEDO);
    return EvalSpectrum(rec, sv.specPool, wl);
}
template <bool TEX> BPT_HD float MatFloat(float v, int tex, const SceneView &sv, const TexEvalCtx *tc) {
    if (TEX && tex >= 0) return EvalFloatTexture<kTexMaxDepth>(sv, tex, *tc);
    return v;
}

struct DiffuseMaterialEval {
    static constexpr int kTag = MAT_DIFFUSE;
    typedef DiffuseBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        Spec r = ClampSpec(MatSpectrum<TEX>(m.reflectance, m.reflectanceTex, sv, wl, tc), 0, 1);
        return DiffuseBxDF{r};
    }
};

BPT_HD float EvalSpectrumAt(const SpectrumRec &s, const float *pool, float lambda) {
    switch (s.kind) {
    case SPEC_CONSTANT: return s.c0;
    case SPEC_RGB_ALBEDO: return SigmoidPoly(s.c0, s.c1, s.c2, lambda);
    case SPEC_RGB_UNBOUNDED: return s.scale * SigmoidPoly(s.c0, s.c1, s.c2, lambda);
    case SPEC_DENSE: return DenseLookup(pool + s.table, lambda);
    case SPEC_PIECEWISE: return PiecewiseLinearEval(pool + s.table, pool + s.table + s.n, s.n, lambda);
    }
    return 0;
}

struct ConductorMaterialEval {
    static constexpr int kTag = MAT_CONDUCTOR;
    typedef ConductorBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        float uRough = MatFloat<TEX>(m.uroughness, m.uroughnessTex, sv, tc), vRough = MatFloat<TEX>(m.vroughness, m.vroughnessTex, sv, tc);
        if (m.flags & 1) { uRough = TRDist::RoughnessToAlpha(uRough); vRough = TRDist::RoughnessToAlpha(vRough); }
        Spec etas, ks;
        if (!m.useReflectance) {
            etas = EvalSpectrum(m.eta, sv.specPool, wl);
            ks = EvalSpectrum(m.k, sv.specPool, wl);
        } else {
            Spec r = ClampSpec(MatSpectrum<TEX>(m.reflectance, m.reflectanceTex, sv, wl, tc), 0, .9999f);
            etas = Spec(1.f);
            ks = 2 * SqrtSpec(r) / SqrtSpec(ClampZero(Spec(1) - r));
        }
        ConductorBxDF b;
        b.mfDistrib = TRDist(uRough, vRough); b.eta = etas; b.k = ks;
        return b;
    }
};

struct DielectricMaterialEval {
    static constexpr int kTag = MAT_DIELECTRIC;
    typedef DielectricBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        float sampledEta = EvalSpectrumAt(m.eta, sv.specPool, wl.lambda[0]);
        if (m.eta.kind != SPEC_CONSTANT) { wl.TerminateSecondary(); *lambdaChanged = true; }
        if (sampledEta == 0) sampledEta = 1;
        float urough = MatFloat<TEX>(m.uroughness, m.uroughnessTex, sv, tc), vrough = MatFloat<TEX>(m.vroughness, m.vroughnessTex, sv, tc);
        if (m.flags & 1) { urough = TRDist::RoughnessToAlpha(urough); vrough = TRDist::RoughnessToAlpha(vrough); }
        DielectricBxDF b;
        b.eta = sampledEta; b.mfDistrib = TRDist(urough, vrough);
        return b;
    }
};

struct CoatedConductorMaterialEval {
    static constexpr int kTag = MAT_COATED_CONDUCTOR;
    typedef CoatedConductorBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        float iurough = m.iface_uroughness, ivrough = m.iface_vroughness;
        if (m.flags & 1) { iurough = TRDist::RoughnessToAlpha(iurough); ivrough = TRDist::RoughnessToAlpha(ivrough); }
        TRDist interfaceDistrib(iurough, ivrough);
        float thick = m.thickness;
        float ieta = EvalSpectrumAt(m.iface_eta, sv.specPool, wl.lambda[0]);
        if (m.iface_eta.kind != SPEC_CONSTANT) { wl.TerminateSecondary(); *lambdaChanged = true; }
        if (ieta == 0) ieta = 1;
        Spec ce, ck;
        if (!m.useReflectance) {
            ce = EvalSpectrum(m.eta, sv.specPool, wl);
            ck = EvalSpectrum(m.k, sv.specPool, wl);
        } else {
            Spec r = ClampSpec(MatSpectrum<TEX>(m.reflectance, m.reflectanceTex, sv, wl, tc), 0, .9999f);
            ce = Spec(1.f);
            ck = 2 * SqrtSpec(r) / SqrtSpec(ClampZero(Spec(1) - r));
        }
        ce /= ieta;
        ck /= ieta;
        float curough = MatFloat<TEX>(m.uroughness, m.uroughnessTex, sv, tc), cvrough = MatFloat<TEX>(m.vroughness, m.vroughnessTex, sv, tc);
        if (m.flags & 1) { curough = TRDist::RoughnessToAlpha(curough); cvrough = TRDist::RoughnessToAlpha(cvrough); }
        TRDist conductorDistrib(curough, cvrough);
        Spec a = ClampSpec(EvalSpectrum(m.albedo, sv.specPool, wl), 0, 1);
        float gg = Clampf(m.g, -1, 1);
        CoatedConductorBxDF b;
        b.top.eta = ieta; b.top.mfDistrib = interfaceDistrib;
        b.bottom.mfDistrib = conductorDistrib; b.bottom.eta = ce; b.bottom.k = ck;
        b.thickness = stdmax(thick, FLT_MIN); b.albedo = a; b.g = gg; b.maxDepth = m.maxDepth; b.nSamples = m.nSamples;
        b.seed = m.optionsSeed;
        return b;
    }
};

// CoatedDiffuseMaterial::GetBxDF (materials.cpp:255-283)
struct CoatedDiffuseMaterialEval {
    static constexpr int kTag = MAT_COATED_DIFFUSE;
    typedef CoatedDiffuseBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        Spec r = ClampSpec(MatSpectrum<TEX>(m.reflectance, m.reflectanceTex, sv, wl, tc), 0, 1);
        float urough = MatFloat<TEX>(m.uroughness, m.uroughnessTex, sv, tc), vrough = MatFloat<TEX>(m.vroughness, m.vroughnessTex, sv, tc);
        if (m.flags & 1) { urough = TRDist::RoughnessToAlpha(urough); vrough = TRDist::RoughnessToAlpha(vrough); }
        TRDist distrib(urough, vrough);
        float thick = m.thickness;
        float sampledEta = EvalSpectrumAt(m.eta, sv.specPool, wl.lambda[0]);
        if (m.eta.kind != SPEC_CONSTANT) { wl.TerminateSecondary(); *lambdaChanged = true; }
        if (sampledEta == 0) sampledEta = 1;
        Spec a = ClampSpec(EvalSpectrum(m.albedo, sv.specPool, wl), 0, 1);
        float gg = Clampf(m.g, -1, 1);
        CoatedDiffuseBxDF b;
        b.top.eta = sampledEta; b.top.mfDistrib = distrib;
        b.bottom.R = r;
        b.thickness = stdmax(thick, FLT_MIN); b.albedo = a; b.g = gg; b.maxDepth = m.maxDepth; b.nSamples = m.nSamples;
        b.seed = m.optionsSeed;
        return b;
    }
};

// DiffuseTransmissionMaterial::GetBxDF (materials.h:815-821)
struct DiffuseTransmissionMaterialEval {
    static constexpr int kTag = MAT_DIFFUSE_TRANSMISSION;
    typedef DiffuseTransmissionBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        Spec r = ClampSpec(MatSpectrum<TEX>(m.reflectance, m.reflectanceTex, sv, wl, tc) * m.scale, 0, 1);
        Spec t = ClampSpec(MatSpectrum<TEX>(m.transmittance, m.transmittanceTex, sv, wl, tc) * m.scale, 0, 1);
        return DiffuseTransmissionBxDF{r, t};
    }
};

// ThinDielectricMaterial::GetBxDF (materials.h:226-240)
struct ThinDielectricMaterialEval {
    static constexpr int kTag = MAT_THIN_DIELECTRIC;
    typedef ThinDielectricBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged, const TexEvalCtx *tc) {
        float sampledEta = EvalSpectrumAt(m.eta, sv.specPool, wl.lambda[0]);
        if (m.eta.kind != SPEC_CONSTANT) { wl.TerminateSecondary(); *lambdaChanged = true; }
        if (sampledEta == 0) sampledEta = 1;
        return ThinDielectricBxDF{sampledEta};
    }
};

struct InterfaceMaterialEval {
    static constexpr int kTag = MAT_INTERFACE;
    typedef DiffuseBxDF BxDF;
    template <bool TEX> BPT_HD static BxDF GetBxDF(const DeviceMaterial &, const SceneView &, Wavelengths &, bool *, const TexEvalCtx *) { return DiffuseBxDF{Spec(0.f)}; }
};

}  // namespace bpt
