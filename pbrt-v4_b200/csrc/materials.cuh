// materials.cuh -- Material::GetBxDF for the per-material shading kernels (one k_shade instantiation per
// material tag, mirroring the reference's per-material MaterialEvalQueue kernels, wavefront/surfscatter.cpp:26-55).
//   DiffuseMaterial::GetBxDF          materials.h:465-469
//   ConductorMaterial::GetBxDF        materials.h:491-513
//   DielectricMaterial::GetBxDF       materials.h:182-203
//   CoatedConductorMaterial::GetBxDF  materials.cpp:346-398
// Textures on this path are constant (SpectrumConstantTexture / FloatConstantTexture), i.e. what
// BasicTextureEvaluator handles (textures.h:1158-1213).
#pragma once
#include "scene.h"
#include "lights.h"
#include "bxdfs.h"

namespace bpt {

struct DiffuseMaterialEval {
    static constexpr int kTag = MAT_DIFFUSE;
    typedef DiffuseBxDF BxDF;
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        Spec r = ClampSpec(EvalSpectrum(m.reflectance, sv.specPool, wl), 0, 1);
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
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        float uRough = m.uroughness, vRough = m.vroughness;
        if (m.flags & 1) { uRough = TRDist::RoughnessToAlpha(uRough); vRough = TRDist::RoughnessToAlpha(vRough); }
        Spec etas, ks;
        if (!m.useReflectance) {
            etas = EvalSpectrum(m.eta, sv.specPool, wl);
            ks = EvalSpectrum(m.k, sv.specPool, wl);
        } else {
            Spec r = ClampSpec(EvalSpectrum(m.reflectance, sv.specPool, wl), 0, .9999f);
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
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        float sampledEta = EvalSpectrumAt(m.eta, sv.specPool, wl.lambda[0]);
        if (m.eta.kind != SPEC_CONSTANT) { wl.TerminateSecondary(); *lambdaChanged = true; }
        if (sampledEta == 0) sampledEta = 1;
        float urough = m.uroughness, vrough = m.vroughness;
        if (m.flags & 1) { urough = TRDist::RoughnessToAlpha(urough); vrough = TRDist::RoughnessToAlpha(vrough); }
        DielectricBxDF b;
        b.eta = sampledEta; b.mfDistrib = TRDist(urough, vrough);
        return b;
    }
};

struct CoatedConductorMaterialEval {
    static constexpr int kTag = MAT_COATED_CONDUCTOR;
    typedef CoatedConductorBxDF BxDF;
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
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
            Spec r = ClampSpec(EvalSpectrum(m.reflectance, sv.specPool, wl), 0, .9999f);
            ce = Spec(1.f);
            ck = 2 * SqrtSpec(r) / SqrtSpec(ClampZero(Spec(1) - r));
        }
        ce /= ieta;
        ck /= ieta;
        float curough = m.uroughness, cvrough = m.vroughness;
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
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        Spec r = ClampSpec(EvalSpectrum(m.reflectance, sv.specPool, wl), 0, 1);
        float urough = m.uroughness, vrough = m.vroughness;
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
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        Spec r = ClampSpec(EvalSpectrum(m.reflectance, sv.specPool, wl) * m.scale, 0, 1);
        Spec t = ClampSpec(EvalSpectrum(m.transmittance, sv.specPool, wl) * m.scale, 0, 1);
        return DiffuseTransmissionBxDF{r, t};
    }
};

// ThinDielectricMaterial::GetBxDF (materials.h:226-240)
struct ThinDielectricMaterialEval {
    static constexpr int kTag = MAT_THIN_DIELECTRIC;
    typedef ThinDielectricBxDF BxDF;
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        float sampledEta = EvalSpectrumAt(m.eta, sv.specPool, wl.lambda[0]);
        if (m.eta.kind != SPEC_CONSTANT) { wl.TerminateSecondary(); *lambdaChanged = true; }
        if (sampledEta == 0) sampledEta = 1;
        return ThinDielectricBxDF{sampledEta};
    }
};

struct InterfaceMaterialEval {
    static constexpr int kTag = MAT_INTERFACE;
    typedef DiffuseBxDF BxDF;
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &, const SceneView &, Wavelengths &, bool *) { return DiffuseBxDF{Spec(0.f)}; }
};

}  // namespace bpt
