// bxdfs.h -- BxDFs evaluated by the per-material shading kernels, in the local shading frame.
//   BxDFFlags / BSDFSample           base/bxdf.h:20-140
//   BSDF (frame wrapper)             bsdf.h
//   DiffuseBxDF                      bxdfs.h:30-82
//   ConductorBxDF                    bxdfs.h:280-383
//   DielectricBxDF                   bxdfs.cpp:77-258
//   TrowbridgeReitzDistribution, FrDielectric, FrComplex   util/scattering.h:61-214
//   LayeredBxDF / CoatedConductorBxDF   bxdfs.h:431-912
#pragma once
#include "spectrum.h"
#include "sampling.h"

namespace bpt {

enum BxDFFlags {
    BXDF_UNSET = 0, BXDF_REFLECTION = 1, BXDF_TRANSMISSION = 2, BXDF_DIFFUSE = 4, BXDF_GLOSSY = 8, BXDF_SPECULAR = 16,
    BXDF_DIFFUSE_REFLECTION = 5, BXDF_DIFFUSE_TRANSMISSION = 6, BXDF_GLOSSY_REFLECTION = 9, BXDF_GLOSSY_TRANSMISSION = 10,
    BXDF_SPECULAR_REFLECTION = 17, BXDF_SPECULAR_TRANSMISSION = 18, BXDF_ALL = 31
};
enum { REFLTRANS_REFLECTION = 1, REFLTRANS_TRANSMISSION = 2, REFLTRANS_ALL = 3 };
enum TransportMode { MODE_RADIANCE = 0, MODE_IMPORTANCE = 1 };
BPT_HD bool IsReflective(int f) { return f & BXDF_REFLECTION; }
BPT_HD bool IsTransmissive(int f) { return f & BXDF_TRANSMISSION; }
BPT_HD bool IsSpecular(int f) { return f & BXDF_SPECULAR; }
BPT_HD bool IsNonSpecular(int f) { return f & (BXDF_DIFFUSE | BXDF_GLOSSY); }

struct BSDFSample {
    Spec f; Vec3 wi; float pdf; int flags; float eta; bool pdfIsProportional; bool valid;
    BPT_HD BSDFSample() : pdf(0), flags(0), eta(1), pdfIsProportional(false), valid(false) {}
    BPT_HD BSDFSample(Spec f_, Vec3 wi_, float pdf_, int flags_, float eta_ = 1, bool prop = false)
        : f(f_), wi(wi_), pdf(pdf_), flags(flags_), eta(eta_), pdfIsProportional(prop), valid(true) {}
};

// spherical geometry helpers (util/vecmath.h:1649-1700)
BPT_HD float CosTheta(Vec3 w) { return w.z; }
BPT_HD float Cos2Theta(Vec3 w) { return Sqr(w.z); }
BPT_HD float AbsCosTheta(Vec3 w) { return fabsf(w.z); }
BPT_HD float Sin2Theta(Vec3 w) { return stdmax(0.f, 1 - Cos2Theta(w)); }
BPT_HD float SinTheta(Vec3 w) { return sqrtf(Sin2Theta(w)); }
BPT_HD float TanTheta(Vec3 w) { return SinTheta(w) / CosTheta(w); }
BPT_HD float Tan2Theta(Vec3 w) { return Sin2Theta(w) / Cos2Theta(w); }
BPT_HD float CosPhi(Vec3 w) { float s = SinTheta(w); return (s == 0) ? 1 : Clampf(w.x / s, -1, 1); }
BPT_HD float SinPhi(Vec3 w) { float s = SinTheta(w); return (s == 0) ? 0 : Clampf(w.y / s, -1, 1); }
BPT_HD bool SameHemisphere(Vec3 w, Vec3 wp) { return w.z * wp.z > 0; }

// ---------------------------------------------------------------- DiffuseBxDF
struct DiffuseBxDF {
    Spec R;
    BPT_HD int Flags() const { return R.IsNonZero() ? BXDF_DIFFUSE_REFLECTION : BXDF_UNSET; }
    BPT_HD Spec f(Vec3 wo, Vec3 wi, int mode) const {
        if (!SameHemisphere(wo, wi)) return Spec(0.f);
        return R * kInvPi;
    }
    BPT_HD BSDFSample Sample_f(Vec3 wo, float uc, Vec2 u, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (!(sampleFlags & REFLTRANS_REFLECTION)) return {};
        Vec3 wi = SampleCosineHemisphere(u);
        if (wo.z < 0) wi.z *= -1;
        float pdf = CosineHemispherePDF(AbsCosTheta(wi));
        return BSDFSample(R * kInvPi, wi, pdf, BXDF_DIFFUSE_REFLECTION);
    }
    BPT_HD float PDF(Vec3 wo, Vec3 wi, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (!(sampleFlags & REFLTRANS_REFLECTION) || !SameHemisphere(wo, wi)) return 0;
        return CosineHemispherePDF(AbsCosTheta(wi));
    }
    BPT_HD void Regularize() {}
};

// ---------------------------------------------------------------- BSDF frame wrapper (bsdf.h)
template <typename BxDF>
struct BSDF {
    BxDF bxdf;
    Frame frame;
    BPT_HD BSDF(Nrm3 ns, Vec3 dpdus, const BxDF &b) : bxdf(b), frame(Frame::FromXZ(Normalize(dpdus), V(ns))) {}
    BPT_HD int Flags() const { return bxdf.Flags(); }
    BPT_HD Vec3 RenderToLocal(Vec3 v) const { return frame.ToLocal(v); }
    BPT_HD Vec3 LocalToRender(Vec3 v) const { return frame.FromLocal(v); }
    BPT_HD Spec f(Vec3 woRender, Vec3 wiRender, int mode = MODE_RADIANCE) const {
        Vec3 wi = RenderToLocal(wiRender), wo = RenderToLocal(woRender);
        if (wo.z == 0) return Spec(0.f);
        return bxdf.f(wo, wi, mode);
    }
    BPT_HD BSDFSample Sample_f(Vec3 woRender, float u, Vec2 u2, int mode = MODE_RADIANCE, int sampleFlags = REFLTRANS_ALL) const {
        Vec3 wo = RenderToLocal(woRender);
        if (wo.z == 0 || !(bxdf.Flags() & sampleFlags)) return {};
        BSDFSample bs = bxdf.Sample_f(wo, u, u2, mode, sampleFlags);
        if (!bs.valid || !bs.f.IsNonZero() || bs.pdf == 0 || bs.wi.z == 0) return {};
        bs.wi = LocalToRender(bs.wi);
        return bs;
    }
    BPT_HD float PDF(Vec3 woRender, Vec3 wiRender, int mode = MODE_RADIANCE, int sampleFlags = REFLTRANS_ALL) const {
        Vec3 wo = RenderToLocal(woRender), wi = RenderToLocal(wiRender);
        if (wo.z == 0) return 0;
        return bxdf.PDF(wo, wi, mode, sampleFlags);
    }
    BPT_HD void Regularize() { bxdf.Regularize(); }
};

}  // namespace bpt
