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
// BPT_OUTLINE_BXDF: the microfacet BxDFs and the LayeredBxDF random walk as real device functions.  ncu r02g: the coated-conductor
// shading kernel (77 k instructions with everything inlined: LayeredBxDF::f / Sample_f / PDF each inline Dielectric and Conductor
// f / Sample_f / PDF at every interface visit) stalls 14.9 cycles per issued instruction on instruction fetch.
#ifndef BPT_OUTLINE_BXDF
#define BPT_OUTLINE_BXDF 1
#endif
#if BPT_OUTLINE_BXDF && defined(__CUDACC__)
#define BPT_BXDF_FN __host__ __device__ __noinline__
#else
#define BPT_BXDF_FN BPT_HD
#endif
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

// ---------------------------------------------------------------- scattering helpers (util/scattering.h)
BPT_HD Vec3 Reflect(Vec3 wo, Vec3 n) { return -wo + 2 * Dot(wo, n) * n; }
BPT_HD bool Refract(Vec3 wi, Nrm3 n, float eta, float *etap, Vec3 *wt) {
    float cosTheta_i = Dot(n, wi);
    if (cosTheta_i < 0) { eta = 1 / eta; cosTheta_i = -cosTheta_i; n = -n; }
    float sin2Theta_i = stdmax(0.f, 1 - Sqr(cosTheta_i));
    float sin2Theta_t = sin2Theta_i / Sqr(eta);
    if (sin2Theta_t >= 1) return false;
    float cosTheta_t = sqrtf(1 - sin2Theta_t);
    *wt = -wi / eta + (cosTheta_i / eta - cosTheta_t) * V(n);
    if (etap) *etap = eta;
    return true;
}
BPT_HD float FrDielectric(float cosTheta_i, float eta) {
    cosTheta_i = Clampf(cosTheta_i, -1, 1);
    if (cosTheta_i < 0) { eta = 1 / eta; cosTheta_i = -cosTheta_i; }
    float sin2Theta_i = 1 - Sqr(cosTheta_i);
    float sin2Theta_t = sin2Theta_i / Sqr(eta);
    if (sin2Theta_t >= 1) return 1.f;
    float cosTheta_t = SafeSqrt(1 - sin2Theta_t);
    float r_parl = (eta * cosTheta_i - cosTheta_t) / (eta * cosTheta_i + cosTheta_t);
    float r_perp = (cosTheta_i - eta * cosTheta_t) / (cosTheta_i + eta * cosTheta_t);
    return (Sqr(r_parl) + Sqr(r_perp)) / 2;
}
// pstd::complex<float> (util/pstd.h:1066-1229), including its implicit float -> complex promotions
struct Cplx {
    float re, im;
    BPT_HD Cplx(float r) : re(r), im(0) {}
    BPT_HD Cplx(float r, float i) : re(r), im(i) {}
    BPT_HD Cplx operator+(Cplx z) const { return {re + z.re, im + z.im}; }
    BPT_HD Cplx operator-(Cplx z) const { return {re - z.re, im - z.im}; }
    BPT_HD Cplx operator*(Cplx z) const { return {re * z.re - im * z.im, re * z.im + im * z.re}; }
    BPT_HD Cplx operator/(Cplx z) const {
        float scale = 1 / (z.re * z.re + z.im * z.im);
        return {scale * (re * z.re + im * z.im), scale * (im * z.re - re * z.im)};
    }
};
BPT_HD float CNorm(Cplx z) { return z.re * z.re + z.im * z.im; }
BPT_HD Cplx CSqrt(Cplx z) {
    float n = sqrtf(CNorm(z)), t1 = sqrtf(.5f * (n + fabsf(z.re))), t2 = .5f * z.im / t1;
    if (n == 0) return Cplx(0);
    if (z.re >= 0) return {t1, t2};
    return {fabsf(t2), copysignf(t1, z.im)};
}
BPT_HD float FrComplex(float cosTheta_i, Cplx eta) {
    cosTheta_i = Clampf(cosTheta_i, 0, 1);
    float sin2Theta_i = 1 - Sqr(cosTheta_i);
    Cplx sin2Theta_t = Cplx(sin2Theta_i) / (eta * eta);
    Cplx cosTheta_t = CSqrt(Cplx(1) - sin2Theta_t);
    Cplx r_parl = (eta * Cplx(cosTheta_i) - cosTheta_t) / (eta * Cplx(cosTheta_i) + cosTheta_t);
    Cplx r_perp = (Cplx(cosTheta_i) - eta * cosTheta_t) / (Cplx(cosTheta_i) + eta * cosTheta_t);
    return (CNorm(r_parl) + CNorm(r_perp)) / 2;
}
BPT_HD Spec FrComplexSpec(float cosTheta_i, Spec eta, Spec k) {
    Spec r;
    for (int i = 0; i < 4; ++i) r[i] = FrComplex(cosTheta_i, Cplx(eta[i], k[i]));
    return r;
}
BPT_HD Vec2 SampleUniformDiskPolar(Vec2 u) {
    float r = sqrtf(u.x);
    float theta = 2 * kPi * u.y;
    return {r * CosF(theta), r * SinF(theta)};
}

// TrowbridgeReitzDistribution (util/scattering.h:109-214)
struct TRDist {
    float alpha_x, alpha_y;
    BPT_HD TRDist() : alpha_x(0), alpha_y(0) {}
    BPT_HD TRDist(float ax, float ay) : alpha_x(ax), alpha_y(ay) {
        if (!EffectivelySmooth()) { alpha_x = stdmax(alpha_x, 1e-4f); alpha_y = stdmax(alpha_y, 1e-4f); }
    }
    BPT_HD bool EffectivelySmooth() const { return stdmax(alpha_x, alpha_y) < 1e-3f; }
    BPT_HD float D(Vec3 wm) const {
        float tan2Theta = Tan2Theta(wm);
        if (IsInf_(tan2Theta)) return 0;
        float cos4Theta = Sqr(Cos2Theta(wm));
        if (cos4Theta < 1e-16f) return 0;
        float e = tan2Theta * (Sqr(CosPhi(wm) / alpha_x) + Sqr(SinPhi(wm) / alpha_y));
        return 1 / (kPi * alpha_x * alpha_y * cos4Theta * Sqr(1 + e));
    }
    BPT_HD float Lambda(Vec3 w) const {
        float tan2Theta = Tan2Theta(w);
        if (IsInf_(tan2Theta)) return 0;
        float alpha2 = Sqr(CosPhi(w) * alpha_x) + Sqr(SinPhi(w) * alpha_y);
        return (sqrtf(1 + alpha2 * tan2Theta) - 1) / 2;
    }
    BPT_HD float G1(Vec3 w) const { return 1 / (1 + Lambda(w)); }
    BPT_HD float G(Vec3 wo, Vec3 wi) const { return 1 / (1 + Lambda(wo) + Lambda(wi)); }
    BPT_HD float D(Vec3 w, Vec3 wm) const { return G1(w) / AbsCosTheta(w) * D(wm) * AbsDot(w, wm); }
    BPT_HD float PDF(Vec3 w, Vec3 wm) const { return D(w, wm); }
    BPT_HD Vec3 Sample_wm(Vec3 w, Vec2 u) const {
        Vec3 wh = Normalize(Vec3(alpha_x * w.x, alpha_y * w.y, w.z));
        if (wh.z < 0) wh = -wh;
        Vec3 T1 = (wh.z < 0.99999f) ? Normalize(Cross(Vec3(0, 0, 1), wh)) : Vec3(1, 0, 0);
        Vec3 T2 = Cross(wh, T1);
        Vec2 p = SampleUniformDiskPolar(u);
        float h = sqrtf(1 - Sqr(p.x));
        p.y = Lerp((1 + wh.z) / 2, h, p.y);
        float pz = sqrtf(stdmax(0.f, 1 - (Sqr(p.x) + Sqr(p.y))));
        Vec3 nh = p.x * T1 + p.y * T2 + pz * wh;
        return Normalize(Vec3(alpha_x * nh.x, alpha_y * nh.y, stdmax(1e-6f, nh.z)));
    }
    BPT_HD static float RoughnessToAlpha(float roughness) { return sqrtf(roughness); }
    BPT_HD void Regularize() {
        if (alpha_x < 0.3f) alpha_x = Clampf(2 * alpha_x, 0.1f, 0.3f);
        if (alpha_y < 0.3f) alpha_y = Clampf(2 * alpha_y, 0.1f, 0.3f);
    }
};

// ---------------------------------------------------------------- ConductorBxDF (bxdfs.h:280-383)
struct ConductorBxDF {
    TRDist mfDistrib;
    Spec eta, k;
    BPT_HD int Flags() const { return mfDistrib.EffectivelySmooth() ? BXDF_SPECULAR_REFLECTION : BXDF_GLOSSY_REFLECTION; }
    BPT_BXDF_FN BSDFSample Sample_f(Vec3 wo, float uc, Vec2 u, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (!(sampleFlags & REFLTRANS_REFLECTION)) return {};
        if (mfDistrib.EffectivelySmooth()) {
            Vec3 wi(-wo.x, -wo.y, wo.z);
            Spec f = FrComplexSpec(AbsCosTheta(wi), eta, k) / AbsCosTheta(wi);
            return BSDFSample(f, wi, 1, BXDF_SPECULAR_REFLECTION);
        }
        if (wo.z == 0) return {};
        Vec3 wm = mfDistrib.Sample_wm(wo, u);
        Vec3 wi = Reflect(wo, wm);
        if (!SameHemisphere(wo, wi)) return {};
        float pdf = mfDistrib.PDF(wo, wm) / (4 * AbsDot(wo, wm));
        float cosTheta_o = AbsCosTheta(wo), cosTheta_i = AbsCosTheta(wi);
        if (cosTheta_i == 0 || cosTheta_o == 0) return {};
        Spec F = FrComplexSpec(AbsDot(wo, wm), eta, k);
        Spec f = mfDistrib.D(wm) * F * mfDistrib.G(wo, wi) / (4 * cosTheta_i * cosTheta_o);
        return BSDFSample(f, wi, pdf, BXDF_GLOSSY_REFLECTION);
    }
    BPT_BXDF_FN Spec f(Vec3 wo, Vec3 wi, int mode) const {
        if (!SameHemisphere(wo, wi)) return Spec(0.f);
        if (mfDistrib.EffectivelySmooth()) return Spec(0.f);
        float cosTheta_o = AbsCosTheta(wo), cosTheta_i = AbsCosTheta(wi);
        if (cosTheta_i == 0 || cosTheta_o == 0) return Spec(0.f);
        Vec3 wm = wi + wo;
        if (LengthSquared(wm) == 0) return Spec(0.f);
        wm = Normalize(wm);
        Spec F = FrComplexSpec(AbsDot(wo, wm), eta, k);
        return mfDistrib.D(wm) * F * mfDistrib.G(wo, wi) / (4 * cosTheta_i * cosTheta_o);
    }
    BPT_BXDF_FN float PDF(Vec3 wo, Vec3 wi, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (!(sampleFlags & REFLTRANS_REFLECTION)) return 0;
        if (!SameHemisphere(wo, wi)) return 0;
        if (mfDistrib.EffectivelySmooth()) return 0;
        Vec3 wm = wo + wi;
        if (LengthSquared(wm) == 0) return 0;
        wm = FaceForward(Normalize(wm), Nrm3(0, 0, 1));
        return mfDistrib.PDF(wo, wm) / (4 * AbsDot(wo, wm));
    }
    BPT_HD void Regularize() { mfDistrib.Regularize(); }
};

// ---------------------------------------------------------------- DielectricBxDF (bxdfs.cpp:77-258)
struct DielectricBxDF {
    float eta;
    TRDist mfDistrib;
    BPT_HD int Flags() const {
        int flags = (eta == 1) ? BXDF_TRANSMISSION : (BXDF_REFLECTION | BXDF_TRANSMISSION);
        return flags | (mfDistrib.EffectivelySmooth() ? BXDF_SPECULAR : BXDF_GLOSSY);
    }
    BPT_BXDF_FN BSDFSample Sample_f(Vec3 wo, float uc, Vec2 u, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (eta == 1 || mfDistrib.EffectivelySmooth()) {
            float R = FrDielectric(CosTheta(wo), eta), T = 1 - R;
            float pr = R, pt = T;
            if (!(sampleFlags & REFLTRANS_REFLECTION)) pr = 0;
            if (!(sampleFlags & REFLTRANS_TRANSMISSION)) pt = 0;
            if (pr == 0 && pt == 0) return {};
            if (uc < pr / (pr + pt)) {
                Vec3 wi(-wo.x, -wo.y, wo.z);
                Spec fr(R / AbsCosTheta(wi));
                return BSDFSample(fr, wi, pr / (pr + pt), BXDF_SPECULAR_REFLECTION);
            } else {
                Vec3 wi; float etap;
                bool valid = Refract(wo, Nrm3(0, 0, 1), eta, &etap, &wi);
                if (!valid) return {};
                Spec ft(T / AbsCosTheta(wi));
                if (mode == MODE_RADIANCE) ft /= Sqr(etap);
                return BSDFSample(ft, wi, pt / (pr + pt), BXDF_SPECULAR_TRANSMISSION, etap);
            }
        } else {
            Vec3 wm = mfDistrib.Sample_wm(wo, u);
            float R = FrDielectric(Dot(wo, wm), eta);
            float T = 1 - R;
            float pr = R, pt = T;
            if (!(sampleFlags & REFLTRANS_REFLECTION)) pr = 0;
            if (!(sampleFlags & REFLTRANS_TRANSMISSION)) pt = 0;
            if (pr == 0 && pt == 0) return {};
            float pdf;
            if (uc < pr / (pr + pt)) {
                Vec3 wi = Reflect(wo, wm);
                if (!SameHemisphere(wo, wi)) return {};
                pdf = mfDistrib.PDF(wo, wm) / (4 * AbsDot(wo, wm)) * pr / (pr + pt);
                Spec f(mfDistrib.D(wm) * mfDistrib.G(wo, wi) * R / (4 * CosTheta(wi) * CosTheta(wo)));
                return BSDFSample(f, wi, pdf, BXDF_GLOSSY_REFLECTION);
            } else {
                float etap; Vec3 wi;
                bool tir = !Refract(wo, Nrm3(wm), eta, &etap, &wi);
                if (tir) return {};   // (wi is uninitialised in the reference when tir; the test order there still returns {})
                if (SameHemisphere(wo, wi) || wi.z == 0) return {};
                float denom = Sqr(Dot(wi, wm) + Dot(wo, wm) / etap);
                float dwm_dwi = AbsDot(wi, wm) / denom;
                pdf = mfDistrib.PDF(wo, wm) * dwm_dwi * pt / (pr + pt);
                Spec ft(T * mfDistrib.D(wm) * mfDistrib.G(wo, wi) * fabsf(Dot(wi, wm) * Dot(wo, wm) / (CosTheta(wi) * CosTheta(wo) * denom)));
                if (mode == MODE_RADIANCE) ft /= Sqr(etap);
                return BSDFSample(ft, wi, pdf, BXDF_GLOSSY_TRANSMISSION, etap);
            }
        }
    }
    BPT_BXDF_FN Spec f(Vec3 wo, Vec3 wi, int mode) const {
        if (eta == 1 || mfDistrib.EffectivelySmooth()) return Spec(0.f);
        float cosTheta_o = CosTheta(wo), cosTheta_i = CosTheta(wi);
        bool reflect = cosTheta_i * cosTheta_o > 0;
        float etap = 1;
        if (!reflect) etap = cosTheta_o > 0 ? eta : (1 / eta);
        Vec3 wm = wi * etap + wo;
        if (cosTheta_i == 0 || cosTheta_o == 0 || LengthSquared(wm) == 0) return Spec(0.f);
        wm = FaceForward(Normalize(wm), Nrm3(0, 0, 1));
        if (Dot(wm, wi) * cosTheta_i < 0 || Dot(wm, wo) * cosTheta_o < 0) return Spec(0.f);
        float F = FrDielectric(Dot(wo, wm), eta);
        if (reflect) return Spec(mfDistrib.D(wm) * mfDistrib.G(wo, wi) * F / fabsf(4 * cosTheta_i * cosTheta_o));
        float denom = Sqr(Dot(wi, wm) + Dot(wo, wm) / etap) * cosTheta_i * cosTheta_o;
        float ft = mfDistrib.D(wm) * (1 - F) * mfDistrib.G(wo, wi) * fabsf(Dot(wi, wm) * Dot(wo, wm) / denom);
        if (mode == MODE_RADIANCE) ft /= Sqr(etap);
        return Spec(ft);
    }
    BPT_BXDF_FN float PDF(Vec3 wo, Vec3 wi, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (eta == 1 || mfDistrib.EffectivelySmooth()) return 0;
        float cosTheta_o = CosTheta(wo), cosTheta_i = CosTheta(wi);
        bool reflect = cosTheta_i * cosTheta_o > 0;
        float etap = 1;
        if (!reflect) etap = cosTheta_o > 0 ? eta : (1 / eta);
        Vec3 wm = wi * etap + wo;
        if (cosTheta_i == 0 || cosTheta_o == 0 || LengthSquared(wm) == 0) return 0;
        wm = FaceForward(Normalize(wm), Nrm3(0, 0, 1));
        if (Dot(wm, wi) * cosTheta_i < 0 || Dot(wm, wo) * cosTheta_o < 0) return 0;
        float R = FrDielectric(Dot(wo, wm), eta);
        float T = 1 - R;
        float pr = R, pt = T;
        if (!(sampleFlags & REFLTRANS_REFLECTION)) pr = 0;
        if (!(sampleFlags & REFLTRANS_TRANSMISSION)) pt = 0;
        if (pr == 0 && pt == 0) return 0;
        float pdf;
        if (reflect) pdf = mfDistrib.PDF(wo, wm) / (4 * AbsDot(wo, wm)) * pr / (pr + pt);
        else {
            float denom = Sqr(Dot(wi, wm) + Dot(wo, wm) / etap);
            float dwm_dwi = AbsDot(wi, wm) / denom;
            pdf = mfDistrib.PDF(wo, wm) * dwm_dwi * pt / (pr + pt);
        }
        return pdf;
    }
    BPT_HD void Regularize() { mfDistrib.Regularize(); }
};

// ---------------------------------------------------------------- Henyey-Greenstein (media.h:43-71, util/scattering.h, util/sampling.cpp)
BPT_HD float HenyeyGreenstein(float cosTheta, float g) {
    g = Clampf(g, -.99f, .99f);
    float denom = 1 + Sqr(g) + 2 * g * cosTheta;
    return kInv4Pi * (1 - Sqr(g)) / (denom * SafeSqrt(denom));
}
BPT_HD Vec3 SampleHenyeyGreenstein(Vec3 wo, float g, Vec2 u, float *pdf) {
    g = Clampf(g, -.99f, .99f);
    float cosTheta;
    if (fabsf(g) < 1e-3f) cosTheta = 1 - 2 * u.x;
    else cosTheta = -1 / (2 * g) * (1 + Sqr(g) - Sqr((1 - Sqr(g)) / (1 + g - 2 * g * u.x)));
    float sinTheta = SafeSqrt(1 - Sqr(cosTheta));
    float phi = 2 * kPi * u.y;
    Frame wFrame = Frame::FromZ(wo);
    Vec3 local(Clampf(sinTheta, -1, 1) * CosF(phi), Clampf(sinTheta, -1, 1) * SinF(phi), Clampf(cosTheta, -1, 1));
    Vec3 wi = wFrame.FromLocal(local);
    *pdf = HenyeyGreenstein(cosTheta, g);
    return wi;
}
BPT_HD float PowerHeuristic(int nf, float fPdf, int ng, float gPdf) {
    float f = nf * fPdf, g = ng * gPdf;
    if (IsInf_(Sqr(f))) return 1;
    return Sqr(f) / (Sqr(f) + Sqr(g));
}
BPT_HD float SampleExponential(float u, float a) { return -LogF(1 - u) / a; }
BPT_HD uint64_t Hash_i32_f32x3(int s, Vec3 v) {   // Hash(int seed, Vector3f)
    uint64_t w[2] = {(uint64_t)(uint32_t)s | ((uint64_t)f2b(v.x) << 32), (uint64_t)f2b(v.y) | ((uint64_t)f2b(v.z) << 32)};
    return MurmurHash64A(w, 16, 0);
}

// ---------------------------------------------------------------- LayeredBxDF (bxdfs.h:431-900), twoSided = true instantiation
// Argument evaluation order: the reference writes e.g. `Sample_f(wo, uc, Point2f(r(), r()), ...)`; the order in which
// the two r() calls run is unspecified in C++ and fixed by the compiler that built the reference (gcc evaluates the
// arguments of a call right to left).  BPT_R2(a, b) draws in that order: b first, then a.
#define BPT_R2(a, b) float b = r(); float a = r();
template <typename TopBxDF, typename BottomBxDF>
struct LayeredBxDF {
    TopBxDF top;
    BottomBxDF bottom;
    float thickness, g;
    Spec albedo;
    int maxDepth, nSamples;
    int seed;   // Options->seed

    BPT_HD static float Tr(float dz, Vec3 w) {
        if (fabsf(dz) <= FLT_MIN) return 1;
        return FastExp(-fabsf(dz / w.z));
    }
    BPT_HD void Regularize() { top.Regularize(); bottom.Regularize(); }
    BPT_HD int Flags() const {
        int topFlags = top.Flags(), bottomFlags = bottom.Flags();
        int flags = BXDF_REFLECTION;
        if (IsSpecular(topFlags)) flags |= BXDF_SPECULAR;
        if ((topFlags & BXDF_DIFFUSE) || (bottomFlags & BXDF_DIFFUSE) || albedo.IsNonZero()) flags |= BXDF_DIFFUSE;
        else if ((topFlags & BXDF_GLOSSY) || (bottomFlags & BXDF_GLOSSY)) flags |= BXDF_GLOSSY;
        if (IsTransmissive(topFlags) && IsTransmissive(bottomFlags)) flags |= BXDF_TRANSMISSION;
        return flags;
    }
    // interface selectors (TopOrBottomBxDF): which = 0 top, 1 bottom
    BPT_HD Spec If(int which, Vec3 wo, Vec3 wi, int mode) const { return which == 0 ? top.f(wo, wi, mode) : bottom.f(wo, wi, mode); }
    BPT_HD BSDFSample ISample(int which, Vec3 wo, float uc, Vec2 u, int mode, int sf = REFLTRANS_ALL) const {
        return which == 0 ? top.Sample_f(wo, uc, u, mode, sf) : bottom.Sample_f(wo, uc, u, mode, sf);
    }
    BPT_HD float IPDF(int which, Vec3 wo, Vec3 wi, int mode, int sf = REFLTRANS_ALL) const {
        return which == 0 ? top.PDF(wo, wi, mode, sf) : bottom.PDF(wo, wi, mode, sf);
    }
    BPT_HD int IFlags(int which) const { return which == 0 ? top.Flags() : bottom.Flags(); }
    BPT_HD static bool Bad(const BSDFSample &b) { return !b.valid || !b.f.IsNonZero() || b.pdf == 0 || b.wi.z == 0; }

    BPT_BXDF_FN Spec f(Vec3 wo, Vec3 wi, int mode) const {
        Spec fsum(0.f);
        if (wo.z < 0) { wo = -wo; wi = -wi; }
        const bool enteredTop = true;
        int enterI = 0;
        int exitI, nonExitI;
        if (SameHemisphere(wo, wi) ^ enteredTop) { exitI = 1; nonExitI = 0; } else { exitI = 0; nonExitI = 1; }
        float exitZ = (SameHemisphere(wo, wi) ^ enteredTop) ? 0 : thickness;
        if (SameHemisphere(wo, wi)) fsum = nSamples * If(enterI, wo, wi, mode);
        RNG rng(Hash_i32_f32x3(seed, wo), Hash_f32x3(wi.x, wi.y, wi.z));
        auto r = [&rng]() { return stdmin(rng.UniformFloat(), kOneMinusEpsilon); };
        for (int s = 0; s < nSamples; ++s) {
            float uc = r();
            BPT_R2(ua, ub)
            BSDFSample wos = ISample(enterI, wo, uc, Vec2(ua, ub), mode, REFLTRANS_TRANSMISSION);
            if (Bad(wos)) continue;
            uc = r();
            BPT_R2(uc2, ud2)
            BSDFSample wis = ISample(exitI, wi, uc, Vec2(uc2, ud2), !mode, REFLTRANS_TRANSMISSION);
            if (Bad(wis)) continue;
            Spec beta = wos.f * AbsCosTheta(wos.wi) / wos.pdf;
            float z = enteredTop ? thickness : 0;
            Vec3 w = wos.wi;
            for (int depth = 0; depth < maxDepth; ++depth) {
                if (depth > 3 && beta.MaxComponentValue() < 0.25f) {
                    float q = stdmax(0.f, 1 - beta.MaxComponentValue());
                    if (r() < q) break;
                    beta /= 1 - q;
                }
                if (!albedo.IsNonZero()) {
                    z = (z == thickness) ? 0 : thickness;
                    beta *= Tr(thickness, w);
                } else {
                    float sigma_t = 1;
                    float dz = SampleExponential(r(), sigma_t / fabsf(w.z));
                    float zp = w.z > 0 ? (z + dz) : (z - dz);
                    if (z == zp) continue;
                    if (0 < zp && zp < thickness) {
                        float wt = 1;
                        if (!IsSpecular(IFlags(exitI))) wt = PowerHeuristic(1, wis.pdf, 1, HenyeyGreenstein(Dot(-w, -wis.wi), g));
                        fsum += beta * albedo * HenyeyGreenstein(Dot(-w, -wis.wi), g) * wt * Tr(zp - exitZ, wis.wi) * wis.f / wis.pdf;
                        float u0 = r(); float u1 = r();   // Point2f u{r(), r()}: braced list, left to right
                        float ppdf;
                        Vec3 pwi = SampleHenyeyGreenstein(-w, g, Vec2(u0, u1), &ppdf);
                        if (ppdf == 0 || pwi.z == 0) continue;
                        beta *= albedo * ppdf / ppdf;
                        w = pwi;
                        z = zp;
                        if (((z < exitZ && w.z > 0) || (z > exitZ && w.z < 0)) && !IsSpecular(IFlags(exitI))) {
                            Spec fExit = If(exitI, -w, wi, mode);
                            if (fExit.IsNonZero()) {
                                float exitPDF = IPDF(exitI, -w, wi, mode, REFLTRANS_TRANSMISSION);
                                float wt2 = PowerHeuristic(1, ppdf, 1, exitPDF);
                                fsum += beta * Tr(zp - exitZ, pwi) * fExit * wt2;
                            }
                        }
                        continue;
                    }
                    z = Clampf(zp, 0, thickness);
                }
                if (z == exitZ) {
                    float uc3 = r();
                    BPT_R2(ue, uf)
                    BSDFSample bs = ISample(exitI, -w, uc3, Vec2(ue, uf), mode, REFLTRANS_REFLECTION);
                    if (Bad(bs)) break;
                    beta *= bs.f * AbsCosTheta(bs.wi) / bs.pdf;
                    w = bs.wi;
                } else {
                    if (!IsSpecular(IFlags(nonExitI))) {
                        float wt = 1;
                        if (!IsSpecular(IFlags(exitI))) wt = PowerHeuristic(1, wis.pdf, 1, IPDF(nonExitI, -w, -wis.wi, mode));
                        fsum += beta * If(nonExitI, -w, -wis.wi, mode) * AbsCosTheta(wis.wi) * wt * Tr(thickness, wis.wi) * wis.f / wis.pdf;
                    }
                    float uc4 = r();
                    BPT_R2(ug, uh)   // Point2f u(r(), r())
                    BSDFSample bs = ISample(nonExitI, -w, uc4, Vec2(ug, uh), mode, REFLTRANS_REFLECTION);
                    if (Bad(bs)) break;
                    beta *= bs.f * AbsCosTheta(bs.wi) / bs.pdf;
                    w = bs.wi;
                    if (!IsSpecular(IFlags(exitI))) {
                        Spec fExit = If(exitI, -w, wi, mode);
                        if (fExit.IsNonZero()) {
                            float wt = 1;
                            if (!IsSpecular(IFlags(nonExitI))) {
                                float exitPDF = IPDF(exitI, -w, wi, mode, REFLTRANS_TRANSMISSION);
                                wt = PowerHeuristic(1, bs.pdf, 1, exitPDF);
                            }
                            fsum += beta * Tr(thickness, bs.wi) * fExit * wt;
                        }
                    }
                }
            }
        }
        return fsum / nSamples;
    }

    BPT_BXDF_FN BSDFSample Sample_f(Vec3 wo, float uc, Vec2 u, int mode, int sampleFlags = REFLTRANS_ALL) const {
        bool flipWi = false;
        if (wo.z < 0) { wo = -wo; flipWi = true; }
        BSDFSample bs = top.Sample_f(wo, uc, u, mode);
        if (Bad(bs)) return {};
        if (IsReflective(bs.flags)) {
            if (flipWi) bs.wi = -bs.wi;
            bs.pdfIsProportional = true;
            return bs;
        }
        Vec3 w = bs.wi;
        bool specularPath = IsSpecular(bs.flags);
        uint64_t wu[2] = {(uint64_t)f2b(uc) | ((uint64_t)f2b(u.x) << 32), (uint64_t)f2b(u.y)};
        RNG rng(Hash_i32_f32x3(seed, wo), MurmurHash64A(wu, 12, 0));   // Hash(uc, u)
        auto r = [&rng]() { return stdmin(rng.UniformFloat(), kOneMinusEpsilon); };
        Spec f = bs.f * AbsCosTheta(bs.wi);
        float pdf = bs.pdf;
        float z = thickness;
        for (int depth = 0; depth < maxDepth; ++depth) {
            float rrBeta = f.MaxComponentValue() / pdf;
            if (depth > 3 && rrBeta < 0.25f) {
                float q = stdmax(0.f, 1 - rrBeta);
                if (r() < q) return {};
                pdf *= 1 - q;
            }
            if (w.z == 0) return {};
            if (albedo.IsNonZero()) {
                float sigma_t = 1;
                float dz = SampleExponential(r(), sigma_t / AbsCosTheta(w));
                float zp = w.z > 0 ? (z + dz) : (z - dz);
                if (zp == z) return {};
                if (0 < zp && zp < thickness) {
                    BPT_R2(up0, up1)   // phase.Sample_p(-w, Point2f(r(), r()))
                    float ppdf;
                    Vec3 pwi = SampleHenyeyGreenstein(-w, g, Vec2(up0, up1), &ppdf);
                    if (ppdf == 0 || pwi.z == 0) return {};
                    f *= albedo * ppdf;
                    pdf *= ppdf;
                    specularPath = false;
                    w = pwi;
                    z = zp;
                    continue;
                }
                z = Clampf(zp, 0, thickness);
            } else {
                z = (z == thickness) ? 0 : thickness;
                f *= Tr(thickness, w);
            }
            int iface = (z == 0) ? 1 : 0;
            float uc2 = r();
            BPT_R2(ux, uy)   // Point2f u(r(), r())
            BSDFSample b2 = ISample(iface, -w, uc2, Vec2(ux, uy), mode);
            if (Bad(b2)) return {};
            f *= b2.f;
            pdf *= b2.pdf;
            specularPath &= IsSpecular(b2.flags);
            w = b2.wi;
            if (IsTransmissive(b2.flags)) {
                int flags = SameHemisphere(wo, w) ? BXDF_REFLECTION : BXDF_TRANSMISSION;
                flags |= specularPath ? BXDF_SPECULAR : BXDF_GLOSSY;
                if (flipWi) w = -w;
                return BSDFSample(f, w, pdf, flags, 1.f, true);
            }
            f *= AbsCosTheta(b2.wi);
        }
        return {};
    }

    BPT_BXDF_FN float PDF(Vec3 wo, Vec3 wi, int mode, int sampleFlags = REFLTRANS_ALL) const {
        if (wo.z < 0) { wo = -wo; wi = -wi; }
        RNG rng(Hash_i32_f32x3(seed, wi), Hash_f32x3(wo.x, wo.y, wo.z));
        auto r = [&rng]() { return stdmin(rng.UniformFloat(), kOneMinusEpsilon); };
        float pdfSum = 0;
        if (SameHemisphere(wo, wi)) pdfSum += nSamples * top.PDF(wo, wi, mode, REFLTRANS_REFLECTION);
        for (int s = 0; s < nSamples; ++s) {
            if (SameHemisphere(wo, wi)) {
                const int rI = 1, tI = 0;
                // tInterface.Sample_f(wo, r(), {r(), r()}, mode, trans): arguments right to left -> the braced pair first
                float a0 = r(); float a1 = r(); float auc = r();
                BSDFSample wos = ISample(tI, wo, auc, Vec2(a0, a1), mode, REFLTRANS_TRANSMISSION);
                float b0 = r(); float b1 = r(); float buc = r();
                BSDFSample wis = ISample(tI, wi, buc, Vec2(b0, b1), !mode, REFLTRANS_TRANSMISSION);
                if (wos.valid && wos.f.IsNonZero() && wos.pdf > 0 && wis.valid && wis.f.IsNonZero() && wis.pdf > 0) {
                    if (!IsNonSpecular(IFlags(tI))) pdfSum += IPDF(rI, -wos.wi, -wis.wi, mode);
                    else {
                        float c0 = r(); float c1 = r(); float cuc = r();
                        BSDFSample rs = ISample(rI, -wos.wi, cuc, Vec2(c0, c1), mode);
                        if (rs.valid && rs.f.IsNonZero() && rs.pdf > 0) {
                            if (!IsNonSpecular(IFlags(rI))) pdfSum += IPDF(tI, -rs.wi, wi, mode);
                            else {
                                float rPDF = IPDF(rI, -wos.wi, -wis.wi, mode);
                                float wt = PowerHeuristic(1, wis.pdf, 1, rPDF);
                                pdfSum += wt * rPDF;
                                float tPDF = IPDF(tI, -rs.wi, wi, mode);
                                wt = PowerHeuristic(1, rs.pdf, 1, tPDF);
                                pdfSum += wt * tPDF;
                            }
                        }
                    }
                }
            } else {
                const int toI = 0, tiI = 1;
                float uc = r();
                BPT_R2(d0, d1)
                BSDFSample wos = ISample(toI, wo, uc, Vec2(d0, d1), mode);
                if (Bad(wos) || IsReflective(wos.flags)) continue;
                uc = r();
                BPT_R2(e0, e1)
                BSDFSample wis = ISample(tiI, wi, uc, Vec2(e0, e1), !mode);
                if (Bad(wis) || IsReflective(wis.flags)) continue;
                if (IsSpecular(IFlags(toI))) pdfSum += IPDF(tiI, -wos.wi, wi, mode);
                else if (IsSpecular(IFlags(tiI))) pdfSum += IPDF(toI, wo, -wis.wi, mode);
                else pdfSum += (IPDF(toI, wo, -wis.wi, mode) + IPDF(tiI, -wos.wi, wi, mode)) / 2;
            }
        }
        return Lerp(0.9f, 1 / (4 * kPi), pdfSum / nSamples);
    }
};
typedef LayeredBxDF<DielectricBxDF, ConductorBxDF> CoatedConductorBxDF;
typedef LayeredBxDF<DielectricBxDF, DiffuseBxDF> CoatedDiffuseBxDF;   // bxdfs.h:905-912 (both are two-sided)

// DiffuseTransmissionBxDF (bxdfs.h:85-165)
struct DiffuseTransmissionBxDF {
    Spec R, T;
    BPT_HD int Flags() const { return (R.IsNonZero() ? BXDF_DIFFUSE_REFLECTION : BXDF_UNSET) | (T.IsNonZero() ? BXDF_DIFFUSE_TRANSMISSION : BXDF_UNSET); }
    BPT_HD Spec f(Vec3 wo, Vec3 wi, int mode) const { return SameHemisphere(wo, wi) ? (R * kInvPi) : (T * kInvPi); }
    BPT_HD BSDFSample Sample_f(Vec3 wo, float uc, Vec2 u, int mode, int sampleFlags = REFLTRANS_ALL) const {
        float pr = R.MaxComponentValue(), pt = T.MaxComponentValue();
        if (!(sampleFlags & REFLTRANS_REFLECTION)) pr = 0;
        if (!(sampleFlags & REFLTRANS_TRANSMISSION)) pt = 0;
        if (pr == 0 && pt == 0) return BSDFSample();
        if (uc < pr / (pr + pt)) {
            Vec3 wi = SampleCosineHemisphere(u);
            if (wo.z < 0) wi.z *= -1;
            float pdf = CosineHemispherePDF(AbsCosTheta(wi)) * pr / (pr + pt);
            return BSDFSample(f(wo, wi, mode), wi, pdf, BXDF_DIFFUSE_REFLECTION);
        } else {
            Vec3 wi = SampleCosineHemisphere(u);
            if (wo.z > 0) wi.z *= -1;
            float pdf = CosineHemispherePDF(AbsCosTheta(wi)) * pt / (pr + pt);
            return BSDFSample(f(wo, wi, mode), wi, pdf, BXDF_DIFFUSE_TRANSMISSION);
        }
    }
    BPT_HD float PDF(Vec3 wo, Vec3 wi, int mode, int sampleFlags = REFLTRANS_ALL) const {
        float pr = R.MaxComponentValue(), pt = T.MaxComponentValue();
        if (!(sampleFlags & REFLTRANS_REFLECTION)) pr = 0;
        if (!(sampleFlags & REFLTRANS_TRANSMISSION)) pt = 0;
        if (pr == 0 && pt == 0) return 0;
        if (SameHemisphere(wo, wi)) return pr / (pr + pt) * CosineHemispherePDF(AbsCosTheta(wi));
        else return pt / (pr + pt) * CosineHemispherePDF(AbsCosTheta(wi));
    }
    BPT_HD void Regularize() {}
};

// ThinDielectricBxDF (bxdfs.h:209-277)
struct ThinDielectricBxDF {
    float eta;
    BPT_HD int Flags() const { return BXDF_REFLECTION | BXDF_TRANSMISSION | BXDF_SPECULAR; }
    BPT_HD Spec f(Vec3, Vec3, int) const { return Spec(0.f); }
    BPT_HD BSDFSample Sample_f(Vec3 wo, float uc, Vec2 u, int mode, int sampleFlags = REFLTRANS_ALL) const {
        float R = FrDielectric(AbsCosTheta(wo), eta), T = 1 - R;
        if (R < 1) { R += Sqr(T) * R / (1 - Sqr(R)); T = 1 - R; }
        float pr = R, pt = T;
        if (!(sampleFlags & REFLTRANS_REFLECTION)) pr = 0;
        if (!(sampleFlags & REFLTRANS_TRANSMISSION)) pt = 0;
        if (pr == 0 && pt == 0) return BSDFSample();
        if (uc < pr / (pr + pt)) {
            Vec3 wi(-wo.x, -wo.y, wo.z);
            Spec fr(R / AbsCosTheta(wi));
            return BSDFSample(fr, wi, pr / (pr + pt), BXDF_SPECULAR_REFLECTION);
        } else {
            Vec3 wi = -wo;
            Spec ft(T / AbsCosTheta(wi));
            return BSDFSample(ft, wi, pt / (pr + pt), BXDF_SPECULAR_TRANSMISSION);
        }
    }
    BPT_HD float PDF(Vec3, Vec3, int, int = REFLTRANS_ALL) const { return 0; }
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
