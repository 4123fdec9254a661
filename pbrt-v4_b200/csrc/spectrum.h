// spectrum.h -- 4-wavelength SampledSpectrum / SampledWavelengths and the spectrum kinds the path evaluates.
//   SampledSpectrum, SampledWavelengths      util/spectrum.h:36-349  (NSpectrumSamples = 4)
//   DenselySampledSpectrum::Sample           util/spectrum.h:  lround(lambda) - 360 table lookup
//   RGBSigmoidPolynomial                     util/color.h
//   RGBAlbedo/Unbounded/IlluminantSpectrum   util/spectrum.h, spectrum.cpp:230-247
#pragma once
#include "pmath.h"
#include "sampling.h"

namespace bpt {

constexpr int kLambdaMin = 360, kLambdaMax = 830, kNLambda = kLambdaMax - kLambdaMin + 1;

struct Spec {
    float v[4];
    BPT_HD Spec() { v[0] = v[1] = v[2] = v[3] = 0; }
    BPT_HD explicit Spec(float c) { v[0] = v[1] = v[2] = v[3] = c; }
    BPT_HD Spec(float a, float b, float c, float d) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
    BPT_HD float operator[](int i) const { return v[i]; }
    BPT_HD float &operator[](int i) { return v[i]; }
    BPT_HD Spec operator+(Spec s) const { return {v[0] + s.v[0], v[1] + s.v[1], v[2] + s.v[2], v[3] + s.v[3]}; }
    BPT_HD Spec operator-(Spec s) const { return {v[0] - s.v[0], v[1] - s.v[1], v[2] - s.v[2], v[3] - s.v[3]}; }
    BPT_HD Spec operator*(Spec s) const { return {v[0] * s.v[0], v[1] * s.v[1], v[2] * s.v[2], v[3] * s.v[3]}; }
    BPT_HD Spec operator/(Spec s) const { return {v[0] / s.v[0], v[1] / s.v[1], v[2] / s.v[2], v[3] / s.v[3]}; }
    BPT_HD Spec operator*(float a) const { return {v[0] * a, v[1] * a, v[2] * a, v[3] * a}; }
    BPT_HD Spec operator/(float a) const { return {v[0] / a, v[1] / a, v[2] / a, v[3] / a}; }
    BPT_HD Spec &operator+=(Spec s) { *this = *this + s; return *this; }
    BPT_HD Spec &operator*=(Spec s) { *this = *this * s; return *this; }
    BPT_HD Spec &operator*=(float a) { *this = *this * a; return *this; }
    BPT_HD Spec &operator/=(float a) { *this = *this / a; return *this; }
    BPT_HD Spec operator-() const { return {-v[0], -v[1], -v[2], -v[3]}; }
    BPT_HD bool IsNonZero() const { return v[0] != 0 || v[1] != 0 || v[2] != 0 || v[3] != 0; }   // operator bool
    BPT_HD float Average() const { float s = v[0]; s += v[1]; s += v[2]; s += v[3]; return s / 4; }
    BPT_HD float MaxComponentValue() const { float m = v[0]; m = stdmax(m, v[1]); m = stdmax(m, v[2]); m = stdmax(m, v[3]); return m; }
    BPT_HD float MinComponentValue() const { float m = v[0]; m = stdmin(m, v[1]); m = stdmin(m, v[2]); m = stdmin(m, v[3]); return m; }
    BPT_HD bool HasNaNs() const { return IsNaN_(v[0]) || IsNaN_(v[1]) || IsNaN_(v[2]) || IsNaN_(v[3]); }
};
BPT_HD Spec operator*(float a, Spec s) { return s * a; }
BPT_HD Spec SafeDiv(Spec a, Spec b) {
    Spec r;
    for (int i = 0; i < 4; ++i) r[i] = (b[i] != 0) ? a[i] / b[i] : 0.f;
    return r;
}
BPT_HD Spec ClampSpec(Spec s, float lo, float hi) {
    Spec r;
    for (int i = 0; i < 4; ++i) r[i] = Clampf(s[i], lo, hi);
    return r;
}
BPT_HD Spec ClampZero(Spec s) {
    Spec r;
    for (int i = 0; i < 4; ++i) r[i] = stdmax(0.f, s[i]);
    return r;
}
BPT_HD Spec SqrtSpec(Spec s) { return {sqrtf(s[0]), sqrtf(s[1]), sqrtf(s[2]), sqrtf(s[3])}; }
BPT_HD Spec SafeSqrtSpec(Spec s) { return {SafeSqrt(s[0]), SafeSqrt(s[1]), SafeSqrt(s[2]), SafeSqrt(s[3])}; }
BPT_HD Spec ExpSpec(Spec s) { return {expf(s[0]), expf(s[1]), expf(s[2]), expf(s[3])}; }
BPT_HD Spec FastExpSpec(Spec s) { return {FastExp(s[0]), FastExp(s[1]), FastExp(s[2]), FastExp(s[3])}; }

struct Wavelengths {
    float lambda[4], pdf[4];
    BPT_HD static Wavelengths SampleVisible(float u) {
        Wavelengths swl;
        for (int i = 0; i < 4; ++i) {
            float up = u + float(i) / 4;
            if (up > 1) up -= 1;
            swl.lambda[i] = SampleVisibleWavelengths(up);
            swl.pdf[i] = VisibleWavelengthsPDF(swl.lambda[i]);
        }
        return swl;
    }
    BPT_HD static Wavelengths SampleUniform(float u) {
        Wavelengths swl;
        float lmin = (float)kLambdaMin, lmax = (float)kLambdaMax;
        swl.lambda[0] = Lerp(u, lmin, lmax);
        float delta = (lmax - lmin) / 4;
        for (int i = 1; i < 4; ++i) {
            swl.lambda[i] = swl.lambda[i - 1] + delta;
            if (swl.lambda[i] > lmax) swl.lambda[i] = lmin + (swl.lambda[i] - lmax);
        }
        for (int i = 0; i < 4; ++i) swl.pdf[i] = 1 / (lmax - lmin);
        return swl;
    }
    BPT_HD Spec PDF() const { return {pdf[0], pdf[1], pdf[2], pdf[3]}; }
    BPT_HD bool SecondaryTerminated() const { return pdf[1] == 0 && pdf[2] == 0 && pdf[3] == 0; }
    BPT_HD void TerminateSecondary() {
        if (SecondaryTerminated()) return;
        pdf[1] = pdf[2] = pdf[3] = 0;
        pdf[0] /= 4;
    }
};

// 471-entry table sampled at integer nm (DenselySampledSpectrum)
BPT_HD float DenseLookup(const float *table, float lambda) {
    int offset = (int)lroundf(lambda) - kLambdaMin;
    if (offset < 0 || offset >= kNLambda) return 0;
    return table[offset];
}
BPT_HD Spec DenseSample(const float *table, const Wavelengths &wl) {
    return {DenseLookup(table, wl.lambda[0]), DenseLookup(table, wl.lambda[1]), DenseLookup(table, wl.lambda[2]),
            DenseLookup(table, wl.lambda[3])};
}

// RGBSigmoidPolynomial::operator()
BPT_HD float SigmoidPoly(float c0, float c1, float c2, float lambda) {
    float x = EvalPoly(lambda, c2, c1, c0);
    if (IsInf_(x)) return x > 0 ? 1.f : 0.f;
    return .5f + x / (2 * sqrtf(1 + Sqr(x)));
}

// Spectrum record: the handful of Spectrum kinds that reach the kernels through constant textures / lights.
enum SpectrumKind { SPEC_CONSTANT = 0, SPEC_RGB_ALBEDO = 1, SPEC_RGB_UNBOUNDED = 2, SPEC_DENSE = 3, SPEC_PIECEWISE = 4 };
struct SpectrumRec {
    int kind;
    float c0, c1, c2;    // sigmoid polynomial, or c0 = constant value
    float scale;         // RGBUnbounded scale
    int table;           // SPEC_DENSE: offset (in floats) into the dense table pool (471 floats);
                         // SPEC_PIECEWISE: offset of [lambdas..., values...] in the pool
    int n;               // SPEC_PIECEWISE: number of samples
    int pad;
};
BPT_HD float PiecewiseLinearEval(const float *lambdas, const float *values, int n, float lambda) {
    // PiecewiseLinearSpectrum::operator() (util/spectrum.cpp)
    if (n == 0 || lambda < lambdas[0] || lambda > lambdas[n - 1]) return 0;
    int o = FindIntervalLE(lambdas, n, lambda);
    float t = (lambda - lambdas[o]) / (lambdas[o + 1] - lambdas[o]);
    return Lerp(t, values[o], values[o + 1]);
}
BPT_HD Spec EvalSpectrum(const SpectrumRec &s, const float *pool, const Wavelengths &wl) {
    Spec r;
    switch (s.kind) {
    case SPEC_CONSTANT: return Spec(s.c0);
    case SPEC_RGB_ALBEDO:
        for (int i = 0; i < 4; ++i) r[i] = SigmoidPoly(s.c0, s.c1, s.c2, wl.lambda[i]);
        return r;
    case SPEC_RGB_UNBOUNDED:
        for (int i = 0; i < 4; ++i) r[i] = s.scale * SigmoidPoly(s.c0, s.c1, s.c2, wl.lambda[i]);
        return r;
    case SPEC_DENSE: return DenseSample(pool + s.table, wl);
    case SPEC_PIECEWISE:
        for (int i = 0; i < 4; ++i) r[i] = PiecewiseLinearEval(pool + s.table, pool + s.table + s.n, s.n, wl.lambda[i]);
        return r;
    }
    return r;
}

}  // namespace bpt
