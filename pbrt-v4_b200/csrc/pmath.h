// pmath.h -- float-exact scalar / vector arithmetic shared by the host scene setup and the sm_100a kernels.
//
// Everything here restates the *CPU* branches of the reference's util layer so that a path computed on the
// GPU goes through the same IEEE-754 operations, in the same order, as the reference CPU VolPathIntegrator:
//   FMA / DifferenceOfProducts / SumOfProducts / InnerProduct  util/math.h:283,553-612
//   gamma(n), NextFloatUp/Down, Add/Sub/Mul/DivRoundUp/Down    util/float.h:36,195,  math.h:744-800 (CPU branch)
//   Interval                                                   util/math.h:816-960
//   Vector/Point/Normal ops (note: Dot(Vector,Vector) is plain, Dot(Normal,*) is FMA+SumOfProducts)
//                                                              util/vecmath.h:932-1120
// The translation unit is compiled with -fmad=false (device) and -ffp-contract=off (host); every fused
// multiply-add below is explicit.
#pragma once
#include "libm_exact.h"
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cfloat>

#if defined(__CUDACC__)
#define BPT_HD __host__ __device__ __forceinline__
#define BPT_D __device__ __forceinline__
#else
#define BPT_HD inline
#define BPT_D inline
#endif

namespace bpt {

constexpr float kPi = 3.14159265358979323846f;
constexpr float kInvPi = 0.31830988618379067154f;
constexpr float kInv2Pi = 0.15915494309189533577f;
constexpr float kInv4Pi = 0.07957747154594766788f;
constexpr float kPiOver2 = 1.57079632679489661923f;
constexpr float kPiOver4 = 0.78539816339744830961f;
constexpr float kSqrt2 = 1.41421356237309504880f;
constexpr float kShadowEpsilon = 0.0001f;
constexpr float kOneMinusEpsilon = 0x1.fffffep-1f;
constexpr float kMachineEpsilon = FLT_EPSILON * 0.5f;
#if defined(__CUDA_ARCH__)
#define BPT_INFINITY __int_as_float(0x7f800000)
#else
#define BPT_INFINITY (std::numeric_limits<float>::infinity())
#endif

BPT_HD constexpr float gamma_n(int n) { return (n * kMachineEpsilon) / (1 - n * kMachineEpsilon); }

BPT_HD uint32_t f2b(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; std::memcpy(&u, &f, 4); return u;
#endif
}
BPT_HD float b2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; std::memcpy(&f, &u, 4); return f;
#endif
}
BPT_HD uint64_t d2b(double f) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(f);
#else
    uint64_t u; std::memcpy(&u, &f, 8); return u;
#endif
}

BPT_HD float FMA(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
BPT_HD double FMA(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
BPT_HD float Sqr(float x) { return x * x; }
BPT_HD float fabs_(float x) { return fabsf(x); }
BPT_HD float fmin_(float a, float b) { return a < b ? a : b; }   // std::min(a,b): (b < a) ? b : a
BPT_HD float fmax_(float a, float b) { return a < b ? b : a; }   // std::max(a,b): (a < b) ? b : a
// std::min / std::max semantic restated exactly (matters only for NaN / signed zero)
BPT_HD float stdmin(float a, float b) { return (b < a) ? b : a; }
BPT_HD float stdmax(float a, float b) { return (a < b) ? b : a; }
BPT_HD int imin(int a, int b) { return b < a ? b : a; }
BPT_HD int imax(int a, int b) { return a < b ? b : a; }
BPT_HD float Clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
BPT_HD int Clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
BPT_HD float Lerp(float x, float a, float b) { return (1 - x) * a + x * b; }
BPT_HD float SafeSqrt(float x) { return sqrtf(stdmax(0.f, x)); }
BPT_HD float Radians(float deg) { return (kPi / 180) * deg; }
BPT_HD bool IsNaN_(float v) { return v != v; }
BPT_HD bool IsInf_(float v) { return fabsf(v) == BPT_INFINITY; }

// Transcendentals.  The reference runs on glibc's float libm.  libdevice's float versions differ from glibc in
// 15-33 % of arguments (measured on the B200 box, scripts/probe_math.py); evaluating in double and rounding once
// differs in 0.06-1.3 % (sin, cos, exp, log, pow, tan) to 7-20 % (asin, acos, atan2, atanh, cosh) -- always by 1 ulp.
// Host code uses libm itself.
#if defined(__CUDA_ARCH__)
#define BPT_DEF_T1(name, dfn, hfn) BPT_HD float name(float x) { return (float)dfn((double)x); }
#else
#define BPT_DEF_T1(name, dfn, hfn) BPT_HD float name(float x) { return hfn(x); }
#endif
// BPT_EXACT_LIBM (default on): glibc's own float algorithms on the device (libm_exact.h: sin, cos, log, exp, asin, acos, atan2,
// atanh, cosh restated; exhaustively equal to the host libm on the CPU, tests/tools/libm_exact_check.c; 0 mismatches on the B200,
// profiles/r01p; the GPU parity suite is green with it).  Every transcendental on the render path then reproduces the reference's
// bits; BPT_EXACT_LIBM=0 restores "evaluate in FP64, round once" (0.06 - 20 % of arguments off by 1 ulp).
#ifndef BPT_EXACT_LIBM
#define BPT_EXACT_LIBM 1
#endif
// BPT_OUTLINE_LIBM: the restated routines as real functions (one copy per kernel module instead of one per call site): the
// shading kernels are 20 k - 90 k instructions when everything is inlined, far beyond the instruction caches
#ifndef BPT_OUTLINE_LIBM
#define BPT_OUTLINE_LIBM 1
#endif
#if BPT_OUTLINE_LIBM && defined(__CUDA_ARCH__)
#define BPT_LIBM_FN __device__ __noinline__
#else
#define BPT_LIBM_FN BPT_HD
#endif
#if BPT_EXACT_LIBM && defined(__CUDA_ARCH__)
BPT_LIBM_FN float SinF(float x) { return bpt_exact_sinf(x); }
BPT_LIBM_FN float CosF(float x) { return bpt_exact_cosf(x); }
BPT_LIBM_FN float ASinF(float x) { return bpt_exact_asinf(x); }
BPT_LIBM_FN float ACosF(float x) { return bpt_exact_acosf(x); }
BPT_LIBM_FN float ATanhF(float x) { return bpt_exact_atanhf(x); }
BPT_LIBM_FN float CoshF(float x) { return bpt_exact_coshf(x); }
BPT_LIBM_FN float ExpF(float x) { return bpt_exact_expf(x); }
BPT_LIBM_FN float LogF(float x) { return bpt_exact_logf(x); }
BPT_LIBM_FN float ATan2F(float y, float x) { return bpt_exact_atan2f(y, x); }
#else
BPT_DEF_T1(SinF, sin, sinf)
BPT_DEF_T1(CosF, cos, cosf)
BPT_DEF_T1(ASinF, asin, asinf)
BPT_DEF_T1(ACosF, acos, acosf)
BPT_DEF_T1(ATanhF, atanh, atanhf)
BPT_DEF_T1(CoshF, cosh, coshf)
BPT_DEF_T1(ExpF, exp, expf)
BPT_DEF_T1(LogF, log, logf)
BPT_HD float ATan2F(float y, float x) {
#if defined(__CUDA_ARCH__)
    return (float)atan2((double)y, (double)x);
#else
    return atan2f(y, x);
#endif
}
#endif
BPT_DEF_T1(TanF, tan, tanf)
BPT_HD float PowF(float x, float y) {
#if defined(__CUDA_ARCH__)
    return (float)pow((double)x, (double)y);
#else
    return powf(x, y);
#endif
}

BPT_HD float DifferenceOfProducts(float a, float b, float c, float d) {
    float cd = c * d;
    float dop = FMA(a, b, -cd);
    float err = FMA(-c, d, cd);
    return dop + err;
}
BPT_HD float SumOfProducts(float a, float b, float c, float d) {
    float cd = c * d;
    float sop = FMA(a, b, cd);
    float err = FMA(c, d, -cd);
    return sop + err;
}

// CompensatedFloat / InnerProduct (util/math.h:553-612)
struct CompF { float v, err; };
BPT_HD CompF TwoProd(float a, float b) { float ab = a * b; return {ab, FMA(a, b, -ab)}; }
BPT_HD CompF TwoSum(float a, float b) {
    float s = a + b, delta = s - a;
    return {s, (a - (s - delta)) + (b - delta)};
}
BPT_HD CompF InnerProductC(float a, float b) { return TwoProd(a, b); }
template <typename... T>
BPT_HD CompF InnerProductC(float a, float b, T... terms) {
    CompF ab = TwoProd(a, b);
    CompF tp = InnerProductC(terms...);
    CompF sum = TwoSum(ab.v, tp.v);
    return {sum.v, ab.err + (tp.err + sum.err)};
}
template <typename... T>
BPT_HD float InnerProduct(T... terms) { CompF ip = InnerProductC(terms...); return ip.v + ip.err; }

// polynomial via Horner with FMA (util/math.h:328-337)
BPT_HD float EvalPoly(float t, float c0) { return c0; }
template <typename... A>
BPT_HD float EvalPoly(float t, float c0, A... rest) { return FMA(t, EvalPoly(t, rest...), c0); }

// FastExp, CPU branch (util/math.h:450-475)
BPT_HD float FastExp(float x) {
    float xp = x * 1.442695041f;
    float fxp = floorf(xp), f = xp - fxp;
    int i = (int)fxp;
    float twoToF = EvalPoly(f, 1.f, 0.695556856f, 0.226173572f, 0.0781455737f);
    int exponent = (int)(f2b(twoToF) >> 23) - 127 + i;
    if (exponent < -126) return 0;
    if (exponent > 127) return BPT_INFINITY;
    uint32_t bits = f2b(twoToF);
    bits &= 0b10000000011111111111111111111111u;
    bits |= (uint32_t)(exponent + 127) << 23;
    return b2f(bits);
}
BPT_HD float Gaussian(float x, float mu, float sigma) {
    return 1 / sqrtf(2 * kPi * sigma * sigma) * FastExp(-Sqr(x - mu) / (2 * sigma * sigma));
}

BPT_HD float NextFloatUp(float v) {
    if (IsInf_(v) && v > 0.f) return v;
    if (v == -0.f) v = 0.f;
    uint32_t ui = f2b(v);
    if (v >= 0) ++ui; else --ui;
    return b2f(ui);
}
BPT_HD float NextFloatDown(float v) {
    if (IsInf_(v) && v < 0.f) return v;
    if (v == 0.f) v = -0.f;
    uint32_t ui = f2b(v);
    if (v > 0) --ui; else ++ui;
    return b2f(ui);
}
// CPU branches of the directed-rounding helpers (util/math.h: AddRoundUp etc. use NextFloatUp(a+b))
BPT_HD float AddRoundUp(float a, float b) { return NextFloatUp(a + b); }
BPT_HD float AddRoundDown(float a, float b) { return NextFloatDown(a + b); }
BPT_HD float SubRoundUp(float a, float b) { return AddRoundUp(a, -b); }
BPT_HD float SubRoundDown(float a, float b) { return AddRoundDown(a, -b); }
BPT_HD float MulRoundUp(float a, float b) { return NextFloatUp(a * b); }
BPT_HD float MulRoundDown(float a, float b) { return NextFloatDown(a * b); }
BPT_HD float DivRoundUp(float a, float b) { return NextFloatUp(a / b); }
BPT_HD float DivRoundDown(float a, float b) { return NextFloatDown(a / b); }

// ---------------------------------------------------------------------------------------------
// Interval (util/math.h:816-960) -- only the operations the path needs
struct Interval {
    float low, high;
    BPT_HD Interval() : low(0), high(0) {}
    BPT_HD explicit Interval(float v) : low(v), high(v) {}
    BPT_HD Interval(float l, float h) : low(stdmin(l, h)), high(stdmax(l, h)) {}
    BPT_HD static Interval FromValueAndError(float v, float err) {
        Interval i;
        if (err == 0) i.low = i.high = v;
        else { i.low = SubRoundDown(v, err); i.high = AddRoundUp(v, err); }
        return i;
    }
    BPT_HD float Midpoint() const { return (low + high) / 2; }
    BPT_HD float Width() const { return high - low; }
    BPT_HD Interval operator+(Interval i) const { return {AddRoundDown(low, i.low), AddRoundUp(high, i.high)}; }
    BPT_HD Interval operator-(Interval i) const { return {SubRoundDown(low, i.high), SubRoundUp(high, i.low)}; }
    BPT_HD Interval operator+(float f) const { return *this + Interval(f); }
    BPT_HD Interval operator-(float f) const { return *this - Interval(f); }
};

// ---------------------------------------------------------------------------------------------
// Tuples.  Distinct Vec3 / Pt3 / Nrm3 types so that overload resolution picks the same arithmetic the
// reference picks (vecmath.h: Vector.Vector dot is a plain sum, Normal dots use FMA + SumOfProducts).
struct Vec2 { float x, y; BPT_HD Vec2() : x(0), y(0) {} BPT_HD Vec2(float a, float b) : x(a), y(b) {} };
typedef Vec2 Pt2;

struct Vec3 {
    float x, y, z;
    BPT_HD Vec3() : x(0), y(0), z(0) {}
    BPT_HD Vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    BPT_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    BPT_HD Vec3 operator-() const { return {-x, -y, -z}; }
    BPT_HD Vec3 operator+(Vec3 v) const { return {x + v.x, y + v.y, z + v.z}; }
    BPT_HD Vec3 operator-(Vec3 v) const { return {x - v.x, y - v.y, z - v.z}; }
    BPT_HD Vec3 operator*(float s) const { return {x * s, y * s, z * s}; }
    BPT_HD Vec3 operator/(float s) const { return {x / s, y / s, z / s}; }
    BPT_HD bool operator==(Vec3 v) const { return x == v.x && y == v.y && z == v.z; }
    BPT_HD bool operator!=(Vec3 v) const { return x != v.x || y != v.y || z != v.z; }
};
struct Nrm3 {
    float x, y, z;
    BPT_HD Nrm3() : x(0), y(0), z(0) {}
    BPT_HD Nrm3(float a, float b, float c) : x(a), y(b), z(c) {}
    BPT_HD explicit Nrm3(Vec3 v) : x(v.x), y(v.y), z(v.z) {}
    BPT_HD Nrm3 operator-() const { return {-x, -y, -z}; }
    BPT_HD Nrm3 operator+(Nrm3 v) const { return {x + v.x, y + v.y, z + v.z}; }
    BPT_HD Nrm3 operator-(Nrm3 v) const { return {x - v.x, y - v.y, z - v.z}; }
    BPT_HD Nrm3 operator*(float s) const { return {x * s, y * s, z * s}; }
    BPT_HD Nrm3 operator/(float s) const { return {x / s, y / s, z / s}; }
    BPT_HD bool operator==(Nrm3 v) const { return x == v.x && y == v.y && z == v.z; }
    BPT_HD bool operator!=(Nrm3 v) const { return x != v.x || y != v.y || z != v.z; }
};
struct Pt3 {
    float x, y, z;
    BPT_HD Pt3() : x(0), y(0), z(0) {}
    BPT_HD Pt3(float a, float b, float c) : x(a), y(b), z(c) {}
    BPT_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    BPT_HD Pt3 operator+(Vec3 v) const { return {x + v.x, y + v.y, z + v.z}; }
    BPT_HD Pt3 operator-(Vec3 v) const { return {x - v.x, y - v.y, z - v.z}; }
    BPT_HD Vec3 operator-(Pt3 p) const { return {x - p.x, y - p.y, z - p.z}; }
    BPT_HD Pt3 operator+(Pt3 p) const { return {x + p.x, y + p.y, z + p.z}; }   // weighted sums
    BPT_HD Pt3 operator*(float s) const { return {x * s, y * s, z * s}; }
    BPT_HD Pt3 operator/(float s) const { return {x / s, y / s, z / s}; }
};
BPT_HD Vec3 operator*(float s, Vec3 v) { return v * s; }
BPT_HD Nrm3 operator*(float s, Nrm3 v) { return v * s; }
BPT_HD Pt3 operator*(float s, Pt3 v) { return v * s; }
BPT_HD Vec3 V(Pt3 p) { return {p.x, p.y, p.z}; }
BPT_HD Vec3 V(Nrm3 p) { return {p.x, p.y, p.z}; }
BPT_HD Pt3 P(Vec3 p) { return {p.x, p.y, p.z}; }

BPT_HD Vec3 Abs(Vec3 v) { return {fabsf(v.x), fabsf(v.y), fabsf(v.z)}; }
BPT_HD Pt3 Abs(Pt3 v) { return {fabsf(v.x), fabsf(v.y), fabsf(v.z)}; }
BPT_HD Nrm3 Abs(Nrm3 v) { return {fabsf(v.x), fabsf(v.y), fabsf(v.z)}; }
BPT_HD float Dot(Vec3 v, Vec3 w) { return v.x * w.x + v.y * w.y + v.z * w.z; }
BPT_HD float Dot(Nrm3 n, Vec3 v) { return FMA(n.x, v.x, SumOfProducts(n.y, v.y, n.z, v.z)); }
BPT_HD float Dot(Vec3 v, Nrm3 n) { return FMA(n.x, v.x, SumOfProducts(n.y, v.y, n.z, v.z)); }
BPT_HD float Dot(Nrm3 a, Nrm3 b) { return FMA(a.x, b.x, SumOfProducts(a.y, b.y, a.z, b.z)); }
BPT_HD float AbsDot(Vec3 v, Vec3 w) { return fabsf(Dot(v, w)); }
BPT_HD float AbsDot(Nrm3 v, Vec3 w) { return fabsf(Dot(v, w)); }
BPT_HD float AbsDot(Vec3 v, Nrm3 w) { return fabsf(Dot(v, w)); }
BPT_HD float AbsDot(Nrm3 v, Nrm3 w) { return fabsf(Dot(v, w)); }
BPT_HD Vec3 Cross(Vec3 v, Vec3 w) {
    return {DifferenceOfProducts(v.y, w.z, v.z, w.y), DifferenceOfProducts(v.z, w.x, v.x, w.z),
            DifferenceOfProducts(v.x, w.y, v.y, w.x)};
}
BPT_HD Vec3 Cross(Nrm3 v, Vec3 w) { return Cross(V(v), w); }
BPT_HD Vec3 Cross(Vec3 v, Nrm3 w) { return Cross(v, V(w)); }
BPT_HD float LengthSquared(Vec3 v) { return Sqr(v.x) + Sqr(v.y) + Sqr(v.z); }
BPT_HD float LengthSquared(Nrm3 v) { return Sqr(v.x) + Sqr(v.y) + Sqr(v.z); }
BPT_HD float Length(Vec3 v) { return sqrtf(LengthSquared(v)); }
BPT_HD float Length(Nrm3 v) { return sqrtf(LengthSquared(v)); }
BPT_HD Vec3 Normalize(Vec3 v) { return v / Length(v); }
BPT_HD Nrm3 Normalize(Nrm3 v) { return v / Length(v); }
BPT_HD float DistanceSquared(Pt3 a, Pt3 b) { return LengthSquared(a - b); }
BPT_HD float Distance(Pt3 a, Pt3 b) { return Length(a - b); }
BPT_HD Nrm3 FaceForward(Nrm3 n, Vec3 v) { return (Dot(n, v) < 0.f) ? -n : n; }
BPT_HD Nrm3 FaceForward(Nrm3 n, Nrm3 v) { return (Dot(n, v) < 0.f) ? -n : n; }
BPT_HD Vec3 FaceForward(Vec3 n, Vec3 v) { return (Dot(n, v) < 0.f) ? -n : n; }
BPT_HD Vec3 FaceForward(Vec3 n, Nrm3 v) { return (Dot(n, v) < 0.f) ? -n : n; }
BPT_HD float MaxComponentValue(Vec3 v) { return stdmax(stdmax(v.x, v.y), v.z); }  // std::max({x,y,z})
BPT_HD int MaxComponentIndex(Vec3 v) { return (v.x > v.y) ? ((v.x > v.z) ? 0 : 2) : ((v.y > v.z) ? 1 : 2); }
BPT_HD Vec3 Permute(Vec3 v, int x, int y, int z) { return {v[x], v[y], v[z]}; }
BPT_HD Pt3 Permute(Pt3 v, int x, int y, int z) { return {v[x], v[y], v[z]}; }
BPT_HD void CoordinateSystem(Vec3 v1, Vec3 *v2, Vec3 *v3) {
    float sign = copysignf(1.f, v1.z);
    float a = -1 / (sign + v1.z);
    float b = v1.x * v1.y * a;
    *v2 = Vec3(1 + sign * Sqr(v1.x) * a, sign * b, -sign * v1.x);
    *v3 = Vec3(b, sign + Sqr(v1.y) * a, -v1.y);
}
BPT_HD Vec3 GramSchmidt(Vec3 v, Vec3 w) { return v - Dot(v, w) * w; }
BPT_HD float SafeASin(float x) { return ASinF(Clampf(x, -1, 1)); }
BPT_HD float SafeACos(float x) { return ACosF(Clampf(x, -1, 1)); }
BPT_HD float AngleBetween(Vec3 v1, Vec3 v2) {
    if (Dot(v1, v2) < 0) return kPi - 2 * SafeASin(Length(v1 + v2) / 2);
    else return 2 * SafeASin(Length(v2 - v1) / 2);
}

// Point3fi (vecmath.h: Point3<Interval>)
struct Pt3i {
    Interval x, y, z;
    BPT_HD Pt3i() {}
    BPT_HD explicit Pt3i(Pt3 p) : x(p.x), y(p.y), z(p.z) {}
    BPT_HD Pt3i(Pt3 p, Vec3 e)
        : x(Interval::FromValueAndError(p.x, e.x)), y(Interval::FromValueAndError(p.y, e.y)),
          z(Interval::FromValueAndError(p.z, e.z)) {}
    BPT_HD Pt3 p() const { return {x.Midpoint(), y.Midpoint(), z.Midpoint()}; }
    BPT_HD Vec3 Error() const { return {x.Width() / 2, y.Width() / 2, z.Width() / 2}; }
    BPT_HD bool IsExact() const { return x.Width() == 0 && y.Width() == 0 && z.Width() == 0; }
};

// Frame (util/vecmath.h Frame::FromZ / FromXZ, ToLocal, FromLocal)
struct Frame {
    Vec3 x, y, z;
    BPT_HD static Frame FromXZ(Vec3 x, Vec3 z) { Frame f; f.x = x; f.y = Cross(z, x); f.z = z; return f; }
    BPT_HD static Frame FromZ(Vec3 z) { Frame f; f.z = z; CoordinateSystem(z, &f.x, &f.y); return f; }
    BPT_HD Vec3 ToLocal(Vec3 v) const { return {Dot(v, x), Dot(v, y), Dot(v, z)}; }
    BPT_HD Vec3 FromLocal(Vec3 v) const { return v.x * x + v.y * y + v.z * z; }
};

// ---------------------------------------------------------------------------------------------
// Ray-origin offsetting (ray.h:77-108)
BPT_HD Pt3 OffsetRayOrigin(Pt3i pi, Nrm3 n, Vec3 w) {
    // Find vector _offset_ to corner of error bounds and compute initial _po_
    float d = Dot(Abs(n), pi.Error());
    Vec3 offset = d * V(n);
    if (Dot(w, n) < 0) offset = -offset;
    Pt3 po = pi.p() + offset;
    // Round offset point _po_ away from _p_
    if (offset.x > 0) po.x = NextFloatUp(po.x); else if (offset.x < 0) po.x = NextFloatDown(po.x);
    if (offset.y > 0) po.y = NextFloatUp(po.y); else if (offset.y < 0) po.y = NextFloatDown(po.y);
    if (offset.z > 0) po.z = NextFloatUp(po.z); else if (offset.z < 0) po.z = NextFloatDown(po.z);
    return po;
}

}  // namespace bpt
