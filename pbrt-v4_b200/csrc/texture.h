// texture.h -- image textures on the shading path: MIP pyramid lookups and the "imagemap" float / spectrum textures with (u,v)
// mapping, restated from the reference so that a lookup returns the reference's bits (host + device; pinned on the host by
// tests/tools/texture_host_check.cpp).
//   MIPMap::Filter / Texel / Bilerp / EWA          util/mipmap.cpp:208-347
//   Image::GetChannel / BilerpChannel              util/image.h:259-292, RemapPixelCoords :96-147
//   UVMapping::Map                                 textures.h:86-106
//   FloatImageTexture / SpectrumImageTexture       textures.h:571-626, textures.cpp:359-387
// The pyramid itself (Image::GeneratePyramid) is built on the host in texture_host.h; levels are stored decoded to float, which is
// value-identical to decoding 8-bit texels through the colour encoding's table at lookup time.
#pragma once
#include "scene.h"
#include "lights.h"

namespace bpt {

enum { TEXFILTER_POINT = 0, TEXFILTER_BILINEAR = 1, TEXFILTER_TRILINEAR = 2, TEXFILTER_EWA = 3 };   // FilterFunction, util/mipmap.h:21
enum { TEXWRAP_BLACK = 0, TEXWRAP_CLAMP = 1, TEXWRAP_REPEAT = 2 };                                  // WrapMode, util/image.h:53
enum { TEXSPEC_ILLUMINANT = 0, TEXSPEC_ALBEDO = 1, TEXSPEC_UNBOUNDED = 2 };                          // SpectrumType, util/spectrum.h

// TextureEvalContext (textures.h:30-69)
struct TexEvalCtx {
    Pt3 p;
    Vec3 dpdx, dpdy;
    Nrm3 n;
    Vec2 uv;
    float dudx = 0, dudy = 0, dvdx = 0, dvdy = 0;
};

BPT_HD int TexMod(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }

// Image::GetChannel(p, c, wrapMode) on one pyramid level
BPT_HD float MipTexelChannel(const DeviceMIPLevel &lv, int nc, int x, int y, int c, int wrap) {
    int p[2] = {x, y};
    const int res[2] = {lv.xres, lv.yres};
    for (int k = 0; k < 2; ++k) {
        if (p[k] >= 0 && p[k] < res[k]) continue;
        if (wrap == TEXWRAP_REPEAT) p[k] = TexMod(p[k], res[k]);
        else if (wrap == TEXWRAP_CLAMP) p[k] = Clampi(p[k], 0, res[k] - 1);
        else return 0;
    }
    return lv.texels[((size_t)p[1] * lv.xres + p[0]) * nc + c];
}

BPT_HD float MipBilerpChannel(const DeviceMIPLevel &lv, int nc, Vec2 st, int c, int wrap) {
    float x = st.x * lv.xres - 0.5f, y = st.y * lv.yres - 0.5f;
    int xi = (int)floorf(x), yi = (int)floorf(y);
    float dx = x - xi, dy = y - yi;
    float v0 = MipTexelChannel(lv, nc, xi, yi, c, wrap), v1 = MipTexelChannel(lv, nc, xi + 1, yi, c, wrap);
    float v2 = MipTexelChannel(lv, nc, xi, yi + 1, c, wrap), v3 = MipTexelChannel(lv, nc, xi + 1, yi + 1, c, wrap);
    return ((1 - dx) * (1 - dy) * v0 + dx * (1 - dy) * v1 + (1 - dx) * dy * v2 + dx * dy * v3);
}

// One value type for both instantiations of MIPMap::Filter<T>: RGB uses all three components, Float only v[0]
template <bool IS_RGB> struct TexVal {
    float v[3];
    BPT_HD TexVal() { v[0] = v[1] = v[2] = 0; }
    BPT_HD TexVal operator*(float a) const { TexVal r; r.v[0] = v[0] * a; if (IS_RGB) { r.v[1] = v[1] * a; r.v[2] = v[2] * a; } return r; }
    BPT_HD TexVal operator+(const TexVal &o) const { TexVal r; r.v[0] = v[0] + o.v[0]; if (IS_RGB) { r.v[1] = v[1] + o.v[1]; r.v[2] = v[2] + o.v[2]; } return r; }
    BPT_HD TexVal operator/(float a) const { TexVal r; r.v[0] = v[0] / a; if (IS_RGB) { r.v[1] = v[1] / a; r.v[2] = v[2] / a; } return r; }
};

template <bool IS_RGB> BPT_HD TexVal<IS_RGB> MipTexel(const DeviceMIPLevel &lv, int nc, int x, int y, int wrap) {
    TexVal<IS_RGB> r;
    if (!IS_RGB) { r.v[0] = MipTexelChannel(lv, nc, x, y, 0, wrap); return r; }   // Texel<Float>: channel 0 whatever the image holds
    if (nc == 3 || nc == 4) { for (int c = 0; c < 3; ++c) r.v[c] = MipTexelChannel(lv, nc, x, y, c, wrap); }
    else { float v = MipTexelChannel(lv, nc, x, y, 0, wrap); r.v[0] = r.v[1] = r.v[2] = v; }
    return r;
}
template <bool IS_RGB> BPT_HD TexVal<IS_RGB> MipBilerp(const DeviceMIPLevel &lv, int nc, Vec2 st, int wrap) {
    TexVal<IS_RGB> r;
    if (IS_RGB) {
        if (nc == 3 || nc == 4) { for (int c = 0; c < 3; ++c) r.v[c] = MipBilerpChannel(lv, nc, st, c, wrap); }
        else { float v = MipBilerpChannel(lv, nc, st, 0, wrap); r.v[0] = r.v[1] = r.v[2] = v; }
        return r;
    }
    if (nc == 1) r.v[0] = MipBilerpChannel(lv, nc, st, 0, wrap);
    else if (nc == 3) {   // Image::Bilerp(st).Average()
        float sum = 0;
        for (int c = 0; c < 3; ++c) sum += MipBilerpChannel(lv, nc, st, c, wrap);
        r.v[0] = sum / 3;
    } else r.v[0] = MipBilerpChannel(lv, nc, st, 3, wrap);   // alpha
    return r;
}
template <bool IS_RGB> BPT_HD TexVal<IS_RGB> TexLerp(float t, const TexVal<IS_RGB> &a, const TexVal<IS_RGB> &b) { return a * (1 - t) + b * t; }

template <bool IS_RGB>
BPT_HD TexVal<IS_RGB> MipEWA(const DeviceTexture &tex, const DeviceMIPLevel *levels, const float *ewaLUT, int level, Vec2 st, Vec2 dst0,
                             Vec2 dst1) {
    if (level >= tex.nLevels) return MipTexel<IS_RGB>(levels[tex.nLevels - 1], tex.nc, 0, 0, tex.wrap);
    const DeviceMIPLevel &lv = levels[level];
    st.x = st.x * lv.xres - 0.5f; st.y = st.y * lv.yres - 0.5f;
    dst0.x *= lv.xres; dst0.y *= lv.yres; dst1.x *= lv.xres; dst1.y *= lv.yres;
    float A = Sqr(dst0.y) + Sqr(dst1.y) + 1;
    float B = -2 * (dst0.x * dst0.y + dst1.x * dst1.y);
    float C = Sqr(dst0.x) + Sqr(dst1.x) + 1;
    float invF = 1 / (A * C - Sqr(B) * 0.25f);
    A *= invF; B *= invF; C *= invF;
    float det = -Sqr(B) + 4 * A * C;
    float invDet = 1 / det;
    float uSqrt = SafeSqrt(det * C), vSqrt = SafeSqrt(A * det);
    int s0 = (int)ceilf(st.x - 2 * invDet * uSqrt), s1 = (int)floorf(st.x + 2 * invDet * uSqrt);
    int t0 = (int)ceilf(st.y - 2 * invDet * vSqrt), t1 = (int)floorf(st.y + 2 * invDet * vSqrt);
    TexVal<IS_RGB> sum;
    float sumWts = 0;
    for (int it = t0; it <= t1; ++it) {
        float tt = it - st.y;
        for (int is = s0; is <= s1; ++is) {
            float ss = is - st.x;
            float r2 = A * Sqr(ss) + B * ss * tt + C * Sqr(tt);
            if (r2 < 1) {
                int index = imin((int)(r2 * 128), 127);
                float weight = ewaLUT[index];
                sum = sum + MipTexel<IS_RGB>(lv, tex.nc, is, it, tex.wrap) * weight;
                sumWts += weight;
            }
        }
    }
    return sum / sumWts;
}

// Outlined on the device (one copy per value type instead of one per call site: bump mapping alone looks the displacement up three
// times), which also keeps the shading kernels' compile time and register allocation in hand.
#if defined(__CUDACC__)
#define BPT_TEX_FN __host__ __device__ __noinline__
#else
#define BPT_TEX_FN inline
#endif
template <bool IS_RGB>
BPT_TEX_FN TexVal<IS_RGB> MipFilter(const DeviceTexture &tex, const DeviceMIPLevel *allLevels, const float *ewaLUT, Vec2 st, Vec2 dst0, Vec2 dst1) {
    const DeviceMIPLevel *levels = allLevels + tex.firstLevel;
    const int nLevels = tex.nLevels;
    if (tex.filter != TEXFILTER_EWA) {
        float width = 2 * stdmax(stdmax(fabsf(dst0.x), fabsf(dst0.y)), stdmax(fabsf(dst1.x), fabsf(dst1.y)));
        float level = nLevels - 1 + LogF(stdmax(width, 1e-8f)) * 1.442695040888963387004650940071f;
        if (level >= nLevels - 1) return MipTexel<IS_RGB>(levels[nLevels - 1], tex.nc, 0, 0, tex.wrap);
        int iLevel = imax(0, (int)floorf(level));
        if (tex.filter == TEXFILTER_POINT) {
            const DeviceMIPLevel &lv = levels[iLevel];
            return MipTexel<IS_RGB>(lv, tex.nc, (int)roundf(st.x * lv.xres - 0.5f), (int)roundf(st.y * lv.yres - 0.5f), tex.wrap);
        }
        if (tex.filter == TEXFILTER_BILINEAR) return MipBilerp<IS_RGB>(levels[iLevel], tex.nc, st, tex.wrap);
        if (iLevel == 0) return MipBilerp<IS_RGB>(levels[0], tex.nc, st, tex.wrap);
        return TexLerp(level - iLevel, MipBilerp<IS_RGB>(levels[iLevel], tex.nc, st, tex.wrap), MipBilerp<IS_RGB>(levels[iLevel + 1], tex.nc, st, tex.wrap));
    }
    if (dst0.x * dst0.x + dst0.y * dst0.y < dst1.x * dst1.x + dst1.y * dst1.y) { Vec2 t = dst0; dst0 = dst1; dst1 = t; }
    float longerVecLength = sqrtf(dst0.x * dst0.x + dst0.y * dst0.y), shorterVecLength = sqrtf(dst1.x * dst1.x + dst1.y * dst1.y);
    if (shorterVecLength * tex.maxAnisotropy < longerVecLength && shorterVecLength > 0) {
        float scale = longerVecLength / (shorterVecLength * tex.maxAnisotropy);
        dst1.x *= scale; dst1.y *= scale;
        shorterVecLength *= scale;
    }
    if (shorterVecLength == 0) return MipBilerp<IS_RGB>(levels[0], tex.nc, st, tex.wrap);
    float lod = stdmax(0.f, nLevels - 1 + LogF(shorterVecLength) * 1.442695040888963387004650940071f);
    int ilod = (int)floorf(lod);
    return TexLerp(lod - ilod, MipEWA<IS_RGB>(tex, levels, ewaLUT, ilod, st, dst0, dst1), MipEWA<IS_RGB>(tex, levels, ewaLUT, ilod + 1, st, dst0, dst1));
}

struct TexCoord2D { Vec2 st; float dsdx, dsdy, dtdx, dtdy; };
// SphericalMapping / CylindricalMapping / PlanarMapping::Map (textures.h:117-140, 151-165, 177-189).  A real function: inlined at
// every texture evaluation of the unrolled texture graphs it multiplied the size (and the ptxas time) of the TEX shading kernels.
static BPT_TEX_FN TexCoord2D PointMap(const DeviceTexture &t, Pt3 p, Vec3 dpdxR, Vec3 dpdyR) {
    TexCoord2D c;
        const float (*m)[4] = t.textureFromRender.m;
        auto xp = [&](Pt3 p) {   // Transform::operator()(Point3f)
            float x = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3], y = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
            float z = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3], w = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
            return w == 1 ? Pt3(x, y, z) : Pt3(x, y, z) / w;
        };
        auto xv = [&](Vec3 v) { return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z, m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z); };
        Pt3 pt = xp(p);
        Vec3 dpdx = xv(dpdxR), dpdy = xv(dpdyR);
        if (t.mapping == 3) {
            Vec3 vec(pt.x, pt.y, pt.z);
            c.dsdx = Dot(t.vs, dpdx); c.dsdy = Dot(t.vs, dpdy); c.dtdx = Dot(t.vt, dpdx); c.dtdy = Dot(t.vt, dpdy);
            c.st = Vec2(t.du + Dot(vec, t.vs), t.dv + Dot(vec, t.vt));
            return c;
        }
        float x2y2 = Sqr(pt.x) + Sqr(pt.y);
        Vec3 dsdp = Vec3(-pt.y, pt.x, 0) / (2 * kPi * x2y2);
        Vec3 dtdp(0, 0, 1);
        if (t.mapping == 1) {
            float sqrtx2y2 = sqrtf(x2y2);
            dtdp = 1 / (kPi * (x2y2 + Sqr(pt.z))) * Vec3(pt.x * pt.z / sqrtx2y2, pt.y * pt.z / sqrtx2y2, -sqrtx2y2);
        }
        c.dsdx = Dot(dsdp, dpdx); c.dsdy = Dot(dsdp, dpdy); c.dtdx = Dot(dtdp, dpdx); c.dtdy = Dot(dtdp, dpdy);
        if (t.mapping == 1) {
            Vec3 vec = Normalize(pt - Pt3(0, 0, 0));
            float phi = ATan2F(vec.y, vec.x);                       // SphericalPhi
            c.st = Vec2(SafeACos(vec.z) * kInvPi, ((phi < 0) ? (phi + 2 * kPi) : phi) * kInv2Pi);
        } else
            c.st = Vec2((kPi + ATan2F(pt.y, pt.x)) * kInv2Pi, pt.z);
        return c;
}
BPT_HD TexCoord2D UVMap(const DeviceTexture &t, const TexEvalCtx &ctx) {
    if (t.mapping != 0) return PointMap(t, ctx.p, ctx.dpdx, ctx.dpdy);
    TexCoord2D c;
    c.dsdx = t.su * ctx.dudx; c.dsdy = t.su * ctx.dudy;
    c.dtdx = t.sv * ctx.dvdx; c.dtdy = t.sv * ctx.dvdy;
    c.st = Vec2(t.su * ctx.uv.x + t.du, t.sv * ctx.uv.y + t.dv);
    return c;
}

// FloatImageTexture::Evaluate (textures.h:579-593)
BPT_HD float EvalFloatImageTexture(const SceneView &sv, const DeviceTexture &t, const TexEvalCtx &ctx) {
    TexCoord2D c = UVMap(t, ctx);
    c.st.y = 1 - c.st.y;
    float v = t.scale * MipFilter<false>(t, sv.mipLevels, sv.ewaLUT, c.st, Vec2(c.dsdx, c.dtdx), Vec2(c.dsdy, c.dtdy)).v[0];
    return t.invert ? stdmax(0.f, 1 - v) : v;
}

// SpectrumImageTexture::Evaluate (textures.cpp:359-387); the image's colour space is sRGB (every reader's default)
BPT_HD Spec EvalSpectrumImageTexture(const SceneView &sv, const DeviceTexture &t, const TexEvalCtx &ctx, const Wavelengths &wl, int spectrumType) {
    TexCoord2D c = UVMap(t, ctx);
    c.st.y = 1 - c.st.y;
    TexVal<true> f = MipFilter<true>(t, sv.mipLevels, sv.ewaLUT, c.st, Vec2(c.dsdx, c.dtdx), Vec2(c.dsdy, c.dtdy));
    float rgb[3];
    for (int k = 0; k < 3; ++k) {
        float v = t.scale * f.v[k];
        if (t.invert) v = 1 - v;
        rgb[k] = stdmax(0.f, v);   // ClampZero
    }
    Spec s;
    float co[3];
    if (spectrumType == TEXSPEC_ALBEDO) {
        float q[3] = {Clampf(rgb[0], 0, 1), Clampf(rgb[1], 0, 1), Clampf(rgb[2], 0, 1)};
        RGBToCoeffsDev(sv, q, co);
        for (int i = 0; i < 4; ++i) s[i] = SigmoidPoly(co[0], co[1], co[2], wl.lambda[i]);
        return s;
    }
    // RGBUnboundedSpectrum / RGBIlluminantSpectrum (util/spectrum.cpp): scale = 2 max(r, g, b), coefficients of rgb / scale
    float m = stdmax(stdmax(rgb[0], rgb[1]), rgb[2]);
    float scale = 2 * m;
    float q[3] = {0, 0, 0};
    if (scale) { q[0] = rgb[0] / scale; q[1] = rgb[1] / scale; q[2] = rgb[2] / scale; }
    RGBToCoeffsDev(sv, q, co);
    for (int i = 0; i < 4; ++i) s[i] = scale * SigmoidPoly(co[0], co[1], co[2], wl.lambda[i]);
    if (spectrumType == TEXSPEC_ILLUMINANT) s = s * DenseSample(sv.illumD65, wl);
    return s;
}

// ---- texture graphs: "constant", "scale", "mix" over image textures (FloatScaledTexture / SpectrumScaledTexture textures.h:843-890,
// FloatMixTexture / SpectrumMixTexture textures.h:199-260).  The engine bounds the depth at commit (kTexMaxDepth); the recursion is
// unrolled by the template.
constexpr int kTexMaxDepth = 3;
// Checkerboard(ctx, map2D, nullptr) (textures.cpp:183-207): the box-filtered weight of the second texture
BPT_HD float CheckerboardWeight(const DeviceTexture &t, const TexEvalCtx &ctx) {
    auto d = [](float x) { float y = x / 2 - floorf(x / 2) - 0.5f; return x / 2 + y * (1 - 2 * fabsf(y)); };
    auto bf = [&](float x, float r) -> float {
        if (floorf(x - r) == floorf(x + r)) return (float)(1 - 2 * ((int)floorf(x) & 1));
        return (d(x + r) - 2 * d(x) + d(x - r)) / Sqr(r);
    };
    TexCoord2D c = UVMap(t, ctx);
    float ds = stdmax(fabsf(c.dsdx), fabsf(c.dsdy)), dt = stdmax(fabsf(c.dtdx), fabsf(c.dtdy));
    ds *= 1.5f; dt *= 1.5f;
    return 0.5f - bf(c.st.x, ds) * bf(c.st.y, dt) / 2;
}
template <int DEPTH> BPT_HD float EvalFloatTexture(const SceneView &sv, int id, const TexEvalCtx &ctx) {
    const DeviceTexture &t = sv.textures[id];
    if (t.kind == TEXKIND_IMAGEMAP) return EvalFloatImageTexture(sv, t, ctx);
    if (t.kind == TEXKIND_CONSTANT) return t.cval;
    if (DEPTH > 0) {
        constexpr int D = DEPTH > 0 ? DEPTH - 1 : 0;
        if (t.kind == TEXKIND_SCALE) {
            float sc = EvalFloatTexture<D>(sv, t.child1, ctx);
            if (sc == 0) return 0;
            return EvalFloatTexture<D>(sv, t.child0, ctx) * sc;
        }
        if (t.kind == TEXKIND_MIX || t.kind == TEXKIND_CHECKERBOARD) {
            float amt = t.kind == TEXKIND_MIX ? EvalFloatTexture<D>(sv, t.child2, ctx) : CheckerboardWeight(t, ctx);
            float t1 = 0, t2 = 0;
            if (amt != 1) t1 = EvalFloatTexture<D>(sv, t.child0, ctx);
            if (amt != 0) t2 = EvalFloatTexture<D>(sv, t.child1, ctx);
            return (1 - amt) * t1 + amt * t2;
        }
    }
    return 0;
}
template <int DEPTH> BPT_HD Spec EvalSpectrumTexture(const SceneView &sv, int id, const TexEvalCtx &ctx, const Wavelengths &wl, int spectrumType) {
    const DeviceTexture &t = sv.textures[id];
    if (t.kind == TEXKIND_IMAGEMAP) return EvalSpectrumImageTexture(sv, t, ctx, wl, spectrumType);
    if (t.kind == TEXKIND_CONSTANT) return EvalSpectrum(t.sval, sv.specPool, wl);
    if (DEPTH > 0) {
        constexpr int D = DEPTH > 0 ? DEPTH - 1 : 0;
        if (t.kind == TEXKIND_SCALE) {
            float sc = EvalFloatTexture<D>(sv, t.child1, ctx);
            if (sc == 0) return Spec(0.f);
            return EvalSpectrumTexture<D>(sv, t.child0, ctx, wl, spectrumType) * sc;
        }
        if (t.kind == TEXKIND_MIX || t.kind == TEXKIND_CHECKERBOARD) {
            float amt = t.kind == TEXKIND_MIX ? EvalFloatTexture<D>(sv, t.child2, ctx) : CheckerboardWeight(t, ctx);
            Spec t1(0.f), t2(0.f);
            if (amt != 1) t1 = EvalSpectrumTexture<D>(sv, t.child0, ctx, wl, spectrumType);
            if (amt != 0) t2 = EvalSpectrumTexture<D>(sv, t.child1, ctx, wl, spectrumType);
            return (1 - amt) * t1 + amt * t2;
        }
    }
    return Spec(0.f);
}

}  // namespace bpt
