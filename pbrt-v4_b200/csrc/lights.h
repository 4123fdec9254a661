// lights.h -- light sampling on the path (device + host).
//   LightSampleContext                         base/light.h / lights.h
//   DiffuseAreaLight::L / SampleLi / PDF_Li     lights.h:438-456, lights.cpp:739-767
//   CompactLightBounds::Importance              lightsamplers.h:144-200
//   BVHLightSampler::Sample / PMF               lightsamplers.h:267-353
//   UniformLightSampler                         lightsamplers.h:27-55
//   ImageInfiniteLight / UniformInfiniteLight   lights.h:500-628, lights.cpp
#pragma once
#include "scene.h"
#include "shapes.h"
#include "quadrics.h"
#include "bilinear.h"

namespace bpt {

struct LightSampleContext {
    Pt3i pi;
    Nrm3 n, ns;
    BPT_HD Pt3 p() const { return pi.p(); }
};

// Everything the kernels need to reach scene data
struct SceneView {
    const float4 *triPool;           // 3 x float4 per triangle, ORIGINAL order (p0.w = meshId, p1.w = triIndex)
    const DeviceMesh *meshes;
    const int *indexPool;
    const float *nPool;              // xyz triples
    const float *uvPool;             // uv pairs
    const DeviceMaterial *materials;
    const DeviceLight *lights;
    const float *specPool;           // dense tables / piecewise data
    DeviceLightSampler ls;
    Pt3 allbMin, allbMax;            // BVHLightSampler::allLightBounds
    const DeviceImageLight *imageLights;
    const DeviceInstance *instances;  // null: single-level scene
    const float *cieX, *cieY, *cieZ; // 471-entry densely sampled matching functions
    const float *illumD65;           // sRGB colour space illuminant (densely sampled)
    const float *rgb2specScale;      // 64 z-nodes
    const float *rgb2specCoeffs;     // [3][64][64][64][3]
    int nLights;
    const DeviceTexture *textures;    // image textures (null: none)
    const DeviceMIPLevel *mipLevels;
    const float *ewaLUT;              // MIPFilterLUT, 128 entries (util/mipmap.cpp:59)
    const DeviceShape *shapes;        // quadrics (null: none)
    const float4 *patches;            // bilinear patches: p00 | area, p10 | isRectangle, p01, p11 (rendering space)
};

#if defined(__CUDACC__)
BPT_D Pt3 ldP3(const float4 *p) { float4 v = __ldg(p); return Pt3(v.x, v.y, v.z); }
#endif

// Transform::operator()(const SurfaceInteraction &) (util/transform.cpp:207-245) for the members the path uses, and the
// wo the reference ends up with for an instanced hit: Normalize(M * (-(Minv * d)))  (cpu/primitive.cpp:114-123).
BPT_HD Pt3i XformPt3i(const Mat4 &T, const Pt3i &p) {
    const float (*m)[4] = T.m;
    float x = p.x.Midpoint(), y = p.y.Midpoint(), z = p.z.Midpoint();
    float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
    float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
    float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
    float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
    Vec3 pError;
    if (p.IsExact()) {
        pError.x = gamma_n(3) * (fabsf(m[0][0] * x) + fabsf(m[0][1] * y) + fabsf(m[0][2] * z) + fabsf(m[0][3]));
        pError.y = gamma_n(3) * (fabsf(m[1][0] * x) + fabsf(m[1][1] * y) + fabsf(m[1][2] * z) + fabsf(m[1][3]));
        pError.z = gamma_n(3) * (fabsf(m[2][0] * x) + fabsf(m[2][1] * y) + fabsf(m[2][2] * z) + fabsf(m[2][3]));
    } else {
        Vec3 pInError = p.Error();
        pError.x = (gamma_n(3) + 1) * (fabsf(m[0][0]) * pInError.x + fabsf(m[0][1]) * pInError.y + fabsf(m[0][2]) * pInError.z) +
                   gamma_n(3) * (fabsf(m[0][0] * x) + fabsf(m[0][1] * y) + fabsf(m[0][2] * z) + fabsf(m[0][3]));
        pError.y = (gamma_n(3) + 1) * (fabsf(m[1][0]) * pInError.x + fabsf(m[1][1]) * pInError.y + fabsf(m[1][2]) * pInError.z) +
                   gamma_n(3) * (fabsf(m[1][0] * x) + fabsf(m[1][1] * y) + fabsf(m[1][2] * z) + fabsf(m[1][3]));
        pError.z = (gamma_n(3) + 1) * (fabsf(m[2][0]) * pInError.x + fabsf(m[2][1]) * pInError.y + fabsf(m[2][2]) * pInError.z) +
                   gamma_n(3) * (fabsf(m[2][0] * x) + fabsf(m[2][1] * y) + fabsf(m[2][2] * z) + fabsf(m[2][3]));
    }
    Pt3i r(Pt3(xp, yp, zp), pError);
    if (wp != 1) {
        auto divI = [](Interval i, float f) { return f > 0 ? Interval(DivRoundDown(i.low, f), DivRoundUp(i.high, f))
                                                          : Interval(DivRoundDown(i.high, f), DivRoundUp(i.low, f)); };
        r.x = divI(r.x, wp); r.y = divI(r.y, wp); r.z = divI(r.z, wp);
    }
    return r;
}
BPT_HD Vec3 XformVec(const Mat4 &T, Vec3 v) {
    const float (*m)[4] = T.m;
    return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
}
BPT_HD Nrm3 XformNrm(const Mat4 &Tinv, Nrm3 n) {   // Transform::operator()(Normal3f): transpose of the inverse
    const float (*mi)[4] = Tinv.m;
    float x = n.x, y = n.y, z = n.z;
    return Nrm3(mi[0][0] * x + mi[1][0] * y + mi[2][0] * z, mi[0][1] * x + mi[1][1] * y + mi[2][1] * z,
                mi[0][2] * x + mi[1][2] * y + mi[2][2] * z);
}
BPT_HD SurfaceIntr XformInteraction(const DeviceInstance &I, const SurfaceIntr &si) {
    SurfaceIntr r;
    r.pi = XformPt3i(I.m, si.pi);
    r.n = Normalize(XformNrm(I.mInv, si.n));
    r.uv = si.uv;
    r.ns = Normalize(XformNrm(I.mInv, si.ns));
    r.dpdus = XformVec(I.m, si.dpdus);
    r.ns = FaceForward(r.ns, r.n);
    return r;
}
BPT_HD SurfaceExtras XformExtras(const DeviceInstance &I, const SurfaceExtras &e) {
    SurfaceExtras r;
    r.dpdu = XformVec(I.m, e.dpdu); r.dpdv = XformVec(I.m, e.dpdv); r.dpdvs = XformVec(I.m, e.dpdvs);
    r.dndus = XformNrm(I.mInv, e.dndus); r.dndvs = XformNrm(I.mInv, e.dndvs);
    return r;
}
BPT_HD Vec3 InstancedWo(const DeviceInstance &I, Vec3 rayD) {
    Vec3 dInst = XformVec(I.mInv, rayD);       // Transform::ApplyInverse(Vector3f)
    // si.wo = Normalize(-ray.d) in instance space (Interaction ctor, interaction.h:32-33), then ret.wo = Normalize(t(si.wo))
    return Normalize(XformVec(I.m, Normalize(-dInst)));
}

struct TriData { Pt3 p0, p1, p2; TriAttribs a; int mesh, triIndex; };

BPT_HD TriData LoadTriangle(const SceneView &sv, int tri) {
    TriData td;
    const float4 *t = sv.triPool + (size_t)tri * 3;
    float4 a = t[0], b = t[1], c = t[2];
    td.p0 = Pt3(a.x, a.y, a.z); td.p1 = Pt3(b.x, b.y, b.z); td.p2 = Pt3(c.x, c.y, c.z);
    td.mesh = (int)f2b(a.w);
    td.triIndex = (int)f2b(b.w);
    const DeviceMesh &m = sv.meshes[td.mesh];
    td.a.flipNormal = m.flags & 1;
    td.a.hasN = (m.flags & 2) != 0;
    td.a.hasUV = (m.flags & 4) != 0;
    if (td.a.hasN || td.a.hasUV) {
        const int *v = sv.indexPool + m.indexOffset + 3 * td.triIndex;
        int v0 = v[0], v1 = v[1], v2 = v[2];
        if (td.a.hasN) {
            const float *n = sv.nPool + 3 * (size_t)m.nOffset;
            td.a.n0 = Nrm3(n[3 * v0], n[3 * v0 + 1], n[3 * v0 + 2]);
            td.a.n1 = Nrm3(n[3 * v1], n[3 * v1 + 1], n[3 * v1 + 2]);
            td.a.n2 = Nrm3(n[3 * v2], n[3 * v2 + 1], n[3 * v2 + 2]);
        }
        if (td.a.hasUV) {
            const float *uv = sv.uvPool + 2 * (size_t)m.uvOffset;
            td.a.uv0 = Vec2(uv[2 * v0], uv[2 * v0 + 1]);
            td.a.uv1 = Vec2(uv[2 * v1], uv[2 * v1 + 1]);
            td.a.uv2 = Vec2(uv[2 * v2], uv[2 * v2 + 1]);
        }
    }
    return td;
}

// Patch i of a bilinear patch mesh with the per-vertex attributes of its mesh
BPT_HD PatchData LoadPatch(const SceneView &sv, const DeviceMesh &m, int i) {
    PatchData pd;
    const float4 *q = sv.patches + 4 * (size_t)(m.patchBase + i);
    float4 a = q[0], b = q[1], c = q[2], d = q[3];
    pd.p00 = Pt3(a.x, a.y, a.z); pd.p10 = Pt3(b.x, b.y, b.z); pd.p01 = Pt3(c.x, c.y, c.z); pd.p11 = Pt3(d.x, d.y, d.z);
    pd.area = a.w; pd.isRectangle = f2b(b.w) != 0;
    pd.flipNormal = m.flags & 1; pd.hasN = (m.flags & 2) != 0; pd.hasUV = (m.flags & 4) != 0;
    if (pd.hasN || pd.hasUV) {
        const int *v = sv.indexPool + m.indexOffset + 4 * i;
        int v0 = v[0], v1 = v[1], v2 = v[2], v3 = v[3];
        if (pd.hasN) {
            const float *n = sv.nPool + 3 * (size_t)m.nOffset;
            pd.n00 = Nrm3(n[3 * v0], n[3 * v0 + 1], n[3 * v0 + 2]); pd.n10 = Nrm3(n[3 * v1], n[3 * v1 + 1], n[3 * v1 + 2]);
            pd.n01 = Nrm3(n[3 * v2], n[3 * v2 + 1], n[3 * v2 + 2]); pd.n11 = Nrm3(n[3 * v3], n[3 * v3 + 1], n[3 * v3 + 2]);
        }
        if (pd.hasUV) {
            const float *uv = sv.uvPool + 2 * (size_t)m.uvOffset;
            pd.uv00 = Vec2(uv[2 * v0], uv[2 * v0 + 1]); pd.uv10 = Vec2(uv[2 * v1], uv[2 * v1 + 1]);
            pd.uv01 = Vec2(uv[2 * v2], uv[2 * v2 + 1]); pd.uv11 = Vec2(uv[2 * v3], uv[2 * v3 + 1]);
        }
    }
    return pd;
}

// SurfaceInteraction of a recorded hit: triangles record barycentrics, quadrics the object-space hit point (quadrics.h), bilinear
// patches (u, v)
// (SHAPES = false: the kernels of scenes without quadrics, compiled without the dispatch)
template <bool SHAPES>
BPT_HD SurfaceIntr HitInteraction(const SceneView &sv, const TriData &td, float b0, float b1, float b2, SurfaceExtras *ex = nullptr) {
    const int shape = SHAPES ? sv.meshes[td.mesh].shape : -1;
    if (SHAPES && shape >= 0) return QuadricInteraction(sv.shapes[shape], Pt3(b0, b1, b2), ex);
    if (SHAPES && sv.meshes[td.mesh].patchBase >= 0) return BilinearInteraction(LoadPatch(sv, sv.meshes[td.mesh], td.triIndex), b0, b1, ex);
    TriHit th{b0, b1, b2, 0.f};
    return TriangleInteraction(td.p0, td.p1, td.p2, td.a, th, ex);
}
// wo of that interaction in the space of the enclosing BVH, for a ray of direction d in that space
template <bool SHAPES>
BPT_HD Vec3 HitWo(const SceneView &sv, const TriData &td, Vec3 d) {
    const int shape = SHAPES ? sv.meshes[td.mesh].shape : -1;
    if (SHAPES && shape >= 0) return QuadricWo(sv.shapes[shape], d);
    return Normalize(-d);
}

// ---------------------------------------------------------------- DiffuseAreaLight
BPT_HD Spec AreaLightL(const SceneView &sv, const DeviceLight &light, Nrm3 n, Vec3 w, const Wavelengths &wl) {
    if (!light.twoSided && Dot(n, w) < 0) return Spec(0.f);
    return light.scale * DenseSample(sv.specPool + light.LemitTable, wl);
}

struct LightLiSample { Spec L; Vec3 wi; float pdf; Pt3i pLight; Nrm3 nLight; bool valid; };

template <bool SHAPES>
BPT_HD LightLiSample AreaLightSampleLi(const SceneView &sv, const DeviceLight &light, const LightSampleContext &ctx, Vec2 u,
                                       const Wavelengths &wl) {
    LightLiSample ls;
    ls.valid = false;
    TriData td = LoadTriangle(sv, light.tri);
    const int shape = SHAPES ? sv.meshes[td.mesh].shape : -1;
    ShapeSample ss;
    if (SHAPES && shape >= 0) ss = QuadricSampleSolidAngle(sv.shapes[shape], ctx.pi, ctx.n, u);
    else if (SHAPES && sv.meshes[td.mesh].patchBase >= 0) ss = PatchSampleSolidAngle(LoadPatch(sv, sv.meshes[td.mesh], td.triIndex), ctx.p(), ctx.ns, u);
    else ss = TriangleSampleSolidAngle(td.p0, td.p1, td.p2, td.a, ctx.p(), ctx.ns, u);
    if (!ss.valid || ss.pdf == 0 || LengthSquared(ss.pi.p() - ctx.p()) == 0) return ls;
    Vec3 wi = Normalize(ss.pi.p() - ctx.p());
    Spec Le = AreaLightL(sv, light, ss.n, -wi, wl);
    if (!Le.IsNonZero()) return ls;
    ls.L = Le; ls.wi = wi; ls.pdf = ss.pdf; ls.pLight = ss.pi; ls.nLight = ss.n; ls.valid = true;
    return ls;
}
template <bool SHAPES>
BPT_HD float AreaLightPDF_Li(const SceneView &sv, const DeviceLight &light, const LightSampleContext &ctx, Vec3 wi) {
    TriData td = LoadTriangle(sv, light.tri);
    const int shape = SHAPES ? sv.meshes[td.mesh].shape : -1;
    if (SHAPES && shape >= 0) return QuadricPDFSolidAngle(sv.shapes[shape], ctx.pi, ctx.n, wi);
    if (SHAPES && sv.meshes[td.mesh].patchBase >= 0) return PatchPDFSolidAngle(LoadPatch(sv, sv.meshes[td.mesh], td.triIndex), ctx.pi, ctx.n, ctx.ns, wi);
    return TrianglePDFSolidAngle(td.p0, td.p1, td.p2, td.a, ctx.pi, ctx.n, ctx.ns, wi);
}

// ---------------------------------------------------------------- infinite lights
// Equal-area octahedral mapping (util/math.cpp:292-383)
BPT_HD Vec3 EqualAreaSquareToSphere(Vec2 p) {
    float u = 2 * p.x - 1, v = 2 * p.y - 1;
    float up = fabsf(u), vp = fabsf(v);
    float signedDistance = 1 - (up + vp);
    float d = fabsf(signedDistance);
    float r = 1 - d;
    float phi = (r == 0 ? 1 : (vp - up) / r + 1) * kPi / 4;
    float z = copysignf(1 - Sqr(r), signedDistance);
    float cosPhi = copysignf(CosF(phi), u);
    float sinPhi = copysignf(SinF(phi), v);
    return Vec3(cosPhi * r * SafeSqrt(2 - Sqr(r)), sinPhi * r * SafeSqrt(2 - Sqr(r)), z);
}
BPT_HD Vec2 EqualAreaSphereToSquare(Vec3 d) {
    float x = fabsf(d.x), y = fabsf(d.y), z = fabsf(d.z);
    float r = SafeSqrt(1 - z);
    float a = stdmax(x, y), b = stdmin(x, y);
    b = a == 0 ? 0 : b / a;
    const float t1 = 0.406758566246788489601959989e-5f, t2 = 0.636226545274016134946890922156f, t3 = 0.61572017898280213493197203466e-2f,
                t4 = -0.247333733281268944196501420480f, t5 = 0.881770664775316294736387951347e-1f,
                t6 = 0.419038818029165735901852432784e-1f, t7 = -0.251390972343483509333252996350e-1f;
    float phi = EvalPoly(b, t1, t2, t3, t4, t5, t6, t7);
    if (x < y) phi = 1 - phi;
    float v = phi * r;
    float u = r - v;
    if (d.z < 0) { float t = u; u = v; v = t; u = 1 - u; v = 1 - v; }
    u = copysignf(u, d.x);
    v = copysignf(v, d.y);
    return Vec2(0.5f * (u + 1), 0.5f * (v + 1));
}
// RGBToSpectrumTable::operator() on the device (util/color.cpp:31-68)
BPT_HD void RGBToCoeffsDev(const SceneView &sv, const float rgb[3], float c[3]) {
    const int res = 64;
    if (rgb[0] == rgb[1] && rgb[1] == rgb[2]) { c[0] = 0; c[1] = 0; c[2] = (rgb[0] - .5f) / sqrtf(rgb[0] * (1 - rgb[0])); return; }
    int maxc = (rgb[0] > rgb[1]) ? ((rgb[0] > rgb[2]) ? 0 : 2) : ((rgb[1] > rgb[2]) ? 1 : 2);
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * (res - 1) / z;
    float y = rgb[(maxc + 2) % 3] * (res - 1) / z;
    int xi = imin((int)x, res - 2), yi = imin((int)y, res - 2);
    int zi = FindIntervalLT(sv.rgb2specScale, res, z);
    float dx = x - xi, dy = y - yi, dz = (z - sv.rgb2specScale[zi]) / (sv.rgb2specScale[zi + 1] - sv.rgb2specScale[zi]);
    const float *co = sv.rgb2specCoeffs;
    for (int i = 0; i < 3; ++i) {
#define BPT_CO(ddx, ddy, ddz) co[((((size_t)maxc * res + (zi + ddz)) * res + (yi + ddy)) * res + (xi + ddx)) * 3 + i]
        c[i] = Lerp(dz, Lerp(dy, Lerp(dx, BPT_CO(0, 0, 0), BPT_CO(1, 0, 0)), Lerp(dx, BPT_CO(0, 1, 0), BPT_CO(1, 1, 0))),
                    Lerp(dy, Lerp(dx, BPT_CO(0, 0, 1), BPT_CO(1, 0, 1)), Lerp(dx, BPT_CO(0, 1, 1), BPT_CO(1, 1, 1))));
#undef BPT_CO
    }
}
BPT_HD Vec3 ApplyM3(const Mat4 &T, Vec3 v) {
    const float (*m)[4] = T.m;
    return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
}
// ImageInfiniteLight::ImageLe (lights.h:612-618)
BPT_HD Spec ImageLe(const SceneView &sv, const DeviceLight &light, Vec2 uv, const Wavelengths &wl) {
    const DeviceImageLight &im = sv.imageLights[light.image];
    int res = im.res;
    int px = (int)(uv.x * res), py = (int)(uv.y * res);
    // RemapPixelCoords, WrapMode::OctahedralSphere (util/image.h:100-125)
    if (px < 0) { px = -px; py = res - 1 - py; } else if (px >= res) { px = 2 * res - 1 - px; py = res - 1 - py; }
    if (py < 0) { px = res - 1 - px; py = -py; } else if (py >= res) { px = res - 1 - px; py = 2 * res - 1 - py; }
    if (res == 1) { px = 0; py = 0; }
    const float *pix = im.rgb + 3 * ((size_t)py * res + px);
    float rgb[3] = {stdmax(0.f, pix[0]), stdmax(0.f, pix[1]), stdmax(0.f, pix[2])};
    // RGBIlluminantSpectrum(cs, rgb).Sample(lambda)
    float m = stdmax(stdmax(rgb[0], rgb[1]), rgb[2]);
    float scale = 2 * m;
    float q[3] = {0, 0, 0};
    if (scale) { q[0] = rgb[0] / scale; q[1] = rgb[1] / scale; q[2] = rgb[2] / scale; }
    float c[3];
    RGBToCoeffsDev(sv, q, c);
    Spec s;
    for (int i = 0; i < 4; ++i) s[i] = scale * SigmoidPoly(c[0], c[1], c[2], wl.lambda[i]);
    return light.scale * (s * DenseSample(sv.illumD65, wl));
}
BPT_HD Spec InfiniteLightLe(const SceneView &sv, const DeviceLight &light, Vec3 rd, const Wavelengths &wl) {
    // a DistantLight has no bounds and therefore sits in the light sampler's list of infinite lights (lightsamplers.cpp:63-72), but it
    // is not LightType::Infinite: the integrator's own list (cpu/integrators.cpp:79-83) leaves it out, escaped rays see nothing of it
    if (light.type == LIGHT_DISTANT) return Spec(0.f);
    if (light.type == LIGHT_UNIFORM_INFINITE) return light.scale * DenseSample(sv.specPool + light.LemitTable, wl);
    Vec3 wLight = Normalize(ApplyM3(light.lightFromRender, rd));
    return ImageLe(sv, light, EqualAreaSphereToSquare(wLight), wl);
}
// allowIncompletePDF is true at both call sites on this path (cpu/integrators.cpp:1092, 1303)
BPT_HD float InfiniteLightPDF_Li(const SceneView &sv, const DeviceLight &light, const LightSampleContext &ctx, Vec3 w) {
    if (light.type == LIGHT_UNIFORM_INFINITE) return 0;
    Vec3 wLight = ApplyM3(light.lightFromRender, w);
    Vec2 uv = EqualAreaSphereToSquare(wLight);
    return PC2DPDF(sv.imageLights[light.image].compensatedDistrib, uv) / (4 * kPi);
}
BPT_HD LightLiSample InfiniteLightSampleLi(const SceneView &sv, const DeviceLight &light, const LightSampleContext &ctx, Vec2 u,
                                           const Wavelengths &wl) {
    LightLiSample ls;
    ls.valid = false;
    if (light.type == LIGHT_UNIFORM_INFINITE) return ls;
    float mapPDF = 0; int ou, ov;
    Vec2 uv = PC2DSample(sv.imageLights[light.image].compensatedDistrib, u, &mapPDF, &ou, &ov);
    if (mapPDF == 0) return ls;
    Vec3 wLight = EqualAreaSquareToSphere(uv);
    Vec3 wi = ApplyM3(light.renderFromLight, wLight);
    ls.pdf = mapPDF / (4 * kPi);
    ls.L = ImageLe(sv, light, uv, wl);
    ls.wi = wi;
    ls.pLight = Pt3i(ctx.p() + wi * (2 * light.sceneRadius));
    ls.nLight = Nrm3(0, 0, 0);
    ls.valid = true;
    return ls;
}

// PointLight / SpotLight / DistantLight::SampleLi (lights.h:221-228, 771-783, 282-289; SpotLight::I lights.cpp:1360-1363):
// pdf = 1, the reference point carries no normal and no error bounds.
BPT_HD bool IsDeltaLight(int type) { return type == LIGHT_POINT || type == LIGHT_SPOT || type == LIGHT_DISTANT; }
// Outlined on the device: scenes without delta lights must not pay registers for it in the shading kernels (r02p: inlined, it cost
// k_shade<Diffuse> 11 % through spills).
#if defined(__CUDACC__)
#define BPT_OUTLINED_FN inline __host__ __device__ __noinline__
#else
#define BPT_OUTLINED_FN inline
#endif
BPT_OUTLINED_FN LightLiSample DeltaLightSampleLi(const float *specPool, const DeviceLight &light, Pt3 ctxP, const Wavelengths &wl) {
    LightLiSample ls;
    ls.valid = false;
    ls.pdf = 1;
    ls.nLight = Nrm3(0, 0, 0);
    if (light.type == LIGHT_DISTANT) {
        Vec3 wi = light.w;
        ls.pLight = Pt3i(ctxP + wi * (2 * light.sceneRadius));
        ls.L = light.scale * DenseSample(specPool + light.LemitTable, wl);
        ls.wi = wi;
        ls.valid = true;
        return ls;
    }
    Pt3 p = light.p;
    Vec3 wi = Normalize(p - ctxP);
    Spec I;
    if (light.type == LIGHT_SPOT) {
        Vec3 wLight = Normalize(ApplyM3(light.lightFromRender, -wi));
        float a = light.cosFalloffEnd, b = light.cosFalloffStart, x = wLight.z, smooth;
        if (a == b) smooth = (x < a) ? 0.f : 1.f;
        else {
            float t = Clampf((x - a) / (b - a), 0.f, 1.f);
            smooth = t * t * (3 - 2 * t);
        }
        I = (smooth * light.scale) * DenseSample(specPool + light.LemitTable, wl);
    } else
        I = light.scale * DenseSample(specPool + light.LemitTable, wl);
    ls.L = I / DistanceSquared(p, ctxP);
    if (light.type == LIGHT_SPOT && !ls.L.IsNonZero()) return ls;
    ls.wi = wi;
    ls.pLight = Pt3i(p);
    ls.valid = true;
    return ls;
}

// ---------------------------------------------------------------- light BVH
BPT_HD float BoundSubtendedCos(Pt3 bmin, Pt3 bmax, Pt3 p) {
    // BoundSubtendedDirections(b, p).cosTheta  (vecmath.h:1815-1828)
    Pt3 pCenter = (bmin + bmax) / 2;
    bool inside = pCenter.x >= bmin.x && pCenter.x <= bmax.x && pCenter.y >= bmin.y && pCenter.y <= bmax.y &&
                  pCenter.z >= bmin.z && pCenter.z <= bmax.z;
    float radius = inside ? Distance(pCenter, bmax) : 0;
    if (DistanceSquared(p, pCenter) < Sqr(radius)) return -1;
    float sin2ThetaMax = Sqr(radius) / DistanceSquared(pCenter, p);
    return SafeSqrt(1 - sin2ThetaMax);
}
BPT_HD float CosSubClamped(float sinTheta_a, float cosTheta_a, float sinTheta_b, float cosTheta_b) {
    if (cosTheta_a > cosTheta_b) return 1;
    return cosTheta_a * cosTheta_b + sinTheta_a * sinTheta_b;
}
BPT_HD float SinSubClamped(float sinTheta_a, float cosTheta_a, float sinTheta_b, float cosTheta_b) {
    if (cosTheta_a > cosTheta_b) return 0;
    return sinTheta_a * cosTheta_b - cosTheta_a * sinTheta_b;
}
BPT_HD float LightBoundsImportance(const DeviceLightBVHNode &b, Pt3 p, Nrm3 n) {
    float cosTheta_o = b.cosTheta_o, cosTheta_e = b.cosTheta_e;
    Pt3 pc = (b.bmin + b.bmax) / 2;
    float d2 = DistanceSquared(p, pc);
    d2 = stdmax(d2, Length(b.bmax - b.bmin) / 2);
    Vec3 wi = Normalize(p - pc);
    float cosTheta_w = Dot(b.w, wi);
    if (b.twoSided) cosTheta_w = fabsf(cosTheta_w);
    float sinTheta_w = SafeSqrt(1 - Sqr(cosTheta_w));
    float cosTheta_b = BoundSubtendedCos(b.bmin, b.bmax, p);
    float sinTheta_b = SafeSqrt(1 - Sqr(cosTheta_b));
    float sinTheta_o = SafeSqrt(1 - Sqr(cosTheta_o));
    float cosTheta_x = CosSubClamped(sinTheta_w, cosTheta_w, sinTheta_o, cosTheta_o);
    float sinTheta_x = SinSubClamped(sinTheta_w, cosTheta_w, sinTheta_o, cosTheta_o);
    float cosThetap = CosSubClamped(sinTheta_x, cosTheta_x, sinTheta_b, cosTheta_b);
    if (cosThetap <= cosTheta_e) return 0;
    float importance = b.phi * cosThetap / d2;
    if (n != Nrm3(0, 0, 0)) {
        float cosTheta_i = AbsDot(wi, n);
        float sinTheta_i = SafeSqrt(1 - Sqr(cosTheta_i));
        float cosThetap_i = CosSubClamped(sinTheta_i, cosTheta_i, sinTheta_b, cosTheta_b);
        importance *= cosThetap_i;
    }
    importance = stdmax(importance, 0.f);
    return importance;
}

BPT_HD int SampleDiscrete2(const float w[2], float u, float *pmf, float *uRemapped) {
    float sumWeights = 0;
    sumWeights += w[0];
    sumWeights += w[1];
    float up = u * sumWeights;
    if (up == sumWeights) up = NextFloatDown(up);
    int offset = 0;
    float sum = 0;
    while (sum + w[offset] <= up) sum += w[offset++];
    *pmf = w[offset] / sumWeights;
    *uRemapped = stdmin((up - sum) / w[offset], kOneMinusEpsilon);
    return offset;
}

// returns light index or -1
BPT_HD int LightSamplerSample(const SceneView &sv, const LightSampleContext &ctx, float u, float *pmfOut) {
    const DeviceLightSampler &ls = sv.ls;
    if (ls.type == LS_UNIFORM) {
        if (ls.nLights == 0) return -1;
        int lightIndex = imin((int)(u * ls.nLights), ls.nLights - 1);
        *pmfOut = 1.f / ls.nLights;
        return lightIndex;
    }
    float pInfinite = float(ls.nInfinite) / float(ls.nInfinite + (ls.nNodes == 0 ? 0 : 1));
    if (u < pInfinite) {
        u /= pInfinite;
        int index = imin((int)(u * ls.nInfinite), ls.nInfinite - 1);
        *pmfOut = pInfinite / ls.nInfinite;
        return ls.infiniteLights[index];
    }
    if (ls.nNodes == 0) return -1;
    Pt3 p = ctx.p();
    Nrm3 n = ctx.ns;
    u = stdmin((u - pInfinite) / (1 - pInfinite), kOneMinusEpsilon);
    int nodeIndex = 0;
    float pmf = 1 - pInfinite;
    while (true) {
        const DeviceLightBVHNode &node = ls.nodes[nodeIndex];
        if (!node.isLeaf) {
            float ci[2] = {LightBoundsImportance(ls.nodes[nodeIndex + 1], p, n),
                           LightBoundsImportance(ls.nodes[node.childOrLightIndex], p, n)};
            if (ci[0] == 0 && ci[1] == 0) return -1;
            float nodePMF;
            int child = SampleDiscrete2(ci, u, &nodePMF, &u);
            pmf *= nodePMF;
            nodeIndex = (child == 0) ? (nodeIndex + 1) : node.childOrLightIndex;
        } else {
            if (nodeIndex > 0 || LightBoundsImportance(node, p, n) > 0) { *pmfOut = pmf; return node.childOrLightIndex; }
            return -1;
        }
    }
}

BPT_HD float LightSamplerPMF(const SceneView &sv, const LightSampleContext &ctx, int light) {
    const DeviceLightSampler &ls = sv.ls;
    if (ls.type == LS_UNIFORM) return ls.nLights == 0 ? 0.f : 1.f / ls.nLights;
    if (!ls.lightInBVH[light]) return 1.f / (ls.nInfinite + (ls.nNodes == 0 ? 0 : 1));
    uint32_t bitTrail = ls.bitTrails[light];
    Pt3 p = ctx.p();
    Nrm3 n = ctx.ns;
    float pInfinite = float(ls.nInfinite) / float(ls.nInfinite + (ls.nNodes == 0 ? 0 : 1));
    float pmf = 1 - pInfinite;
    int nodeIndex = 0;
    while (true) {
        const DeviceLightBVHNode &node = ls.nodes[nodeIndex];
        if (node.isLeaf) return pmf;
        float ci[2] = {LightBoundsImportance(ls.nodes[nodeIndex + 1], p, n),
                       LightBoundsImportance(ls.nodes[node.childOrLightIndex], p, n)};
        pmf *= ci[bitTrail & 1] / (ci[0] + ci[1]);
        nodeIndex = (bitTrail & 1) ? node.childOrLightIndex : (nodeIndex + 1);
        bitTrail >>= 1;
    }
}

}  // namespace bpt
