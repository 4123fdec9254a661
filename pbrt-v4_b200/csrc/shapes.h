// shapes.h -- Triangle: watertight intersection, surface interaction, area / spherical sampling.
//   IntersectTriangle                           shapes.cpp:168-269
//   Triangle::InteractionFromIntersection       shapes.h:884-1010
//   Triangle::Sample(u), Sample(ctx,u), PDF     shapes.h:1013-1174
//   SurfaceInteraction ctor / SetShadingGeometry interaction.h
#pragma once
#include "pmath.h"
#include "sampling.h"

namespace bpt {

struct TriHit { float b0, b1, b2, t; };

// Watertight ray/triangle test; returns false on miss.  Identical operation order to the reference,
// including the double-precision fallback on a zero edge function and the conservative t > deltaT test.
BPT_HD bool IntersectTriangle(Pt3 o, Vec3 dir, float tMax, Pt3 p0, Pt3 p1, Pt3 p2, TriHit *hit) {
    if (LengthSquared(Cross(p2 - p0, p1 - p0)) == 0) return false;
    Vec3 p0t = p0 - o, p1t = p1 - o, p2t = p2 - o;
    int kz = MaxComponentIndex(Abs(dir));
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    Vec3 d = Permute(dir, kx, ky, kz);
    p0t = Permute(p0t, kx, ky, kz);
    p1t = Permute(p1t, kx, ky, kz);
    p2t = Permute(p2t, kx, ky, kz);
    float Sx = -d.x / d.z, Sy = -d.y / d.z, Sz = 1 / d.z;
    p0t.x += Sx * p0t.z; p0t.y += Sy * p0t.z;
    p1t.x += Sx * p1t.z; p1t.y += Sy * p1t.z;
    p2t.x += Sx * p2t.z; p2t.y += Sy * p2t.z;
    float e0 = DifferenceOfProducts(p1t.x, p2t.y, p1t.y, p2t.x);
    float e1 = DifferenceOfProducts(p2t.x, p0t.y, p2t.y, p0t.x);
    float e2 = DifferenceOfProducts(p0t.x, p1t.y, p0t.y, p1t.x);
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        double p2txp1ty = (double)p2t.x * (double)p1t.y;
        double p2typ1tx = (double)p2t.y * (double)p1t.x;
        e0 = (float)(p2typ1tx - p2txp1ty);
        double p0txp2ty = (double)p0t.x * (double)p2t.y;
        double p0typ2tx = (double)p0t.y * (double)p2t.x;
        e1 = (float)(p0typ2tx - p0txp2ty);
        double p1txp0ty = (double)p1t.x * (double)p0t.y;
        double p1typ0tx = (double)p1t.y * (double)p0t.x;
        e2 = (float)(p1typ0tx - p1txp0ty);
    }
    if ((e0 < 0 || e1 < 0 || e2 < 0) && (e0 > 0 || e1 > 0 || e2 > 0)) return false;
    float det = e0 + e1 + e2;
    if (det == 0) return false;
    p0t.z *= Sz; p1t.z *= Sz; p2t.z *= Sz;
    float tScaled = e0 * p0t.z + e1 * p1t.z + e2 * p2t.z;
    if (det < 0 && (tScaled >= 0 || tScaled < tMax * det)) return false;
    else if (det > 0 && (tScaled <= 0 || tScaled > tMax * det)) return false;
    float invDet = 1 / det;
    float b0 = e0 * invDet, b1 = e1 * invDet, b2 = e2 * invDet;
    float t = tScaled * invDet;
    float maxZt = MaxComponentValue(Abs(Vec3(p0t.z, p1t.z, p2t.z)));
    float deltaZ = gamma_n(3) * maxZt;
    float maxXt = MaxComponentValue(Abs(Vec3(p0t.x, p1t.x, p2t.x)));
    float maxYt = MaxComponentValue(Abs(Vec3(p0t.y, p1t.y, p2t.y)));
    float deltaX = gamma_n(5) * (maxXt + maxZt);
    float deltaY = gamma_n(5) * (maxYt + maxZt);
    float deltaE = 2 * (gamma_n(2) * maxXt * maxYt + deltaY * maxXt + deltaX * maxYt);
    float maxE = MaxComponentValue(Abs(Vec3(e0, e1, e2)));
    float deltaT = 3 * (gamma_n(3) * maxE * maxZt + deltaE * maxZt + deltaZ * maxE) * fabsf(invDet);
    if (t <= deltaT) return false;
    hit->b0 = b0; hit->b1 = b1; hit->b2 = b2; hit->t = t;
    return true;
}

// What the shading kernels need of a SurfaceInteraction.
struct SurfaceIntr {
    Pt3i pi;
    Nrm3 n;        // geometric normal
    Nrm3 ns;       // shading.n
    Vec3 dpdus;    // shading.dpdu
    Vec2 uv;
    BPT_HD Pt3 p() const { return pi.p(); }
};

// Optional per-vertex attributes of the triangle's mesh
struct TriAttribs {
    bool hasN, hasUV, flipNormal;   // flipNormal = reverseOrientation ^ transformSwapsHandedness
    Nrm3 n0, n1, n2;
    Vec2 uv0, uv1, uv2;
};

// The members of SurfaceInteraction only texture filtering and bump mapping read (interaction.h: dpdu, dpdv, shading.dpdv,
// shading.dndu / dndv); filled when the caller passes a record
struct SurfaceExtras {
    Vec3 dpdu, dpdv;     // geometric
    Vec3 dpdvs;          // shading.dpdv
    Nrm3 dndus, dndvs;   // shading.dndu, shading.dndv
};

BPT_HD SurfaceIntr TriangleInteraction(Pt3 p0, Pt3 p1, Pt3 p2, const TriAttribs &a, TriHit ti, SurfaceExtras *ex = nullptr) {
    Vec2 uv0(0, 0), uv1(1, 0), uv2(1, 1);
    if (a.hasUV) { uv0 = a.uv0; uv1 = a.uv1; uv2 = a.uv2; }
    Vec2 duv02(uv0.x - uv2.x, uv0.y - uv2.y), duv12(uv1.x - uv2.x, uv1.y - uv2.y);
    Vec3 dp02 = p0 - p2, dp12 = p1 - p2;
    float determinant = DifferenceOfProducts(duv02.x, duv12.y, duv02.y, duv12.x);
    Vec3 dpdu, dpdv;
    bool degenerateUV = fabsf(determinant) < 1e-9f;
    if (!degenerateUV) {
        float invdet = 1 / determinant;
        // DifferenceOfProducts(Float, Vector3f, Float, Vector3f) is evaluated per component
        dpdu = Vec3(DifferenceOfProducts(duv12.y, dp02.x, duv02.y, dp12.x), DifferenceOfProducts(duv12.y, dp02.y, duv02.y, dp12.y),
                    DifferenceOfProducts(duv12.y, dp02.z, duv02.y, dp12.z)) * invdet;
        dpdv = Vec3(DifferenceOfProducts(duv02.x, dp12.x, duv12.x, dp02.x), DifferenceOfProducts(duv02.x, dp12.y, duv12.x, dp02.y),
                    DifferenceOfProducts(duv02.x, dp12.z, duv12.x, dp02.z)) * invdet;
    }
    if (degenerateUV || LengthSquared(Cross(dpdu, dpdv)) == 0) {
        Vec3 ng = Cross(p2 - p0, p1 - p0);
        if (LengthSquared(ng) == 0) {
            // double-precision cross product (Vector3<double> Cross uses DifferenceOfProducts in double)
            Vec3 a2 = p2 - p0, b2 = p1 - p0;
            double ax = a2.x, ay = a2.y, az = a2.z, bx = b2.x, by = b2.y, bz = b2.z;
            auto dop = [](double a, double b, double c, double d) {
                double cd = c * d; double r = FMA(a, b, -cd); double e = FMA(-c, d, cd); return r + e; };
            ng = Vec3((float)dop(ay, bz, az, by), (float)dop(az, bx, ax, bz), (float)dop(ax, by, ay, bx));
        }
        CoordinateSystem(Normalize(ng), &dpdu, &dpdv);
    }
    Pt3 pHit = ti.b0 * p0 + ti.b1 * p1 + ti.b2 * p2;
    Vec2 uvHit(ti.b0 * uv0.x + ti.b1 * uv1.x + ti.b2 * uv2.x, ti.b0 * uv0.y + ti.b1 * uv1.y + ti.b2 * uv2.y);
    Pt3 pAbsSum = Abs(ti.b0 * p0) + Abs(ti.b1 * p1) + Abs(ti.b2 * p2);
    Vec3 pError = gamma_n(7) * V(pAbsSum);

    SurfaceIntr isect;
    isect.pi = Pt3i(pHit, pError);
    isect.uv = uvHit;
    isect.dpdus = dpdu;
    isect.n = isect.ns = Nrm3(Normalize(Cross(dp02, dp12)));
    if (a.flipNormal) isect.n = isect.ns = -isect.n;
    if (ex) { ex->dpdu = dpdu; ex->dpdv = dpdv; ex->dpdvs = dpdv; ex->dndus = ex->dndvs = Nrm3(0, 0, 0); }
    if (a.hasN) {
        Nrm3 ns = ti.b0 * a.n0 + ti.b1 * a.n1 + ti.b2 * a.n2;
        ns = LengthSquared(ns) > 0 ? Normalize(ns) : isect.n;
        Vec3 ss = dpdu;
        Vec3 ts = Cross(ns, ss);
        if (LengthSquared(ts) > 0) ss = Cross(ts, ns);
        else CoordinateSystem(V(ns), &ss, &ts);
        // SetShadingGeometry(ns, ss, ts, ..., orientationIsAuthoritative = true)
        isect.ns = ns;
        isect.n = FaceForward(isect.n, isect.ns);
        isect.dpdus = ss;
        if (ex) {
            // dndu / dndv of the interpolated normal (shapes.h:967-1000)
            ex->dpdvs = ts;
            Nrm3 dn1 = a.n0 - a.n2, dn2 = a.n1 - a.n2;
            if (degenerateUV) {
                Vec3 dn = Cross(V(a.n2 - a.n0), V(a.n1 - a.n0));
                if (LengthSquared(dn) == 0) ex->dndus = ex->dndvs = Nrm3(0, 0, 0);
                else {
                    Vec3 dnu, dnv;
                    CoordinateSystem(dn, &dnu, &dnv);
                    ex->dndus = Nrm3(dnu); ex->dndvs = Nrm3(dnv);
                }
            } else {
                float invDet = 1 / determinant;
                ex->dndus = Nrm3(DifferenceOfProducts(duv12.y, dn1.x, duv02.y, dn2.x), DifferenceOfProducts(duv12.y, dn1.y, duv02.y, dn2.y),
                                 DifferenceOfProducts(duv12.y, dn1.z, duv02.y, dn2.z)) * invDet;
                ex->dndvs = Nrm3(DifferenceOfProducts(duv02.x, dn2.x, duv12.x, dn1.x), DifferenceOfProducts(duv02.x, dn2.y, duv12.x, dn1.y),
                                 DifferenceOfProducts(duv02.x, dn2.z, duv12.x, dn1.z)) * invDet;
            }
        }
    }
    return isect;
}

// ShapeSample for triangles
struct ShapeSample { Pt3i pi; Nrm3 n; Vec2 uv; float pdf; bool valid; };

BPT_HD float TriangleArea(Pt3 p0, Pt3 p1, Pt3 p2) { return 0.5f * Length(Cross(p1 - p0, p2 - p0)); }
BPT_HD float TriangleSolidAngle(Pt3 p0, Pt3 p1, Pt3 p2, Pt3 p) {
    return SphericalTriangleArea(Normalize(p0 - p), Normalize(p1 - p), Normalize(p2 - p));
}
constexpr float kMinSphericalSampleArea = 3e-4f;
constexpr float kMaxSphericalSampleArea = 6.22f;

BPT_HD Nrm3 TriSampleNormal(Pt3 p0, Pt3 p1, Pt3 p2, const TriAttribs &a, float b0, float b1) {
    Nrm3 n = Normalize(Nrm3(Cross(p1 - p0, p2 - p0)));
    if (a.hasN) {
        Nrm3 ns(b0 * a.n0 + b1 * a.n1 + (1 - b0 - b1) * a.n2);
        n = FaceForward(n, ns);
    } else if (a.flipNormal)
        n = n * -1.f;
    return n;
}

BPT_HD ShapeSample TriangleSampleArea(Pt3 p0, Pt3 p1, Pt3 p2, const TriAttribs &a, Vec2 u) {
    float b[3];
    SampleUniformTriangle(u, b);
    ShapeSample ss;
    Pt3 p = b[0] * p0 + b[1] * p1 + b[2] * p2;
    ss.n = TriSampleNormal(p0, p1, p2, a, b[0], b[1]);
    Vec2 uv0(0, 0), uv1(1, 0), uv2(1, 1);
    if (a.hasUV) { uv0 = a.uv0; uv1 = a.uv1; uv2 = a.uv2; }
    ss.uv = Vec2(b[0] * uv0.x + b[1] * uv1.x + b[2] * uv2.x, b[0] * uv0.y + b[1] * uv1.y + b[2] * uv2.y);
    Pt3 pAbsSum = Abs(b[0] * p0) + Abs(b[1] * p1) + Abs((1 - b[0] - b[1]) * p2);
    Vec3 pError = V(gamma_n(6) * pAbsSum);
    ss.pi = Pt3i(p, pError);
    ss.pdf = 1 / TriangleArea(p0, p1, p2);
    ss.valid = true;
    return ss;
}

// Triangle::Sample(const ShapeSampleContext &ctx, Point2f u): ctxP = ctx.p(), ctxNs = ctx.ns
BPT_HD ShapeSample TriangleSampleSolidAngle(Pt3 p0, Pt3 p1, Pt3 p2, const TriAttribs &a, Pt3 ctxP, Nrm3 ctxNs, Vec2 u) {
    ShapeSample ss;
    ss.valid = false;
    float solidAngle = TriangleSolidAngle(p0, p1, p2, ctxP);
    if (solidAngle < kMinSphericalSampleArea || solidAngle > kMaxSphericalSampleArea) {
        ss = TriangleSampleArea(p0, p1, p2, a, u);
        Vec3 wi = ss.pi.p() - ctxP;
        if (LengthSquared(wi) == 0) { ss.valid = false; return ss; }
        wi = Normalize(wi);
        ss.pdf /= AbsDot(ss.n, -wi) / DistanceSquared(ctxP, ss.pi.p());
        if (IsInf_(ss.pdf)) { ss.valid = false; return ss; }
        return ss;
    }
    float pdf = 1;
    if (ctxNs != Nrm3(0, 0, 0)) {
        Pt3 rp = ctxP;
        Vec3 wi0 = Normalize(p0 - rp), wi1 = Normalize(p1 - rp), wi2 = Normalize(p2 - rp);
        float w[4] = {stdmax(0.01f, AbsDot(ctxNs, wi1)), stdmax(0.01f, AbsDot(ctxNs, wi1)),
                      stdmax(0.01f, AbsDot(ctxNs, wi0)), stdmax(0.01f, AbsDot(ctxNs, wi2))};
        u = SampleBilinear(u, w);
        pdf = BilinearPDF(u, w);
    }
    float triPDF, b[3];
    SampleSphericalTriangle(p0, p1, p2, ctxP, u, &triPDF, b);
    if (triPDF == 0) return ss;
    pdf *= triPDF;
    Pt3 pAbsSum = Abs(b[0] * p0) + Abs(b[1] * p1) + Abs((1 - b[0] - b[1]) * p2);
    Vec3 pError = V(gamma_n(6) * pAbsSum);
    Pt3 p = b[0] * p0 + b[1] * p1 + b[2] * p2;
    ss.n = TriSampleNormal(p0, p1, p2, a, b[0], b[1]);
    Vec2 uv0(0, 0), uv1(1, 0), uv2(1, 1);
    if (a.hasUV) { uv0 = a.uv0; uv1 = a.uv1; uv2 = a.uv2; }
    ss.uv = Vec2(b[0] * uv0.x + b[1] * uv1.x + b[2] * uv2.x, b[0] * uv0.y + b[1] * uv1.y + b[2] * uv2.y);
    ss.pi = Pt3i(p, pError);
    ss.pdf = pdf;
    ss.valid = true;
    return ss;
}

// Triangle::PDF(const ShapeSampleContext &ctx, Vector3f wi)
BPT_HD float TrianglePDFSolidAngle(Pt3 p0, Pt3 p1, Pt3 p2, const TriAttribs &a, Pt3i ctxPi, Nrm3 ctxN, Nrm3 ctxNs, Vec3 wi) {
    Pt3 ctxP = ctxPi.p();
    float solidAngle = TriangleSolidAngle(p0, p1, p2, ctxP);
    if (solidAngle < kMinSphericalSampleArea || solidAngle > kMaxSphericalSampleArea) {
        // Ray ray = ctx.SpawnRay(wi); isect = Intersect(ray)  (tMax = Infinity)
        Pt3 o = OffsetRayOrigin(ctxPi, ctxN, wi);
        TriHit th;
        if (!IntersectTriangle(o, wi, BPT_INFINITY, p0, p1, p2, &th)) return 0;
        SurfaceIntr isect = TriangleInteraction(p0, p1, p2, a, th);
        float pdf = (1 / TriangleArea(p0, p1, p2)) / (AbsDot(isect.n, -wi) / DistanceSquared(ctxP, isect.p()));
        if (IsInf_(pdf)) pdf = 0;
        return pdf;
    }
    float pdf = 1 / solidAngle;
    if (ctxNs != Nrm3(0, 0, 0)) {
        Vec2 u = InvertSphericalTriangleSample(p0, p1, p2, ctxP, wi);
        Pt3 rp = ctxP;
        Vec3 wi0 = Normalize(p0 - rp), wi1 = Normalize(p1 - rp), wi2 = Normalize(p2 - rp);
        float w[4] = {stdmax(0.01f, AbsDot(ctxNs, wi1)), stdmax(0.01f, AbsDot(ctxNs, wi1)),
                      stdmax(0.01f, AbsDot(ctxNs, wi0)), stdmax(0.01f, AbsDot(ctxNs, wi2))};
        pdf *= BilinearPDF(u, w);
    }
    return pdf;
}

}  // namespace bpt
