// quadrics.h -- Sphere, Disk, Cylinder: intersection, surface interaction, area / solid-angle sampling (host + device).
// Operation-for-operation restatements, interval arithmetic included:
//   Interval * / Sqr Sqrt                         util/math.h:818-1075
//   Transform::operator()(Point3fi / Vector3fi)   util/transform.h:133-177, 272-305
//   Sphere    BasicIntersect / InteractionFromIntersection / Sample / PDF    shapes.h:146-400, shapes.cpp:33-58
//   Disk      ...                                                           shapes.h:445-565, shapes.cpp:84-96
//   Cylinder  ...                                                           shapes.h:606-790, shapes.cpp:121-124
// The reference's GPU path intersects these as custom primitives (gpu/optix/optix.cu:301-377, aggregate.cpp:848-1025); here a
// quadric is one leaf record of the 8-wide BVH (its bounding box in place of triangle vertices, class bit kClsShape) and the
// traversal calls IntersectQuadricLeaf (outlined) for it.  A hit is reported as the object-space hit point (b0, b1, b2) = pObj;
// phi is a function of pObj (atan2), so the shading kernels rebuild the QuadricIntersection from the three floats.
#pragma once
#include "pmath.h"
#include "sampling.h"
#include "scene.h"
#include "shapes.h"
#include "camera.h"   // XformPoint, XformVector

namespace bpt {

// ---------------------------------------------------------------- Interval arithmetic beyond + and -
BPT_HD float min4(float a, float b, float c, float d) { return stdmin(stdmin(stdmin(a, b), c), d); }   // std::min({a, b, c, d})
BPT_HD float max4(float a, float b, float c, float d) { return stdmax(stdmax(stdmax(a, b), c), d); }
BPT_HD bool InRange(float v, Interval i) { return v >= i.low && v <= i.high; }
BPT_HD Interval operator*(Interval a, Interval i) {
    return Interval(min4(MulRoundDown(a.low, i.low), MulRoundDown(a.high, i.low), MulRoundDown(a.low, i.high), MulRoundDown(a.high, i.high)),
                    max4(MulRoundUp(a.low, i.low), MulRoundUp(a.high, i.low), MulRoundUp(a.low, i.high), MulRoundUp(a.high, i.high)));
}
BPT_HD Interval operator/(Interval a, Interval i) {
    if (InRange(0, i)) return Interval(-BPT_INFINITY, BPT_INFINITY);
    return Interval(min4(DivRoundDown(a.low, i.low), DivRoundDown(a.high, i.low), DivRoundDown(a.low, i.high), DivRoundDown(a.high, i.high)),
                    max4(DivRoundUp(a.low, i.low), DivRoundUp(a.high, i.low), DivRoundUp(a.low, i.high), DivRoundUp(a.high, i.high)));
}
BPT_HD Interval operator*(float f, Interval i) {
    if (f > 0) return Interval(MulRoundDown(f, i.low), MulRoundUp(f, i.high));
    return Interval(MulRoundDown(f, i.high), MulRoundUp(f, i.low));
}
BPT_HD Interval SqrI(Interval i) {
    float alow = fabsf(i.low), ahigh = fabsf(i.high);
    if (alow > ahigh) { float t = alow; alow = ahigh; ahigh = t; }
    if (InRange(0, i)) return Interval(0, MulRoundUp(ahigh, ahigh));
    return Interval(MulRoundDown(alow, alow), MulRoundUp(ahigh, ahigh));
}
BPT_HD Interval SqrtI(Interval i) { return Interval(stdmax(0.f, NextFloatDown(sqrtf(i.low))), NextFloatUp(sqrtf(i.high))); }
BPT_HD bool SameInterval(Interval a, Interval b) { return a.low == b.low && a.high == b.high; }

struct Vec3i { Interval x, y, z; };
// Transform::operator()(const Vector3fi &) for an exact input vector
BPT_HD Vec3i XformVec3iExact(const Mat4 &T, Vec3 v) {
    const float (*m)[4] = T.m;
    float x = v.x, y = v.y, z = v.z;
    Vec3 e;
    e.x = gamma_n(3) * (fabsf(m[0][0] * x) + fabsf(m[0][1] * y) + fabsf(m[0][2] * z));
    e.y = gamma_n(3) * (fabsf(m[1][0] * x) + fabsf(m[1][1] * y) + fabsf(m[1][2] * z));
    e.z = gamma_n(3) * (fabsf(m[2][0] * x) + fabsf(m[2][1] * y) + fabsf(m[2][2] * z));
    float xp = m[0][0] * x + m[0][1] * y + m[0][2] * z;
    float yp = m[1][0] * x + m[1][1] * y + m[1][2] * z;
    float zp = m[2][0] * x + m[2][1] * y + m[2][2] * z;
    Vec3i r;
    r.x = Interval::FromValueAndError(xp, e.x); r.y = Interval::FromValueAndError(yp, e.y); r.z = Interval::FromValueAndError(zp, e.z);
    return r;
}
// Transform::operator()(const Point3fi &) (util/transform.h:133-177)
BPT_HD Pt3i XformPoint3i(const Mat4 &T, const Pt3i &p) {
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
BPT_HD Nrm3 XformNormal(const Mat4 &Tinv, Nrm3 n) {   // Transform::operator()(Normal3f): transpose of the inverse
    const float (*mi)[4] = Tinv.m;
    float x = n.x, y = n.y, z = n.z;
    return Nrm3(mi[0][0] * x + mi[1][0] * y + mi[2][0] * z, mi[0][1] * x + mi[1][1] * y + mi[2][1] * z,
                mi[0][2] * x + mi[1][2] * y + mi[2][2] * z);
}
// ---------------------------------------------------------------- the shapes
enum ShapeType { SHAPE_SPHERE = 0, SHAPE_DISK = 1, SHAPE_CYLINDER = 2 };

struct DeviceShape {
    int type;
    int reverseOrientation, swapsHandedness, pad;
    float radius, zMin, zMax, thetaZMin, thetaZMax, phiMax;   // sphere / cylinder (constructor-clamped values)
    float height, innerRadius;                                // disk
    Mat4 renderFromObject, objectFromRender;                  // "render" = the space of the enclosing BVH (instance space inside an object)
};

// Sphere / Disk / Cylinder constructors (shapes.h:118-128, 407-417, 801-812) from the parameters of the .pbrt file, and the object-space
// box the shape's Bounds() transforms (shapes.cpp:33-36, 85-88, 121-124).  Host only (scene preparation).
inline DeviceShape MakeQuadric(int type, float radius, float zmin, float zmax, float phimaxDeg, float height, float innerRadius, bool reverseOrientation,
                               bool swapsHandedness, const Mat4 &renderFromObject, const Mat4 &objectFromRender, Pt3 *objMin, Pt3 *objMax) {
    DeviceShape s{};
    s.type = type;
    s.reverseOrientation = reverseOrientation; s.swapsHandedness = swapsHandedness;
    s.radius = radius;
    s.phiMax = Radians(Clampf(phimaxDeg, 0, 360));
    if (type == SHAPE_SPHERE) {
        s.zMin = Clampf(stdmin(zmin, zmax), -radius, radius);
        s.zMax = Clampf(stdmax(zmin, zmax), -radius, radius);
        s.thetaZMin = acosf(Clampf(stdmin(zmin, zmax) / radius, -1, 1));
        s.thetaZMax = acosf(Clampf(stdmax(zmin, zmax) / radius, -1, 1));
        *objMin = Pt3(-radius, -radius, s.zMin); *objMax = Pt3(radius, radius, s.zMax);
    } else if (type == SHAPE_DISK) {
        s.height = height; s.innerRadius = innerRadius;
        *objMin = Pt3(-radius, -radius, height); *objMax = Pt3(radius, radius, height);
    } else {
        s.zMin = stdmin(zmin, zmax); s.zMax = stdmax(zmin, zmax);
        *objMin = Pt3(-radius, -radius, s.zMin); *objMax = Pt3(radius, radius, s.zMax);
    }
    s.renderFromObject = renderFromObject; s.objectFromRender = objectFromRender;
    return s;
}

struct QuadricHit { float t; Pt3 pObj; float phi; };

BPT_HD float QuadricPhi(Pt3 pObj) {
    float phi = ATan2F(pObj.y, pObj.x);
    if (phi < 0) phi += 2 * kPi;
    return phi;
}

#if defined(__CUDA_ARCH__)
#define BPT_SHAPE_FN __device__ __noinline__
#else
#define BPT_SHAPE_FN inline
#endif

BPT_SHAPE_FN bool SphereBasicIntersect(const DeviceShape &s, Pt3 ro, Vec3 rd, float tMax, QuadricHit *hit) {
    const float radius = s.radius;
    Pt3i oi = XformPoint3i(s.objectFromRender, Pt3i(ro));
    Vec3i di = XformVec3iExact(s.objectFromRender, rd);
    Interval a = SqrI(di.x) + SqrI(di.y) + SqrI(di.z);
    Interval b = 2.f * (di.x * oi.x + di.y * oi.y + di.z * oi.z);
    Interval c = SqrI(oi.x) + SqrI(oi.y) + SqrI(oi.z) - SqrI(Interval(radius));
    Interval f = b / (2.f * a);
    Interval vx = oi.x - di.x * f, vy = oi.y - di.y * f, vz = oi.z - di.z * f;
    Interval length = SqrtI(SqrI(vx) + SqrI(vy) + SqrI(vz));
    Interval discrim = 4.f * a * (Interval(radius) + length) * (Interval(radius) - length);
    if (discrim.low < 0) return false;
    Interval rootDiscrim = SqrtI(discrim);
    Interval q;
    if (b.Midpoint() < 0) q = -.5f * (b - rootDiscrim);
    else q = -.5f * (b + rootDiscrim);
    Interval t0 = q / a, t1 = c / q;
    if (t0.low > t1.low) { Interval t = t0; t0 = t1; t1 = t; }
    if (t0.high > tMax || t1.low <= 0) return false;
    Interval tShapeHit = t0;
    if (tShapeHit.low <= 0) {
        tShapeHit = t1;
        if (tShapeHit.high > tMax) return false;
    }
    const Pt3 o = oi.p();
    const Vec3 d(di.x.Midpoint(), di.y.Midpoint(), di.z.Midpoint());
    Pt3 pHit = o + tShapeHit.Midpoint() * d;
    pHit = pHit * (radius / Distance(pHit, Pt3(0, 0, 0)));
    if (pHit.x == 0 && pHit.y == 0) pHit.x = 1e-5f * radius;
    float phi = QuadricPhi(pHit);
    if ((s.zMin > -radius && pHit.z < s.zMin) || (s.zMax < radius && pHit.z > s.zMax) || phi > s.phiMax) {
        if (SameInterval(tShapeHit, t1)) return false;
        if (t1.high > tMax) return false;
        tShapeHit = t1;
        pHit = o + tShapeHit.Midpoint() * d;
        pHit = pHit * (radius / Distance(pHit, Pt3(0, 0, 0)));
        if (pHit.x == 0 && pHit.y == 0) pHit.x = 1e-5f * radius;
        phi = QuadricPhi(pHit);
        if ((s.zMin > -radius && pHit.z < s.zMin) || (s.zMax < radius && pHit.z > s.zMax) || phi > s.phiMax) return false;
    }
    hit->t = tShapeHit.Midpoint(); hit->pObj = pHit; hit->phi = phi;
    return true;
}

BPT_SHAPE_FN bool DiskBasicIntersect(const DeviceShape &s, Pt3 ro, Vec3 rd, float tMax, QuadricHit *hit) {
    Pt3i oi = XformPoint3i(s.objectFromRender, Pt3i(ro));
    Vec3i di = XformVec3iExact(s.objectFromRender, rd);
    if (di.z.Midpoint() == 0) return false;
    float tShapeHit = (s.height - oi.z.Midpoint()) / di.z.Midpoint();
    if (tShapeHit <= 0 || tShapeHit >= tMax) return false;
    Pt3 pHit = oi.p() + tShapeHit * Vec3(di.x.Midpoint(), di.y.Midpoint(), di.z.Midpoint());
    float dist2 = Sqr(pHit.x) + Sqr(pHit.y);
    if (dist2 > Sqr(s.radius) || dist2 < Sqr(s.innerRadius)) return false;
    float phi = QuadricPhi(pHit);
    if (phi > s.phiMax) return false;
    hit->t = tShapeHit; hit->pObj = pHit; hit->phi = phi;
    return true;
}

BPT_SHAPE_FN bool CylinderBasicIntersect(const DeviceShape &s, Pt3 ro, Vec3 rd, float tMax, QuadricHit *hit) {
    const float radius = s.radius;
    Pt3i oi = XformPoint3i(s.objectFromRender, Pt3i(ro));
    Vec3i di = XformVec3iExact(s.objectFromRender, rd);
    Interval a = SqrI(di.x) + SqrI(di.y);
    Interval b = 2.f * (di.x * oi.x + di.y * oi.y);
    Interval c = SqrI(oi.x) + SqrI(oi.y) - SqrI(Interval(radius));
    Interval f = b / (2.f * a);
    Interval vx = oi.x - f * di.x, vy = oi.y - f * di.y;
    Interval length = SqrtI(SqrI(vx) + SqrI(vy));
    Interval discrim = 4.f * a * (Interval(radius) + length) * (Interval(radius) - length);
    if (discrim.low < 0) return false;
    Interval rootDiscrim = SqrtI(discrim);
    Interval q;
    if (b.Midpoint() < 0) q = -.5f * (b - rootDiscrim);
    else q = -.5f * (b + rootDiscrim);
    Interval t0 = q / a, t1 = c / q;
    if (t0.low > t1.low) { Interval t = t0; t0 = t1; t1 = t; }
    if (t0.high > tMax || t1.low <= 0) return false;
    Interval tShapeHit = t0;
    if (tShapeHit.low <= 0) {
        tShapeHit = t1;
        if (tShapeHit.high > tMax) return false;
    }
    const Pt3 o = oi.p();
    const Vec3 d(di.x.Midpoint(), di.y.Midpoint(), di.z.Midpoint());
    Pt3 pHit = o + tShapeHit.Midpoint() * d;
    float hitRad = sqrtf(Sqr(pHit.x) + Sqr(pHit.y));
    pHit.x *= radius / hitRad; pHit.y *= radius / hitRad;
    float phi = QuadricPhi(pHit);
    if (pHit.z < s.zMin || pHit.z > s.zMax || phi > s.phiMax) {
        if (SameInterval(tShapeHit, t1)) return false;
        tShapeHit = t1;
        if (t1.high > tMax) return false;
        pHit = o + tShapeHit.Midpoint() * d;
        hitRad = sqrtf(Sqr(pHit.x) + Sqr(pHit.y));
        pHit.x *= radius / hitRad; pHit.y *= radius / hitRad;
        phi = QuadricPhi(pHit);
        if (pHit.z < s.zMin || pHit.z > s.zMax || phi > s.phiMax) return false;
    }
    hit->t = tShapeHit.Midpoint(); hit->pObj = pHit; hit->phi = phi;
    return true;
}

BPT_HD bool QuadricBasicIntersect(const DeviceShape &s, Pt3 o, Vec3 d, float tMax, QuadricHit *hit) {
    if (s.type == SHAPE_SPHERE) return SphereBasicIntersect(s, o, d, tMax, hit);
    if (s.type == SHAPE_DISK) return DiskBasicIntersect(s, o, d, tMax, hit);
    return CylinderBasicIntersect(s, o, d, tMax, hit);
}

// (*renderFromObject)(SurfaceInteraction(pi, uv, woObject, dpdu, dpdv, dndu, dndv, time, flipNormal)) for the members the path reads
// (interaction.h:164-183, util/transform.cpp:227-260)
BPT_HD SurfaceIntr FinishQuadricInteraction(const DeviceShape &s, Pt3 pHit, Vec3 pError, Vec2 uv, Vec3 dpdu, Vec3 dpdv, Nrm3 dndu, Nrm3 dndv,
                                            SurfaceExtras *ex) {
    Nrm3 n(Normalize(Cross(dpdu, dpdv)));
    if ((s.reverseOrientation != 0) ^ (s.swapsHandedness != 0)) n = -n;
    SurfaceIntr r;
    r.pi = XformPoint3i(s.renderFromObject, Pt3i(pHit, pError));
    r.n = Normalize(XformNormal(s.objectFromRender, n));
    r.uv = uv;
    r.ns = Normalize(XformNormal(s.objectFromRender, n));
    r.dpdus = XformVector(s.renderFromObject, dpdu);
    r.ns = FaceForward(r.ns, r.n);
    if (ex) {
        ex->dpdu = XformVector(s.renderFromObject, dpdu); ex->dpdv = XformVector(s.renderFromObject, dpdv);
        ex->dpdvs = ex->dpdv;
        ex->dndus = XformNormal(s.objectFromRender, dndu); ex->dndvs = XformNormal(s.objectFromRender, dndv);
    }
    return r;
}

BPT_HD void FundamentalForms(Vec3 dpdu, Vec3 dpdv, Vec3 d2Pduu, Vec3 d2Pduv, Vec3 d2Pdvv, Nrm3 *dndu, Nrm3 *dndv) {
    float E = Dot(dpdu, dpdu), F = Dot(dpdu, dpdv), G = Dot(dpdv, dpdv);
    Vec3 n = Normalize(Cross(dpdu, dpdv));
    float e = Dot(n, d2Pduu), f = Dot(n, d2Pduv), g = Dot(n, d2Pdvv);
    float EGF2 = DifferenceOfProducts(E, G, F, F);
    float invEGF2 = (EGF2 == 0) ? 0.f : 1 / EGF2;
    *dndu = Nrm3((f * F - e * G) * invEGF2 * dpdu + (e * F - f * E) * invEGF2 * dpdv);
    *dndv = Nrm3((g * F - f * G) * invEGF2 * dpdu + (f * F - g * E) * invEGF2 * dpdv);
}

// Shape::InteractionFromIntersection for the three quadrics; pHit = QuadricIntersection::pObj, phi recomputed from it
BPT_SHAPE_FN SurfaceIntr QuadricInteraction(const DeviceShape &s, Pt3 pHit, SurfaceExtras *ex) {
    const float phi = QuadricPhi(pHit);
    const float phiMax = s.phiMax;
    if (s.type == SHAPE_SPHERE) {
        const float radius = s.radius;
        float u = phi / phiMax;
        float cosTheta = pHit.z / radius;
        float theta = SafeACos(cosTheta);
        float v = (theta - s.thetaZMin) / (s.thetaZMax - s.thetaZMin);
        float zRadius = sqrtf(Sqr(pHit.x) + Sqr(pHit.y));
        float cosPhi = pHit.x / zRadius, sinPhi = pHit.y / zRadius;
        Vec3 dpdu(-phiMax * pHit.y, phiMax * pHit.x, 0);
        float sinTheta = SafeSqrt(1 - Sqr(cosTheta));
        Vec3 dpdv = (s.thetaZMax - s.thetaZMin) * Vec3(pHit.z * cosPhi, pHit.z * sinPhi, -radius * sinTheta);
        Vec3 d2Pduu = -phiMax * phiMax * Vec3(pHit.x, pHit.y, 0);
        Vec3 d2Pduv = (s.thetaZMax - s.thetaZMin) * pHit.z * phiMax * Vec3(-sinPhi, cosPhi, 0.f);
        Vec3 d2Pdvv = -Sqr(s.thetaZMax - s.thetaZMin) * Vec3(pHit.x, pHit.y, pHit.z);
        Nrm3 dndu, dndv;
        FundamentalForms(dpdu, dpdv, d2Pduu, d2Pduv, d2Pdvv, &dndu, &dndv);
        Vec3 pError = gamma_n(5) * Abs(Vec3(pHit.x, pHit.y, pHit.z));
        return FinishQuadricInteraction(s, pHit, pError, Vec2(u, v), dpdu, dpdv, dndu, dndv, ex);
    }
    if (s.type == SHAPE_DISK) {
        float u = phi / phiMax;
        float rHit = sqrtf(Sqr(pHit.x) + Sqr(pHit.y));
        float v = (s.radius - rHit) / (s.radius - s.innerRadius);
        Vec3 dpdu(-phiMax * pHit.y, phiMax * pHit.x, 0);
        Vec3 dpdv = Vec3(pHit.x, pHit.y, 0) * (s.innerRadius - s.radius) / rHit;
        pHit.z = s.height;
        return FinishQuadricInteraction(s, pHit, Vec3(0, 0, 0), Vec2(u, v), dpdu, dpdv, Nrm3(0, 0, 0), Nrm3(0, 0, 0), ex);
    }
    float u = phi / phiMax;
    float v = (pHit.z - s.zMin) / (s.zMax - s.zMin);
    Vec3 dpdu(-phiMax * pHit.y, phiMax * pHit.x, 0);
    Vec3 dpdv(0, 0, s.zMax - s.zMin);
    Vec3 d2Pduu = -phiMax * phiMax * Vec3(pHit.x, pHit.y, 0);
    Nrm3 dndu, dndv;
    FundamentalForms(dpdu, dpdv, d2Pduu, Vec3(0, 0, 0), Vec3(0, 0, 0), &dndu, &dndv);
    Vec3 pError = gamma_n(3) * Abs(Vec3(pHit.x, pHit.y, 0));
    return FinishQuadricInteraction(s, pHit, pError, Vec2(u, v), dpdu, dpdv, dndu, dndv, ex);
}

// wo of the interaction a quadric returns for a ray of direction d (InteractionFromIntersection(isect, -ray.d, ...)): the
// Interaction constructor normalises woObject, Transform::operator()(SurfaceInteraction) normalises the transformed vector again
BPT_HD Vec3 QuadricWo(const DeviceShape &s, Vec3 rayD) {
    Vec3 woObject = XformVector(s.objectFromRender, -rayD);
    return Normalize(XformVector(s.renderFromObject, Normalize(woObject)));
}

BPT_HD float QuadricArea(const DeviceShape &s) {
    if (s.type == SHAPE_SPHERE) return s.phiMax * s.radius * (s.zMax - s.zMin);
    if (s.type == SHAPE_DISK) return s.phiMax * 0.5f * (Sqr(s.radius) - Sqr(s.innerRadius));
    return (s.zMax - s.zMin) * s.radius * s.phiMax;
}

BPT_HD Vec3 SampleUniformSphere(Vec2 u) {
    float z = 1 - 2 * u.x;
    float r = SafeSqrt(1 - Sqr(z));
    float phi = 2 * kPi * u.y;
    return Vec3(r * CosF(phi), r * SinF(phi), z);
}
BPT_HD Vec3 SphericalDirection(float sinTheta, float cosTheta, float phi) {
    return Vec3(Clampf(sinTheta, -1, 1) * CosF(phi), Clampf(sinTheta, -1, 1) * SinF(phi), Clampf(cosTheta, -1, 1));
}

// Shape::Sample(Point2f u): a point by area
BPT_HD ShapeSample QuadricSampleArea(const DeviceShape &s, Vec2 u) {
    ShapeSample ss;
    ss.valid = true;
    if (s.type == SHAPE_SPHERE) {
        Vec3 w = s.radius * SampleUniformSphere(u);
        Pt3 pObj = Pt3(0, 0, 0) + w;
        pObj = pObj * (s.radius / Distance(pObj, Pt3(0, 0, 0)));
        Vec3 pObjError = gamma_n(5) * Abs(Vec3(pObj.x, pObj.y, pObj.z));
        Nrm3 n = Normalize(XformNormal(s.objectFromRender, Nrm3(pObj.x, pObj.y, pObj.z)));
        if (s.reverseOrientation) n = n * -1.f;
        float theta = SafeACos(pObj.z / s.radius);
        float phi = QuadricPhi(pObj);
        ss.uv = Vec2(phi / s.phiMax, (theta - s.thetaZMin) / (s.thetaZMax - s.thetaZMin));
        ss.pi = XformPoint3i(s.renderFromObject, Pt3i(pObj, pObjError));
        ss.n = n;
    } else if (s.type == SHAPE_DISK) {
        Vec2 pd = SampleUniformDiskConcentric(u);
        Pt3 pObj(pd.x * s.radius, pd.y * s.radius, s.height);
        ss.pi = XformPoint3i(s.renderFromObject, Pt3i(pObj));
        Nrm3 n = Normalize(XformNormal(s.objectFromRender, Nrm3(0, 0, 1)));
        if (s.reverseOrientation) n = n * -1.f;
        float phi = ATan2F(pd.y, pd.x);
        if (phi < 0) phi += 2 * kPi;
        float radiusSample = sqrtf(Sqr(pObj.x) + Sqr(pObj.y));
        ss.uv = Vec2(phi / s.phiMax, (s.radius - radiusSample) / (s.radius - s.innerRadius));
        ss.n = n;
    } else {
        float z = Lerp(u.x, s.zMin, s.zMax);
        float phi = u.y * s.phiMax;
        Pt3 pObj(s.radius * CosF(phi), s.radius * SinF(phi), z);
        float hitRad = sqrtf(Sqr(pObj.x) + Sqr(pObj.y));
        pObj.x *= s.radius / hitRad; pObj.y *= s.radius / hitRad;
        Vec3 pObjError = gamma_n(3) * Abs(Vec3(pObj.x, pObj.y, 0));
        ss.pi = XformPoint3i(s.renderFromObject, Pt3i(pObj, pObjError));
        Nrm3 n = Normalize(XformNormal(s.objectFromRender, Nrm3(pObj.x, pObj.y, 0)));
        if (s.reverseOrientation) n = n * -1.f;
        ss.uv = Vec2(phi / s.phiMax, (pObj.z - s.zMin) / (s.zMax - s.zMin));
        ss.n = n;
    }
    ss.pdf = 1 / QuadricArea(s);
    return ss;
}

// Shape::Sample(const ShapeSampleContext &ctx, Point2f u); ctx = (pi, n, ns) of the reference point
BPT_SHAPE_FN ShapeSample QuadricSampleSolidAngle(const DeviceShape &s, Pt3i ctxPi, Nrm3 ctxN, Vec2 u) {
    const Pt3 ctxP = ctxPi.p();
    if (s.type == SHAPE_SPHERE) {
        Pt3 pCenter = XformPoint(s.renderFromObject, Pt3(0, 0, 0));
        // ctx.OffsetRayOrigin(pCenter) = OffsetRayOrigin(pi, n, pCenter - p())
        Pt3 pOrigin = OffsetRayOrigin(ctxPi, ctxN, pCenter - ctxP);
        if (!(DistanceSquared(pOrigin, pCenter) <= Sqr(s.radius))) {
            ShapeSample ss;
            float sinThetaMax = s.radius / Distance(ctxP, pCenter);
            float sin2ThetaMax = Sqr(sinThetaMax);
            float cosThetaMax = SafeSqrt(1 - sin2ThetaMax);
            float oneMinusCosThetaMax = 1 - cosThetaMax;
            float cosTheta = (cosThetaMax - 1) * u.x + 1;
            float sin2Theta = 1 - Sqr(cosTheta);
            if (sin2ThetaMax < 0.00068523f) {
                sin2Theta = sin2ThetaMax * u.x;
                cosTheta = sqrtf(1 - sin2Theta);
                oneMinusCosThetaMax = sin2ThetaMax / 2;
            }
            float cosAlpha = sin2Theta / sinThetaMax + cosTheta * SafeSqrt(1 - sin2Theta / Sqr(sinThetaMax));
            float sinAlpha = SafeSqrt(1 - Sqr(cosAlpha));
            float phi = u.y * 2 * kPi;
            Vec3 w = SphericalDirection(sinAlpha, cosAlpha, phi);
            Frame samplingFrame = Frame::FromZ(Normalize(pCenter - ctxP));
            Nrm3 n(samplingFrame.FromLocal(-w));
            Pt3 p = pCenter + s.radius * Vec3(n.x, n.y, n.z);
            if (s.reverseOrientation) n = n * -1.f;
            Vec3 pError = gamma_n(5) * Abs(Vec3(p.x, p.y, p.z));
            Pt3 pObj = XformPoint(s.objectFromRender, p);
            float theta = SafeACos(pObj.z / s.radius);
            float spherePhi = QuadricPhi(pObj);
            ss.uv = Vec2(spherePhi / s.phiMax, (theta - s.thetaZMin) / (s.thetaZMax - s.thetaZMin));
            ss.pi = Pt3i(p, pError);
            ss.n = n;
            ss.pdf = 1 / (2 * kPi * oneMinusCosThetaMax);
            ss.valid = true;
            return ss;
        }
    }
    // sample by area and convert to solid angle
    ShapeSample ss = QuadricSampleArea(s, u);
    Vec3 wi = ss.pi.p() - ctxP;
    if (LengthSquared(wi) == 0) { ss.valid = false; return ss; }
    wi = Normalize(wi);
    ss.pdf /= AbsDot(ss.n, -wi) / DistanceSquared(ctxP, ss.pi.p());
    if (IsInf_(ss.pdf)) ss.valid = false;
    return ss;
}

// Shape::PDF(const ShapeSampleContext &ctx, Vector3f wi)
BPT_SHAPE_FN float QuadricPDFSolidAngle(const DeviceShape &s, Pt3i ctxPi, Nrm3 ctxN, Vec3 wi) {
    const Pt3 ctxP = ctxPi.p();
    if (s.type == SHAPE_SPHERE) {
        Pt3 pCenter = XformPoint(s.renderFromObject, Pt3(0, 0, 0));
        Pt3 pOrigin = OffsetRayOrigin(ctxPi, ctxN, pCenter - ctxP);
        if (!(DistanceSquared(pOrigin, pCenter) <= Sqr(s.radius))) {
            float sin2ThetaMax = s.radius * s.radius / DistanceSquared(ctxP, pCenter);
            float cosThetaMax = SafeSqrt(1 - sin2ThetaMax);
            float oneMinusCosThetaMax = 1 - cosThetaMax;
            if (sin2ThetaMax < 0.00068523f) oneMinusCosThetaMax = sin2ThetaMax / 2;
            return 1 / (2 * kPi * oneMinusCosThetaMax);
        }
    }
    // Ray ray = ctx.SpawnRay(wi); isect = Intersect(ray)  (tMax = Infinity)
    Pt3 o = OffsetRayOrigin(ctxPi, ctxN, wi);
    QuadricHit qh;
    if (!QuadricBasicIntersect(s, o, wi, BPT_INFINITY, &qh)) return 0;
    SurfaceIntr isect = QuadricInteraction(s, qh.pObj, nullptr);
    float pdf = (1 / QuadricArea(s)) / (AbsDot(isect.n, -wi) / DistanceSquared(ctxP, isect.p()));
    if (IsInf_(pdf)) pdf = 0;
    return pdf;
}

}  // namespace bpt
