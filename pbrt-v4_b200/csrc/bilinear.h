// bilinear.h -- BilinearPatch: intersection, surface interaction, area and solid-angle sampling (host + device).
// Operation-for-operation restatements of
//   IntersectBilinearPatch                         shapes.h:1279-1347   (Quadratic util/math.h:613-636, Determinant :1419-1425)
//   BilinearPatch::InteractionFromIntersection     shapes.h:1396-1508   (RotateFromTo util/transform.h:249-270)
//   BilinearPatch ctor (area), IsRectangle         shapes.cpp:1038-1069, shapes.h:1521-1543
//   Bounds / NormalBounds / Sample / PDF           shapes.cpp:1071-1357
//   SampleSphericalRectangle / Invert...           util/sampling.cpp:163-346,  SphericalQuadArea util/vecmath.h:1646-1664,
//   InvertBilinear                                 util/vecmath.h:623-652
// In the BVH a patch is one leaf record holding its bounds (class bit kClsShape, like the quadrics); its four vertices come from
// the patch pool.  A hit is recorded as (u, v) in (b0, b1).
#pragma once
#include "pmath.h"
#include "sampling.h"
#include "shapes.h"

namespace bpt {

struct PatchData {
    Pt3 p00, p10, p01, p11;
    bool hasN, hasUV, flipNormal;     // flipNormal = reverseOrientation ^ transformSwapsHandedness
    Nrm3 n00, n10, n01, n11;
    Vec2 uv00, uv10, uv01, uv11;
    bool isRectangle;
    float area;
};

BPT_HD Pt3 LerpP(float t, Pt3 a, Pt3 b) { return (1 - t) * a + t * b; }
BPT_HD Nrm3 LerpN(float t, Nrm3 a, Nrm3 b) { return (1 - t) * a + t * b; }
BPT_HD Vec2 LerpUV(float t, Vec2 a, Vec2 b) { return Vec2((1 - t) * a.x + t * b.x, (1 - t) * a.y + t * b.y); }

BPT_HD bool QuadraticF(float a, float b, float c, float *t0, float *t1) {
    if (a == 0) {
        if (b == 0) return false;
        *t0 = *t1 = -c / b;
        return true;
    }
    float discrim = DifferenceOfProducts(b, b, 4 * a, c);
    if (discrim < 0) return false;
    float rootDiscrim = sqrtf(discrim);
    float q = -0.5f * (b + copysignf(rootDiscrim, b));
    *t0 = q / a;
    *t1 = c / q;
    if (*t0 > *t1) { float t = *t0; *t0 = *t1; *t1 = t; }
    return true;
}
// Determinant(SquareMatrix<3>(m00 ... m22))
BPT_HD float Determinant3(float m00, float m01, float m02, float m10, float m11, float m12, float m20, float m21, float m22) {
    float minor12 = DifferenceOfProducts(m11, m22, m12, m21);
    float minor02 = DifferenceOfProducts(m10, m22, m12, m20);
    float minor01 = DifferenceOfProducts(m10, m21, m11, m20);
    return FMA(m02, minor01, DifferenceOfProducts(m00, minor12, m01, minor02));
}

#if defined(__CUDA_ARCH__)
#define BPT_PATCH_FN __device__ __noinline__
#else
#define BPT_PATCH_FN inline
#endif

BPT_PATCH_FN bool IntersectBilinearPatch(Pt3 ro, Vec3 rd, float tMax, Pt3 p00, Pt3 p10, Pt3 p01, Pt3 p11, float *uOut, float *vOut, float *tOut) {
    float a = Dot(Cross(p10 - p00, p01 - p11), rd);
    float c = Dot(Cross(p00 - ro, rd), p01 - p00);
    float b = Dot(Cross(p10 - ro, rd), p11 - p10) - (a + c);
    float u1, u2;
    if (!QuadraticF(a, b, c, &u1, &u2)) return false;
    float eps = gamma_n(30) * (MaxComponentValue(Abs(Vec3(ro.x, ro.y, ro.z))) + MaxComponentValue(Abs(rd)) + MaxComponentValue(Abs(Vec3(p00.x, p00.y, p00.z))) +
                               MaxComponentValue(Abs(Vec3(p10.x, p10.y, p10.z))) + MaxComponentValue(Abs(Vec3(p01.x, p01.y, p01.z))) +
                               MaxComponentValue(Abs(Vec3(p11.x, p11.y, p11.z))));
    float t = tMax, u = 0, v = 0;
    if (0 <= u1 && u1 <= 1) {
        Pt3 uo = LerpP(u1, p00, p10);
        Vec3 ud = LerpP(u1, p01, p11) - uo;
        Vec3 deltao = uo - ro;
        Vec3 perp = Cross(rd, ud);
        float p2 = LengthSquared(perp);
        float v1 = Determinant3(deltao.x, rd.x, perp.x, deltao.y, rd.y, perp.y, deltao.z, rd.z, perp.z);
        float t1 = Determinant3(deltao.x, ud.x, perp.x, deltao.y, ud.y, perp.y, deltao.z, ud.z, perp.z);
        if (t1 > p2 * eps && 0 <= v1 && v1 <= p2) {
            u = u1;
            v = v1 / p2;
            t = t1 / p2;
        }
    }
    if (0 <= u2 && u2 <= 1 && u2 != u1) {
        Pt3 uo = LerpP(u2, p00, p10);
        Vec3 ud = LerpP(u2, p01, p11) - uo;
        Vec3 deltao = uo - ro;
        Vec3 perp = Cross(rd, ud);
        float p2 = LengthSquared(perp);
        float v2 = Determinant3(deltao.x, rd.x, perp.x, deltao.y, rd.y, perp.y, deltao.z, rd.z, perp.z);
        float t2 = Determinant3(deltao.x, ud.x, perp.x, deltao.y, ud.y, perp.y, deltao.z, ud.z, perp.z);
        t2 /= p2;
        if (0 <= v2 && v2 <= p2 && t > t2 && t2 > eps) {
            t = t2;
            u = u2;
            v = v2 / p2;
        }
    }
    if (t >= tMax) return false;
    *uOut = u; *vOut = v; *tOut = t;
    return true;
}

// Rotation taking `from` onto `to` (both unit): RotateFromTo applied to a vector (util/transform.h:249-270, then Transform::operator()(Vector3f))
BPT_HD void RotateFromToMatrix(Vec3 from, Vec3 to, float r[3][3]) {
    Vec3 refl;
    if (fabsf(from.x) < 0.72f && fabsf(to.x) < 0.72f) refl = Vec3(1, 0, 0);
    else if (fabsf(from.y) < 0.72f && fabsf(to.y) < 0.72f) refl = Vec3(0, 1, 0);
    else refl = Vec3(0, 0, 1);
    Vec3 u = refl - from, v = refl - to;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r[i][j] = ((i == j) ? 1 : 0) - 2 / Dot(u, u) * u[i] * u[j] - 2 / Dot(v, v) * v[i] * v[j] + 4 * Dot(u, v) / (Dot(u, u) * Dot(v, v)) * v[i] * u[j];
}
BPT_HD Vec3 Mul33(const float r[3][3], Vec3 v) {
    return Vec3(r[0][0] * v.x + r[0][1] * v.y + r[0][2] * v.z, r[1][0] * v.x + r[1][1] * v.y + r[1][2] * v.z, r[2][0] * v.x + r[2][1] * v.y + r[2][2] * v.z);
}

// BilinearPatch::InteractionFromIntersection (the mesh's vertices are already in rendering space, BilinearPatchMesh ctor)
BPT_PATCH_FN SurfaceIntr BilinearInteraction(const PatchData &pd, float uu, float vv, SurfaceExtras *ex) {
    const Pt3 p00 = pd.p00, p10 = pd.p10, p01 = pd.p01, p11 = pd.p11;
    Pt3 p = LerpP(uu, LerpP(vv, p00, p01), LerpP(vv, p10, p11));
    Vec3 dpdu = LerpP(vv, p10, p11) - LerpP(vv, p00, p01);
    Vec3 dpdv = LerpP(uu, p01, p11) - LerpP(uu, p00, p10);
    Vec2 st(uu, vv);
    float duds = 1, dudt = 0, dvds = 0, dvdt = 1;
    if (pd.hasUV) {
        st = LerpUV(uu, LerpUV(vv, pd.uv00, pd.uv01), LerpUV(vv, pd.uv10, pd.uv11));
        Vec2 a = LerpUV(vv, pd.uv10, pd.uv11), b = LerpUV(vv, pd.uv00, pd.uv01);
        Vec2 dstdu(a.x - b.x, a.y - b.y);
        Vec2 c = LerpUV(uu, pd.uv01, pd.uv11), d = LerpUV(uu, pd.uv00, pd.uv10);
        Vec2 dstdv(c.x - d.x, c.y - d.y);
        duds = fabsf(dstdu.x) < 1e-8f ? 0 : 1 / dstdu.x;
        dvds = fabsf(dstdv.x) < 1e-8f ? 0 : 1 / dstdv.x;
        dudt = fabsf(dstdu.y) < 1e-8f ? 0 : 1 / dstdu.y;
        dvdt = fabsf(dstdv.y) < 1e-8f ? 0 : 1 / dstdv.y;
        Vec3 dpds = dpdu * duds + dpdv * dvds;
        Vec3 dpdt = dpdu * dudt + dpdv * dvdt;
        if (Cross(dpds, dpdt) != Vec3(0, 0, 0)) {
            if (Dot(Cross(dpdu, dpdv), Cross(dpds, dpdt)) < 0) dpdt = -dpdt;
            dpdu = dpds;
            dpdv = dpdt;
        }
    }
    Vec3 d2Pduv = (p00 - p01) + (p11 - p10);
    float E = Dot(dpdu, dpdu), F = Dot(dpdu, dpdv), G = Dot(dpdv, dpdv);
    Vec3 n = Normalize(Cross(dpdu, dpdv));
    float e = Dot(n, Vec3(0, 0, 0)), f = Dot(n, d2Pduv), g = Dot(n, Vec3(0, 0, 0));
    float EGF2 = DifferenceOfProducts(E, G, F, F);
    float invEGF2 = (EGF2 == 0) ? 0.f : 1 / EGF2;
    Nrm3 dndu = Nrm3((f * F - e * G) * invEGF2 * dpdu + (e * F - f * E) * invEGF2 * dpdv);
    Nrm3 dndv = Nrm3((g * F - f * G) * invEGF2 * dpdu + (f * F - g * E) * invEGF2 * dpdv);
    Nrm3 dnds = dndu * duds + dndv * dvds;
    Nrm3 dndt = dndu * dudt + dndv * dvdt;
    dndu = dnds;
    dndv = dndt;
    Pt3 pAbsSum = Abs(p00) + Abs(p01) + Abs(p10) + Abs(p11);
    Vec3 pError = gamma_n(6) * V(pAbsSum);

    SurfaceIntr isect;
    isect.pi = Pt3i(p, pError);
    isect.uv = st;
    isect.n = isect.ns = Nrm3(n);
    if (pd.flipNormal) isect.n = isect.ns = -isect.n;
    isect.dpdus = dpdu;
    if (ex) { ex->dpdu = dpdu; ex->dpdv = dpdv; ex->dpdvs = dpdv; ex->dndus = dndu; ex->dndvs = dndv; }
    if (pd.hasN) {
        Nrm3 ns = LerpN(uu, LerpN(vv, pd.n00, pd.n01), LerpN(vv, pd.n10, pd.n11));
        if (LengthSquared(ns) > 0) {
            ns = Normalize(ns);
            Nrm3 dndu2 = LerpN(vv, pd.n10, pd.n11) - LerpN(vv, pd.n00, pd.n01);
            Nrm3 dndv2 = LerpN(uu, pd.n01, pd.n11) - LerpN(uu, pd.n00, pd.n10);
            Nrm3 dnds2 = dndu2 * duds + dndv2 * dvds;
            Nrm3 dndt2 = dndu2 * dudt + dndv2 * dvdt;
            float r[3][3];
            RotateFromToMatrix(V(Normalize(isect.n)), V(ns), r);
            // SetShadingGeometry(ns, r(dpdu), r(dpdv), dndu, dndv, orientationIsAuthoritative = true)
            isect.ns = ns;
            isect.n = FaceForward(isect.n, isect.ns);
            Vec3 dpdus = Mul33(r, dpdu), dpdvs = Mul33(r, dpdv);
            while (LengthSquared(dpdus) > 1e16f || LengthSquared(dpdvs) > 1e16f) { dpdus = dpdus / 1e8f; dpdvs = dpdvs / 1e8f; }
            isect.dpdus = dpdus;
            if (ex) { ex->dpdvs = dpdvs; ex->dndus = dnds2; ex->dndvs = dndt2; }
        }
    }
    return isect;
}

// ---------------------------------------------------------------- construction (host)
BPT_HD bool PatchIsRectangle(Pt3 p00, Pt3 p10, Pt3 p01, Pt3 p11) {
    auto same = [](Pt3 a, Pt3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
    if (same(p00, p01) || same(p01, p11) || same(p11, p10) || same(p10, p00)) return false;
    Nrm3 n(Normalize(Cross(p10 - p00, p01 - p00)));
    if (AbsDot(Normalize(p11 - p00), n) > 1e-5f) return false;
    Pt3 pCenter = (p00 + p01 + p10 + p11) / 4;
    float d2[4] = {DistanceSquared(p00, pCenter), DistanceSquared(p01, pCenter), DistanceSquared(p10, pCenter), DistanceSquared(p11, pCenter)};
    for (int i = 1; i < 4; ++i)
        if (fabsf(d2[i] - d2[0]) / d2[0] > 1e-4f) return false;
    return true;
}
BPT_HD float PatchArea(Pt3 p00, Pt3 p10, Pt3 p01, Pt3 p11) {
    if (PatchIsRectangle(p00, p10, p01, p11)) return Distance(p00, p01) * Distance(p00, p10);
    const int na = 3;
    Pt3 p[na + 1][na + 1];
    for (int i = 0; i <= na; ++i) {
        float u = float(i) / float(na);
        for (int j = 0; j <= na; ++j) {
            float v = float(j) / float(na);
            p[i][j] = LerpP(u, LerpP(v, p00, p01), LerpP(v, p10, p11));
        }
    }
    float area = 0;
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < na; ++j) area += 0.5f * Length(Cross(p[i + 1][j + 1] - p[i][j], p[i + 1][j] - p[i][j + 1]));
    return area;
}
// BilinearPatch::NormalBounds: the cone's axis and cosTheta
BPT_HD void PatchNormalBounds(const PatchData &pd, Vec3 *w, float *cosTheta) {
    const Pt3 p00 = pd.p00, p10 = pd.p10, p01 = pd.p01, p11 = pd.p11;
    auto same = [](Pt3 a, Pt3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; };
    if (same(p00, p10) || same(p10, p11) || same(p11, p01) || same(p01, p00)) {
        Vec3 dpdu = LerpP(0.5f, p10, p11) - LerpP(0.5f, p00, p01);
        Vec3 dpdv = LerpP(0.5f, p01, p11) - LerpP(0.5f, p00, p10);
        Vec3 n = Normalize(Cross(dpdu, dpdv));
        if (pd.hasN) {
            Nrm3 ns = (pd.n00 + pd.n10 + pd.n01 + pd.n11) / 4;
            n = FaceForward(n, ns);
        } else if (pd.flipNormal) n = -n;
        *w = Normalize(n); *cosTheta = 1;
        return;
    }
    Vec3 n00 = Normalize(Cross(p10 - p00, p01 - p00));
    if (pd.hasN) n00 = FaceForward(n00, pd.n00);
    else if (pd.flipNormal) n00 = -n00;
    Vec3 n10 = Normalize(Cross(p11 - p10, p00 - p10));
    Vec3 n01 = Normalize(Cross(p00 - p01, p11 - p01));
    Vec3 n11 = Normalize(Cross(p01 - p11, p10 - p11));
    if (pd.hasN) {
        n10 = FaceForward(n10, pd.n10);
        n01 = FaceForward(n01, pd.n01);
        n11 = FaceForward(n11, pd.n11);
    } else if (pd.flipNormal) { n10 = -n10; n01 = -n01; n11 = -n11; }
    Vec3 n = Normalize(n00 + n10 + n01 + n11);
    float c = stdmin(stdmin(stdmin(Dot(n, n00), Dot(n, n01)), Dot(n, n10)), Dot(n, n11));
    *w = Normalize(n);   // DirectionCone(w, cosTheta): w(Normalize(w))
    *cosTheta = Clampf(c, -1, 1);
}

// ---------------------------------------------------------------- sampling
BPT_HD float SphericalQuadArea(Vec3 a, Vec3 b, Vec3 c, Vec3 d) {
    Vec3 axb = Cross(a, b), bxc = Cross(b, c);
    Vec3 cxd = Cross(c, d), dxa = Cross(d, a);
    if (LengthSquared(axb) == 0 || LengthSquared(bxc) == 0 || LengthSquared(cxd) == 0 || LengthSquared(dxa) == 0) return 0;
    axb = Normalize(axb); bxc = Normalize(bxc); cxd = Normalize(cxd); dxa = Normalize(dxa);
    float alpha = AngleBetween(dxa, -axb);
    float beta = AngleBetween(axb, -bxc);
    float gamma = AngleBetween(bxc, -cxd);
    float delta = AngleBetween(cxd, -dxa);
    return fabsf(alpha + beta + gamma + delta - 2 * kPi);
}

BPT_PATCH_FN Pt3 SampleSphericalRectangle(Pt3 pRef, Pt3 s, Vec3 ex, Vec3 ey, Vec2 u, float *pdf) {
    float exl = Length(ex), eyl = Length(ey);
    Frame R; R.x = ex / exl; R.y = ey / eyl; R.z = Cross(R.x, R.y);   // Frame::FromXY
    Vec3 dLocal = R.ToLocal(s - pRef);
    float z0 = dLocal.z;
    if (z0 > 0) { R.z = -R.z; z0 *= -1; }
    float x0 = dLocal.x, y0 = dLocal.y;
    float x1 = x0 + exl, y1 = y0 + eyl;
    Vec3 v00(x0, y0, z0), v01(x0, y1, z0);
    Vec3 v10(x1, y0, z0), v11(x1, y1, z0);
    Vec3 n0 = Normalize(Cross(v00, v10)), n1 = Normalize(Cross(v10, v11));
    Vec3 n2 = Normalize(Cross(v11, v01)), n3 = Normalize(Cross(v01, v00));
    float g0 = AngleBetween(-n0, n1), g1 = AngleBetween(-n1, n2);
    float g2 = AngleBetween(-n2, n3), g3 = AngleBetween(-n3, n0);
    float solidAngle = g0 + g1 + g2 + g3 - 2 * kPi;
    if (solidAngle <= 0) {
        *pdf = 0;
        return s + u.x * ex + u.y * ey;
    }
    *pdf = stdmax(0.f, 1 / solidAngle);
    if (solidAngle < 1e-3) return s + u.x * ex + u.y * ey;
    float b0 = n0.z, b1 = n2.z;
    float au = u.x * (g0 + g1 - 2 * kPi) + (u.x - 1) * (g2 + g3);
    float fu = (CosF(au) * b0 - b1) / SinF(au);
    float cu = copysignf(1 / sqrtf(Sqr(fu) + Sqr(b0)), fu);
    cu = Clampf(cu, -kOneMinusEpsilon, kOneMinusEpsilon);
    float xu = -(cu * z0) / SafeSqrt(1 - Sqr(cu));
    xu = Clampf(xu, x0, x1);
    float dd = sqrtf(Sqr(xu) + Sqr(z0));
    float h0 = y0 / sqrtf(Sqr(dd) + Sqr(y0));
    float h1 = y1 / sqrtf(Sqr(dd) + Sqr(y1));
    float hv = h0 + u.y * (h1 - h0), hvsq = Sqr(hv);
    float yv = (hvsq < 1 - 1e-6f) ? (hv * dd) / sqrtf(1 - hvsq) : y1;
    return pRef + R.FromLocal(Vec3(xu, yv, z0));
}

BPT_PATCH_FN Vec2 InvertSphericalRectangleSample(Pt3 pRef, Pt3 s, Vec3 ex, Vec3 ey, Pt3 pRect) {
    float exl = Length(ex), eyl = Length(ey);
    Frame R; R.x = ex / exl; R.y = ey / eyl; R.z = Cross(R.x, R.y);
    Vec3 d = s - pRef;
    Vec3 dLocal = R.ToLocal(d);
    float z0 = dLocal.z;
    if (z0 > 0) { R.z = -R.z; z0 *= -1; }
    float z0sq = Sqr(z0);
    float x0 = dLocal.x, y0 = dLocal.y;
    float x1 = x0 + exl, y1 = y0 + eyl;
    float y0sq = Sqr(y0), y1sq = Sqr(y1);
    Vec3 v00(x0, y0, z0), v01(x0, y1, z0);
    Vec3 v10(x1, y0, z0), v11(x1, y1, z0);
    Vec3 n0 = Normalize(Cross(v00, v10)), n1 = Normalize(Cross(v10, v11));
    Vec3 n2 = Normalize(Cross(v11, v01)), n3 = Normalize(Cross(v01, v00));
    float g0 = AngleBetween(-n0, n1), g1 = AngleBetween(-n1, n2);
    float g2 = AngleBetween(-n2, n3), g3 = AngleBetween(-n3, n0);
    float b0 = n0.z, b1 = n2.z, b0sq = Sqr(b0);
    float solidAngle = (float)((double)g0 + (double)g1 + (double)g2 + (double)g3 - 2. * (double)kPi);
    if (solidAngle < 1e-3) {
        Vec3 pq = pRect - s;
        return Vec2(Dot(pq, ex) / LengthSquared(ex), Dot(pq, ey) / LengthSquared(ey));
    }
    Vec3 v = R.ToLocal(pRect - pRef);
    float xu = v.x, yv = v.y;
    xu = Clampf(xu, x0, x1);
    if (xu == 0) xu = (float)1e-10;
    float invcusq = 1 + z0sq / Sqr(xu);
    float fusq = invcusq - b0sq;
    float fu = copysignf(sqrtf(fusq), xu);
    float sq = SafeSqrt(DifferenceOfProducts(b0, b0, b1, b1) + fusq);
    float au = ATan2F(-(b1 * fu) - copysignf(b0 * sq, fu * b0), b0 * b1 - sq * fabsf(fu));
    if (au > 0) au -= 2 * kPi;
    if (fu == 0) au = kPi;
    float u0 = (au + g2 + g3) / solidAngle;
    float ddsq = Sqr(xu) + z0sq;
    float dd = sqrtf(ddsq);
    float h0 = y0 / sqrtf(ddsq + y0sq);
    float h1 = y1 / sqrtf(ddsq + y1sq);
    float yvsq = Sqr(yv);
    float u1[2] = {(DifferenceOfProducts(h0, h0, h0, h1) - fabsf(h0 - h1) * sqrtf(yvsq * (ddsq + yvsq)) / (ddsq + yvsq)) / Sqr(h0 - h1),
                   (DifferenceOfProducts(h0, h0, h0, h1) + fabsf(h0 - h1) * sqrtf(yvsq * (ddsq + yvsq)) / (ddsq + yvsq)) / Sqr(h0 - h1)};
    float hv[2] = {Lerp(u1[0], h0, h1), Lerp(u1[1], h0, h1)};
    float hvsq[2] = {Sqr(hv[0]), Sqr(hv[1])};
    float yz[2] = {(hv[0] * dd) / sqrtf(1 - hvsq[0]), (hv[1] * dd) / sqrtf(1 - hvsq[1])};
    return (fabsf(yz[0] - yv) < fabsf(yz[1] - yv)) ? Vec2(Clampf(u0, 0, 1), u1[0]) : Vec2(Clampf(u0, 0, 1), u1[1]);
}

BPT_HD Vec2 InvertBilinear(Vec2 p, Vec2 v0, Vec2 v1, Vec2 v2, Vec2 v3) {
    Vec2 a = v0, b = v1, c = v3, d = v2;
    Vec2 e(b.x - a.x, b.y - a.y), f(d.x - a.x, d.y - a.y), g((a.x - b.x) + (c.x - d.x), (a.y - b.y) + (c.y - d.y)), h(p.x - a.x, p.y - a.y);
    auto cross2d = [](Vec2 a, Vec2 b) { return DifferenceOfProducts(a.x, b.y, a.y, b.x); };
    float k2 = cross2d(g, f);
    float k1 = cross2d(e, f) + cross2d(h, g);
    float k0 = cross2d(h, e);
    if (fabsf(k2) < 0.001f) {
        if (fabsf(e.x * k1 - g.x * k0) < 1e-5f) return Vec2((h.y * k1 + f.y * k0) / (e.y * k1 - g.y * k0), -k0 / k1);
        else return Vec2((h.x * k1 + f.x * k0) / (e.x * k1 - g.x * k0), -k0 / k1);
    }
    float w0, w1;
    if (!QuadraticF(k2, k1, k0, &w0, &w1)) return Vec2(0, 0);
    float u = (h.x - f.x * w0) / (e.x + g.x * w0);
    if (u < 0 || u > 1 || w0 < 0 || w0 > 1) return Vec2((h.x - f.x * w1) / (e.x + g.x * w1), w1);
    return Vec2(u, w0);
}

BPT_HD Nrm3 PatchSampleNormal(const PatchData &pd, Vec3 dpdu, Vec3 dpdv, float uu, float vv) {
    Nrm3 n = Nrm3(Normalize(Cross(dpdu, dpdv)));
    if (pd.hasN) {
        Nrm3 ns = LerpN(uu, LerpN(vv, pd.n00, pd.n01), LerpN(vv, pd.n10, pd.n11));
        n = FaceForward(n, ns);
    } else if (pd.flipNormal) n = -n;
    return n;
}

// BilinearPatch::Sample(Point2f u)
BPT_HD ShapeSample PatchSampleArea(const PatchData &pd, Vec2 u) {
    const Pt3 p00 = pd.p00, p10 = pd.p10, p01 = pd.p01, p11 = pd.p11;
    ShapeSample ss; ss.valid = false;
    float pdf = 1;
    Vec2 uv;
    if (!pd.isRectangle) {
        float w[4] = {Length(Cross(p10 - p00, p01 - p00)), Length(Cross(p10 - p00, p11 - p10)), Length(Cross(p01 - p00, p11 - p01)), Length(Cross(p11 - p10, p11 - p01))};
        uv = SampleBilinear(u, w);
        pdf = BilinearPDF(uv, w);
    } else uv = u;
    Pt3 pu0 = LerpP(uv.y, p00, p01), pu1 = LerpP(uv.y, p10, p11);
    Pt3 p = LerpP(uv.x, pu0, pu1);
    Vec3 dpdu = pu1 - pu0;
    Vec3 dpdv = LerpP(uv.x, p01, p11) - LerpP(uv.x, p00, p10);
    if (LengthSquared(dpdu) == 0 || LengthSquared(dpdv) == 0) return ss;
    Vec2 st = uv;
    if (pd.hasUV) st = LerpUV(uv.x, LerpUV(uv.y, pd.uv00, pd.uv01), LerpUV(uv.y, pd.uv10, pd.uv11));
    ss.n = PatchSampleNormal(pd, dpdu, dpdv, uv.x, uv.y);
    Pt3 pAbsSum = Abs(p00) + Abs(p01) + Abs(p10) + Abs(p11);
    Vec3 pError = gamma_n(6) * V(pAbsSum);
    ss.pi = Pt3i(p, pError);
    ss.uv = st;
    ss.pdf = pdf / Length(Cross(dpdu, dpdv));
    ss.valid = true;
    return ss;
}
// BilinearPatch::PDF(const Interaction &): intrUV = intr.uv
BPT_HD float PatchPDFArea(const PatchData &pd, Vec2 intrUV) {
    const Pt3 p00 = pd.p00, p10 = pd.p10, p01 = pd.p01, p11 = pd.p11;
    Vec2 uv = intrUV;
    if (pd.hasUV) uv = InvertBilinear(uv, pd.uv00, pd.uv10, pd.uv01, pd.uv11);
    float pdf;
    if (!pd.isRectangle) {
        float w[4] = {Length(Cross(p10 - p00, p01 - p00)), Length(Cross(p10 - p00, p11 - p10)), Length(Cross(p01 - p00, p11 - p01)), Length(Cross(p11 - p10, p11 - p01))};
        pdf = BilinearPDF(uv, w);
    } else pdf = 1;
    Pt3 pu0 = LerpP(uv.y, p00, p01), pu1 = LerpP(uv.y, p10, p11);
    Vec3 dpdu = pu1 - pu0;
    Vec3 dpdv = LerpP(uv.x, p01, p11) - LerpP(uv.x, p00, p10);
    return pdf / Length(Cross(dpdu, dpdv));
}

constexpr float kPatchMinSphericalSampleArea = 1e-4f;

// BilinearPatch::Sample(const ShapeSampleContext &ctx, Point2f u)
BPT_PATCH_FN ShapeSample PatchSampleSolidAngle(const PatchData &pd, Pt3 ctxP, Nrm3 ctxNs, Vec2 u) {
    const Pt3 p00 = pd.p00, p10 = pd.p10, p01 = pd.p01, p11 = pd.p11;
    Vec3 v00 = Normalize(p00 - ctxP), v10 = Normalize(p10 - ctxP);
    Vec3 v01 = Normalize(p01 - ctxP), v11 = Normalize(p11 - ctxP);
    if (!pd.isRectangle || SphericalQuadArea(v00, v10, v11, v01) <= kPatchMinSphericalSampleArea) {
        ShapeSample ss = PatchSampleArea(pd, u);
        if (!ss.valid) return ss;   // (the reference DCHECKs has_value())
        Vec3 wi = ss.pi.p() - ctxP;
        if (LengthSquared(wi) == 0) { ss.valid = false; return ss; }
        wi = Normalize(wi);
        ss.pdf /= AbsDot(ss.n, -wi) / DistanceSquared(ctxP, ss.pi.p());
        if (IsInf_(ss.pdf)) ss.valid = false;
        return ss;
    }
    float pdf = 1;
    if (ctxNs != Nrm3(0, 0, 0)) {
        float w[4] = {stdmax(0.01f, AbsDot(v00, ctxNs)), stdmax(0.01f, AbsDot(v10, ctxNs)), stdmax(0.01f, AbsDot(v01, ctxNs)), stdmax(0.01f, AbsDot(v11, ctxNs))};
        u = SampleBilinear(u, w);
        pdf *= BilinearPDF(u, w);
    }
    Vec3 eu = p10 - p00, ev = p01 - p00;
    float quadPDF;
    Pt3 p = SampleSphericalRectangle(ctxP, p00, eu, ev, u, &quadPDF);
    pdf *= quadPDF;
    Vec2 uv(Dot(p - p00, eu) / DistanceSquared(p10, p00), Dot(p - p00, ev) / DistanceSquared(p01, p00));
    Nrm3 n = Nrm3(Normalize(Cross(eu, ev)));
    if (pd.hasN) {
        Nrm3 ns = LerpN(uv.x, LerpN(uv.y, pd.n00, pd.n01), LerpN(uv.y, pd.n10, pd.n11));
        n = FaceForward(n, ns);
    } else if (pd.flipNormal) n = -n;
    Vec2 st = uv;
    if (pd.hasUV) st = LerpUV(uv.x, LerpUV(uv.y, pd.uv00, pd.uv01), LerpUV(uv.y, pd.uv10, pd.uv11));
    ShapeSample ss;
    ss.pi = Pt3i(p);
    ss.n = n; ss.uv = st; ss.pdf = pdf; ss.valid = true;
    return ss;
}

// BilinearPatch::PDF(const ShapeSampleContext &ctx, Vector3f wi)
BPT_PATCH_FN float PatchPDFSolidAngle(const PatchData &pd, Pt3i ctxPi, Nrm3 ctxN, Nrm3 ctxNs, Vec3 wi) {
    const Pt3 p00 = pd.p00, p10 = pd.p10, p01 = pd.p01, p11 = pd.p11;
    const Pt3 ctxP = ctxPi.p();
    Pt3 o = OffsetRayOrigin(ctxPi, ctxN, wi);
    float uu, vv, t;
    if (!IntersectBilinearPatch(o, wi, BPT_INFINITY, p00, p10, p01, p11, &uu, &vv, &t)) return 0;
    SurfaceIntr isect = BilinearInteraction(pd, uu, vv, nullptr);
    Vec3 v00 = Normalize(p00 - ctxP), v10 = Normalize(p10 - ctxP);
    Vec3 v01 = Normalize(p01 - ctxP), v11 = Normalize(p11 - ctxP);
    if (!pd.isRectangle || SphericalQuadArea(v00, v10, v11, v01) <= kPatchMinSphericalSampleArea) {
        float pdf = PatchPDFArea(pd, isect.uv) * (DistanceSquared(ctxP, isect.p()) / AbsDot(isect.n, -wi));
        return IsInf_(pdf) ? 0 : pdf;
    }
    float pdf = 1 / SphericalQuadArea(v00, v10, v11, v01);
    if (ctxNs != Nrm3(0, 0, 0)) {
        float w[4] = {stdmax(0.01f, AbsDot(v00, ctxNs)), stdmax(0.01f, AbsDot(v10, ctxNs)), stdmax(0.01f, AbsDot(v01, ctxNs)), stdmax(0.01f, AbsDot(v11, ctxNs))};
        Vec2 u = InvertSphericalRectangleSample(ctxP, p00, p10 - p00, p01 - p00, isect.p());
        return BilinearPDF(u, w) * pdf;
    }
    return pdf;
}

}  // namespace bpt
