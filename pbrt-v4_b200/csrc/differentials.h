// differentials.h -- ray differentials along a path and the screen-space derivatives texture filtering reads (host + device).
//   SurfaceInteraction::ComputeDifferentials      interaction.cpp:41-89
//   Camera::Approximate_dp_dxy                    cameras.h:155-183   (RotateFromTo util/transform.h:249-270)
//   SurfaceInteraction::SpawnRay(rayi, bsdf, ...) interaction.cpp:99-150  (differentials survive specular bounces only)
//   SurfaceInteraction::SkipIntersection          interaction.cpp:91-97
//   BumpMap                                       materials.h:86-130
#pragma once
#include "scene.h"
#include "camera.h"
#include "shapes.h"
#include "texture.h"
#include "media.h"

namespace bpt {

struct RayDiffState {
    bool has = false;
    Pt3 rxO, ryO;
    Vec3 rxD, ryD;
};

// Transform::ApplyInverse(Normal3f) (ApplyInverse(Point3f): ApplyInversePoint in media.h)
BPT_HD bool IsFinite_(float v) { return !IsInf_(v) && !IsNaN_(v); }
BPT_HD Nrm3 ApplyInverseNormal(const Mat4 &M, Nrm3 n) {
    const float (*m)[4] = M.m;
    float x = n.x, y = n.y, z = n.z;
    return Nrm3(m[0][0] * x + m[1][0] * y + m[2][0] * z, m[0][1] * x + m[1][1] * y + m[2][1] * z, m[0][2] * x + m[1][2] * y + m[2][2] * z);
}

// Camera::Approximate_dp_dxy: the rotation that takes the direction to p onto +z, two fixed "minimum differential" rays intersected
// with the tangent plane there
BPT_HD void ApproximateDpDxy(const DeviceCamera &cam, Pt3 p, Nrm3 n, Vec3 *dpdx, Vec3 *dpdy) {
    Pt3 pCamera = ApplyInversePoint(cam.cameraFromRender, p);
    Vec3 from = Normalize(Vec3(pCamera.x, pCamera.y, pCamera.z)), to(0, 0, 1);
    Vec3 refl;
    if (fabsf(from.x) < 0.72f && fabsf(to.x) < 0.72f) refl = Vec3(1, 0, 0);
    else if (fabsf(from.y) < 0.72f && fabsf(to.y) < 0.72f) refl = Vec3(0, 1, 0);
    else refl = Vec3(0, 0, 1);
    Vec3 u = refl - from, v = refl - to;
    float r[3][3];
    const float uu = Dot(u, u), vv = Dot(v, v), uv = Dot(u, v);
    const float ua[3] = {u.x, u.y, u.z}, va[3] = {v.x, v.y, v.z};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            r[i][j] = ((i == j) ? 1 : 0) - 2 / uu * ua[i] * ua[j] - 2 / vv * va[i] * va[j] + 4 * uv / (uu * vv) * va[i] * ua[j];
    // DownZFromCamera = Transform(r, Transpose(r))
    auto fwdP = [&](Pt3 q) { return Pt3(r[0][0] * q.x + r[0][1] * q.y + r[0][2] * q.z + 0.f, r[1][0] * q.x + r[1][1] * q.y + r[1][2] * q.z + 0.f,
                                        r[2][0] * q.x + r[2][1] * q.y + r[2][2] * q.z + 0.f); };
    auto fwdN = [&](Nrm3 q) { return Nrm3(r[0][0] * q.x + r[0][1] * q.y + r[0][2] * q.z, r[1][0] * q.x + r[1][1] * q.y + r[1][2] * q.z,
                                          r[2][0] * q.x + r[2][1] * q.y + r[2][2] * q.z); };
    auto invV = [&](Vec3 q) { return Vec3(r[0][0] * q.x + r[1][0] * q.y + r[2][0] * q.z, r[0][1] * q.x + r[1][1] * q.y + r[2][1] * q.z,
                                          r[0][2] * q.x + r[1][2] * q.y + r[2][2] * q.z); };
    Pt3 pDownZ = fwdP(pCamera);
    Nrm3 nDownZ = fwdN(ApplyInverseNormal(cam.renderFromCamera, n));
    float d = nDownZ.z * pDownZ.z;
    Pt3 xo = Pt3(0, 0, 0) + cam.minPosDifferentialX; Vec3 xd = Vec3(0, 0, 1) + cam.minDirDifferentialX;
    float tx = -(Dot(nDownZ, Vec3(xo.x, xo.y, xo.z)) - d) / Dot(nDownZ, xd);
    Pt3 yo = Pt3(0, 0, 0) + cam.minPosDifferentialY; Vec3 yd = Vec3(0, 0, 1) + cam.minDirDifferentialY;
    float ty = -(Dot(nDownZ, Vec3(yo.x, yo.y, yo.z)) - d) / Dot(nDownZ, yd);
    Pt3 px = xo + xd * tx, py = yo + yd * ty;
    *dpdx = XformVector(cam.renderFromCamera, invV(px - pDownZ)) * cam.sppScale;
    *dpdy = XformVector(cam.renderFromCamera, invV(py - pDownZ)) * cam.sppScale;
}

struct UVDiffs { Vec3 dpdx, dpdy; float dudx, dudy, dvdx, dvdy; };

BPT_HD UVDiffs ComputeDifferentials(const DeviceCamera &cam, const SurfaceIntr &isect, const SurfaceExtras &ex, const RayDiffState &ray) {
    UVDiffs o;
    const Nrm3 n = isect.n;
    const Pt3 P = isect.p();
    if (ray.has && Dot(n, ray.rxD) != 0 && Dot(n, ray.ryD) != 0) {
        float d = -Dot(n, Vec3(P.x, P.y, P.z));
        float tx = (-Dot(n, Vec3(ray.rxO.x, ray.rxO.y, ray.rxO.z)) - d) / Dot(n, ray.rxD);
        Pt3 px = ray.rxO + ray.rxD * tx;
        float ty = (-Dot(n, Vec3(ray.ryO.x, ray.ryO.y, ray.ryO.z)) - d) / Dot(n, ray.ryD);
        Pt3 py = ray.ryO + ray.ryD * ty;
        o.dpdx = px - P; o.dpdy = py - P;
    } else
        ApproximateDpDxy(cam, P, n, &o.dpdx, &o.dpdy);
    float ata00 = Dot(ex.dpdu, ex.dpdu), ata01 = Dot(ex.dpdu, ex.dpdv), ata11 = Dot(ex.dpdv, ex.dpdv);
    float invDet = 1 / DifferenceOfProducts(ata00, ata11, ata01, ata01);
    invDet = IsFinite_(invDet) ? invDet : 0.f;
    float atb0x = Dot(ex.dpdu, o.dpdx), atb1x = Dot(ex.dpdv, o.dpdx);
    float atb0y = Dot(ex.dpdu, o.dpdy), atb1y = Dot(ex.dpdv, o.dpdy);
    o.dudx = DifferenceOfProducts(ata11, atb0x, ata01, atb1x) * invDet;
    o.dvdx = DifferenceOfProducts(ata00, atb1x, ata01, atb0x) * invDet;
    o.dudy = DifferenceOfProducts(ata11, atb0y, ata01, atb1y) * invDet;
    o.dvdy = DifferenceOfProducts(ata00, atb1y, ata01, atb0y) * invDet;
    o.dudx = IsFinite_(o.dudx) ? Clampf(o.dudx, -1e8f, 1e8f) : 0.f;
    o.dvdx = IsFinite_(o.dvdx) ? Clampf(o.dvdx, -1e8f, 1e8f) : 0.f;
    o.dudy = IsFinite_(o.dudy) ? Clampf(o.dudy, -1e8f, 1e8f) : 0.f;
    o.dvdy = IsFinite_(o.dvdy) ? Clampf(o.dvdy, -1e8f, 1e8f) : 0.f;
    return o;
}

// SurfaceInteraction::SpawnRay(rayi, bsdf, wi, flags, eta): the differentials of the spawned ray.  ns / dndus / dndvs are the
// shading geometry the BSDF was built on (after bump mapping), flags the sampled component's BxDFFlags.
BPT_HD RayDiffState SpawnRayDifferentials(const RayDiffState &rayi, Pt3 P, Vec3 wo, Vec3 wi, Nrm3 ns, Nrm3 dndus, Nrm3 dndvs, const UVDiffs &dv,
                                          int flags, float eta) {
    RayDiffState rd;
    if (rayi.has) {
        Nrm3 n = ns;
        Nrm3 dndx = dndus * dv.dudx + dndvs * dv.dvdx;
        Nrm3 dndy = dndus * dv.dudy + dndvs * dv.dvdy;
        Vec3 dwodx = -rayi.rxD - wo, dwody = -rayi.ryD - wo;
        if (flags == (BXDF_SPECULAR | BXDF_REFLECTION)) {
            rd.has = true;
            rd.rxO = P + dv.dpdx; rd.ryO = P + dv.dpdy;
            float dwoDotn_dx = Dot(dwodx, n) + Dot(wo, dndx);
            float dwoDotn_dy = Dot(dwody, n) + Dot(wo, dndy);
            rd.rxD = wi - dwodx + 2 * V(Dot(wo, n) * dndx + dwoDotn_dx * n);
            rd.ryD = wi - dwody + 2 * V(Dot(wo, n) * dndy + dwoDotn_dy * n);
        } else if (flags == (BXDF_SPECULAR | BXDF_TRANSMISSION)) {
            rd.has = true;
            rd.rxO = P + dv.dpdx; rd.ryO = P + dv.dpdy;
            if (Dot(wo, n) < 0) { n = -n; dndx = -dndx; dndy = -dndy; }
            float dwoDotn_dx = Dot(dwodx, n) + Dot(wo, dndx);
            float dwoDotn_dy = Dot(dwody, n) + Dot(wo, dndy);
            float mu = Dot(wo, n) / eta - AbsDot(wi, n);
            float dmudx = dwoDotn_dx * (1 / eta + 1 / Sqr(eta) * Dot(wo, n) / Dot(wi, n));
            float dmudy = dwoDotn_dy * (1 / eta + 1 / Sqr(eta) * Dot(wo, n) / Dot(wi, n));
            rd.rxD = wi - eta * dwodx + V(mu * dndx + dmudx * n);
            rd.ryD = wi - eta * dwody + V(mu * dndy + dmudy * n);
        }
    }
    if (LengthSquared(rd.rxD) > 1e16f || LengthSquared(rd.ryD) > 1e16f || LengthSquared(Vec3(rd.rxO.x, rd.rxO.y, rd.rxO.z)) > 1e16f ||
        LengthSquared(Vec3(rd.ryO.x, rd.ryO.y, rd.ryO.z)) > 1e16f)
        rd.has = false;
    return rd;
}

}  // namespace bpt
