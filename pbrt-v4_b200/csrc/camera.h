// camera.h -- transforms and the perspective camera of the path (host + device so that tests/tools can pin them to the reference):
//   Transform::operator()(Point3f / Vector3f / Normal3f / Ray)            util/transform.h:310-349
//   GetCameraSample                                                        samplers.h:797-814
//   PerspectiveCamera::GenerateRayDifferential                             cameras.cpp:429-478
//   RayDifferential::ScaleDifferentials                                    ray.h:53-58
#pragma once
#include "pmath.h"
#include "sampling.h"
#include "scene.h"

namespace bpt {

BPT_HD Pt3 XformPoint(const Mat4 &T, Pt3 p) {
    const float (*m)[4] = T.m;
    float xp = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
    float yp = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
    float zp = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
    float wp = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
    if (wp == 1) return Pt3(xp, yp, zp);
    return Pt3(xp, yp, zp) / wp;
}
BPT_HD Vec3 XformVector(const Mat4 &T, Vec3 v) {
    const float (*m)[4] = T.m;
    return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
}
// Normal3f transformed by the transpose of `A` (Transform::operator()(Normal3f): A = mInv; Transform::ApplyInverse(Normal3f): A = m)
BPT_HD Nrm3 XformNormalByTranspose(const Mat4 &A, Nrm3 n) {
    const float (*m)[4] = A.m;
    float x = n.x, y = n.y, z = n.z;
    return Nrm3(m[0][0] * x + m[1][0] * y + m[2][0] * z, m[0][1] * x + m[1][1] * y + m[2][1] * z, m[0][2] * x + m[1][2] * y + m[2][2] * z);
}
// Transform::operator()(const Ray &r) (util/transform.h:336-349) for an exact origin
BPT_HD void XformRay(const Mat4 &T, Pt3 &o, Vec3 &d) {
    const float (*m)[4] = T.m;
    float x = o.x, y = o.y, z = o.z;
    float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
    float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
    float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
    float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
    Vec3 pError;
    pError.x = gamma_n(3) * (fabsf(m[0][0] * x) + fabsf(m[0][1] * y) + fabsf(m[0][2] * z) + fabsf(m[0][3]));
    pError.y = gamma_n(3) * (fabsf(m[1][0] * x) + fabsf(m[1][1] * y) + fabsf(m[1][2] * z) + fabsf(m[1][3]));
    pError.z = gamma_n(3) * (fabsf(m[2][0] * x) + fabsf(m[2][1] * y) + fabsf(m[2][2] * z) + fabsf(m[2][3]));
    Pt3i oi(Pt3(xp, yp, zp), pError);
    if (wp != 1) {
        // Point3fi / Float: Interval::operator/=(Float)
        auto divI = [](Interval i, float f) { return f > 0 ? Interval(DivRoundDown(i.low, f), DivRoundUp(i.high, f))
                                                          : Interval(DivRoundDown(i.high, f), DivRoundUp(i.low, f)); };
        oi.x = divI(oi.x, wp); oi.y = divI(oi.y, wp); oi.z = divI(oi.z, wp);
    }
    Vec3 dn = XformVector(T, d);
    float lengthSquared = LengthSquared(dn);
    if (lengthSquared > 0) {
        float dt = Dot(Abs(dn), oi.Error()) / lengthSquared;
        Vec3 off = dn * dt;
        oi.x = oi.x + off.x; oi.y = oi.y + off.y; oi.z = oi.z + off.z;
    }
    o = oi.p();
    d = dn;
}


// wavelength sample + GetCameraSample (samplers.h:797-814): dimensions 0..5 of the pixel sample
struct CamSample { float lu; Vec2 pFilm; float fweight; float time; Vec2 pLens; };
BPT_HD CamSample GetCameraSample(const SamplerParams &sp, const DeviceFilter &filter, Sampler &sampler, int px, int py, int sampleIndex) {
    CamSample cs;
    sampler.StartPixelSample(&sp, px, py, sampleIndex, 0);
    cs.lu = sampler.Get1D();
    Vec2 uf = sampler.GetPixel2D();
    Vec2 fp;
    if (filter.sampled == 2) {   // TriangleFilter::Sample (filters.h:224-227): SampleTent per axis, weight 1
        fp = Vec2(SampleTent(uf.x, filter.radiusX), SampleTent(uf.y, filter.radiusY));
        cs.fweight = 1;
    } else if (filter.sampled) {
        float pdf; int ou, ov;
        fp = PC2DSample(filter.distrib, uf, &pdf, &ou, &ov);
        cs.fweight = filter.f[ov * filter.nx + ou] / pdf;
    } else {   // BoxFilter::Sample (filters.h)
        fp = Vec2(Lerp(uf.x, -filter.radiusX, filter.radiusX), Lerp(uf.y, -filter.radiusY, filter.radiusY));
        cs.fweight = 1;
    }
    cs.pFilm = Vec2((px + fp.x) + 0.5f, (py + fp.y) + 0.5f);
    cs.time = sampler.Get1D();
    cs.pLens = sampler.Get2D();
    return cs;
}

// PerspectiveCamera::GenerateRayDifferential (cameras.cpp:429-447), main ray, taken to render space (CameraTransform::RenderFromCamera)
BPT_HD void GenerateCameraRay(const DeviceCamera &cam, Vec2 pFilm, Vec2 pLens, Pt3 *oOut, Vec3 *dOut) {
    Pt3 pCamera = XformPoint(cam.cameraFromRaster, Pt3(pFilm.x, pFilm.y, 0));
    if (cam.type == 1) {   // OrthographicCamera::GenerateRayDifferential (cameras.cpp:308-329): origin on the film plane, direction +z
        Pt3 o = pCamera;
        Vec3 dir(0, 0, 1);
        if (cam.lensRadius > 0) {
            Vec2 dsk = SampleUniformDiskConcentric(pLens);
            Vec2 pl(cam.lensRadius * dsk.x, cam.lensRadius * dsk.y);
            float ft = cam.focalDistance / dir.z;
            Pt3 pFocus = o + dir * ft;
            o = Pt3(pl.x, pl.y, 0);
            dir = Normalize(pFocus - o);
        }
        XformRay(cam.renderFromCamera, o, dir);
        *oOut = o; *dOut = dir;
        return;
    }
    Vec3 dir = Normalize(Vec3(pCamera.x, pCamera.y, pCamera.z));
    Pt3 o(0, 0, 0);
    if (cam.lensRadius > 0) {
        Vec2 dsk = SampleUniformDiskConcentric(pLens);
        Vec2 pl(cam.lensRadius * dsk.x, cam.lensRadius * dsk.y);
        float ft = cam.focalDistance / dir.z;
        Pt3 pFocus = o + dir * ft;
        o = Pt3(pl.x, pl.y, 0);
        dir = Normalize(pFocus - o);
    }
    XformRay(cam.renderFromCamera, o, dir);
    *oOut = o; *dOut = dir;
}

// GBufferFilm: the camera ray differentials of a pixel sample in render space, scaled as the integrator does
// (PerspectiveCamera::GenerateRayDifferential cameras.cpp:450-477, Transform::operator()(RayDifferential) util/transform.h,
// RayDifferential::ScaleDifferentials ray.h:53-58).  o, d: the main ray in render space.
struct RayDiffs { Pt3 rxO, ryO; Vec3 rxD, ryD; };
BPT_HD RayDiffs CameraRayDifferentials(const DeviceCamera &cam, const CamSample &cs, Pt3 o, Vec3 d, bool scaleDifferentials = true) {
    Pt3 pCamera = XformPoint(cam.cameraFromRaster, Pt3(cs.pFilm.x, cs.pFilm.y, 0));
    RayDiffs r;
    if (cam.type == 1) {   // OrthographicCamera::GenerateRayDifferential (cameras.cpp:330-352), camera space
        if (cam.lensRadius > 0) {
            Vec2 dsk = SampleUniformDiskConcentric(cs.pLens);
            Vec2 pl(cam.lensRadius * dsk.x, cam.lensRadius * dsk.y);
            // ft from the main ray's (already refocused) direction, as the reference computes it
            Pt3 pFocus0 = pCamera + Vec3(0, 0, 1) * (cam.focalDistance / 1.f);
            Vec3 dMain = Normalize(pFocus0 - Pt3(pl.x, pl.y, 0));
            float ft = cam.focalDistance / dMain.z;
            Pt3 pFocus = pCamera + cam.dxCamera + (ft * Vec3(0, 0, 1));
            r.rxO = Pt3(pl.x, pl.y, 0);
            r.rxD = Normalize(pFocus - r.rxO);
            pFocus = pCamera + cam.dyCamera + (ft * Vec3(0, 0, 1));
            r.ryO = Pt3(pl.x, pl.y, 0);
            r.ryD = Normalize(pFocus - r.ryO);
        } else {
            r.rxO = pCamera + cam.dxCamera; r.ryO = pCamera + cam.dyCamera;
            r.rxD = r.ryD = Vec3(0, 0, 1);
        }
    } else if (cam.lensRadius > 0) {
        Vec2 dsk = SampleUniformDiskConcentric(cs.pLens);
        Vec2 pl(cam.lensRadius * dsk.x, cam.lensRadius * dsk.y);
        Vec3 dx = Normalize(Vec3(pCamera.x + cam.dxCamera.x, pCamera.y + cam.dxCamera.y, pCamera.z + cam.dxCamera.z));
        float ft = cam.focalDistance / dx.z;
        Pt3 pFocus = Pt3(0, 0, 0) + (dx * ft);
        r.rxO = Pt3(pl.x, pl.y, 0);
        r.rxD = Normalize(pFocus - r.rxO);
        Vec3 dy = Normalize(Vec3(pCamera.x + cam.dyCamera.x, pCamera.y + cam.dyCamera.y, pCamera.z + cam.dyCamera.z));
        ft = cam.focalDistance / dy.z;
        pFocus = Pt3(0, 0, 0) + (dy * ft);
        r.ryO = Pt3(pl.x, pl.y, 0);
        r.ryD = Normalize(pFocus - r.ryO);
    } else {
        r.rxO = r.ryO = Pt3(0, 0, 0);
        r.rxD = Normalize(Vec3(pCamera.x, pCamera.y, pCamera.z) + cam.dxCamera);
        r.ryD = Normalize(Vec3(pCamera.x, pCamera.y, pCamera.z) + cam.dyCamera);
    }
    r.rxO = XformPoint(cam.renderFromCamera, r.rxO); r.ryO = XformPoint(cam.renderFromCamera, r.ryO);
    r.rxD = XformVector(cam.renderFromCamera, r.rxD); r.ryD = XformVector(cam.renderFromCamera, r.ryD);
    if (!scaleDifferentials) return r;   // CameraBase::FindMinimumDifferentials looks at the camera's own differentials
    float s = cam.rayDiffScale;
    r.rxO = o + (r.rxO - o) * s; r.ryO = o + (r.ryO - o) * s;
    r.rxD = d + (r.rxD - d) * s; r.ryD = d + (r.ryD - d) * s;
    return r;
}


// CameraBase::FindMinimumDifferentials (cameras.cpp:154-200): the smallest position / direction differentials over 512 camera rays
// along the film diagonal, in camera space; Camera::Approximate_dp_dxy starts from them.  cameraFromRender = renderFromCamera's inverse.
inline void FindMinimumDifferentials(DeviceCamera *cam, int fullResX, int fullResY) {
    if (cam->type == 1) {   // OrthographicCamera's constructor (cameras.h:291-297)
        cam->minDirDifferentialX = cam->minDirDifferentialY = Vec3(0, 0, 0);
        cam->minPosDifferentialX = cam->dxCamera; cam->minPosDifferentialY = cam->dyCamera;
        return;
    }
    const float inf = BPT_INFINITY;
    cam->minPosDifferentialX = cam->minPosDifferentialY = cam->minDirDifferentialX = cam->minDirDifferentialY = Vec3(inf, inf, inf);
    CamSample cs{};
    cs.pLens = Vec2(0.5f, 0.5f);
    cs.time = 0.5f;
    const int n = 512;
    for (int i = 0; i < n; ++i) {
        cs.pFilm.x = float(i) / (n - 1) * fullResX;
        cs.pFilm.y = float(i) / (n - 1) * fullResY;
        Pt3 o; Vec3 d;
        GenerateCameraRay(*cam, cs.pFilm, cs.pLens, &o, &d);
        RayDiffs rd = CameraRayDifferentials(*cam, cs, o, d, false);
        Vec3 dox = XformVector(cam->cameraFromRender, rd.rxO - o);
        if (Length(dox) < Length(cam->minPosDifferentialX)) cam->minPosDifferentialX = dox;
        Vec3 doy = XformVector(cam->cameraFromRender, rd.ryO - o);
        if (Length(doy) < Length(cam->minPosDifferentialY)) cam->minPosDifferentialY = doy;
        d = Normalize(d);
        Vec3 rxD = Normalize(rd.rxD), ryD = Normalize(rd.ryD);
        Frame f = Frame::FromZ(d);
        Vec3 df = f.ToLocal(d);
        Vec3 dxf = Normalize(f.ToLocal(rxD)), dyf = Normalize(f.ToLocal(ryD));
        if (Length(dxf - df) < Length(cam->minDirDifferentialX)) cam->minDirDifferentialX = dxf - df;
        if (Length(dyf - df) < Length(cam->minDirDifferentialY)) cam->minDirDifferentialY = dyf - df;
    }
}

}  // namespace bpt
