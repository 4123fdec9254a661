// media.h -- participating media on the path, float-exact restatement of the reference's CPU semantics:
//   HGPhaseFunction / HenyeyGreenstein / SampleHenyeyGreenstein   media.h:43-71, util/scattering.h:49-58, util/sampling.cpp:348-372
//   HomogeneousMajorantIterator, DDAMajorantIterator              media.h:80-102, 136-214
//   HomogeneousMedium / GridMedium ::SamplePoint / ::SampleRay    media.h:217-262, 265-352 (emission: Le / Lescale; no temperature grids)
//   SampledGrid<Float>::Lookup (trilinear, zero outside)          util/containers.h:803-835
//   Bounds3f::IntersectP(o, d, tMax, &t0, &t1), Bounds3f::Offset  util/vecmath.h:1547-1571, 1323-1332
//   SampleT_maj                                                   media.h:736-800
// Host + device so that tests/tools can pin it against vectors produced by the reference itself.
#pragma once
#include "pmath.h"
#include "sampling.h"
#include "spectrum.h"
#include "scene.h"
#include "bxdfs.h"   // SampleExponential

namespace bpt {

enum { MEDIUM_HOMOGENEOUS = 0, MEDIUM_GRID = 1, MEDIUM_RGBGRID = 2 };

struct DeviceMedium {
    int type;
    int sigmaA, sigmaS;          // offsets (floats) of the 471-entry DenselySampledSpectrum tables, already multiplied by "scale"
    float g;
    Mat4 mInv;                   // renderFromMedium.mInv (GridMedium)
    float bmin[3], bmax[3];      // GridMedium::bounds (medium space)
    const float *density;        // nx*ny*nz, x fastest
    int nx, ny, nz;
    const float *majorant;       // mres^3 maximum densities (MajorantGrid::voxels)
    int mres[3];
    // emission (media.h:223-246 homogeneous: Le_spec scaled by Lescale / photometric norm; :296-318 grid: LeScale grid x Le_spec)
    int Le;                      // offset of the 471-entry Le_spec table, -1: the medium does not emit
    const float *leScale;        // GridMedium::LeScale (SampledGrid<Float>), lnx*lny*lnz
    int lnx, lny, lnz;
    // RGBGridMedium (media.h:355-427): per voxel an RGBUnboundedSpectrum / RGBIlluminantSpectrum as (scale, c0, c1, c2); a null grid is
    // the constant 1 for the coefficients and "no emission" for Le; `majorant` already holds sigmaScale * (max sigma_a + max sigma_s)
    const float4 *rgbA, *rgbS, *rgbLe;
    float sigmaScale, rgbLeScale;
    int illum;                   // 471-entry table of the colour space's illuminant (RGBIlluminantSpectrum::Sample)
};

// ------------------------------------------------------------------ phase function
// (HenyeyGreenstein / SampleHenyeyGreenstein live in bxdfs.h, shared with LayeredBxDF)
BPT_HD float PhaseHG_p(float g, Vec3 wo, Vec3 wi) { return HenyeyGreenstein(Dot(wo, wi), g); }

// SampleDiscrete over three weights (util/sampling.h:79-113), offset only
BPT_HD int SampleDiscrete3(float w0, float w1, float w2, float u) {
    float w[3] = {w0, w1, w2};
    float sumWeights = 0;
    for (int i = 0; i < 3; ++i) sumWeights += w[i];
    float up = u * sumWeights;
    if (up == sumWeights) up = NextFloatDown(up);
    int offset = 0;
    float sum = 0;
    while (offset < 2 && sum + w[offset] <= up) sum += w[offset++];   // (the reference DCHECKs offset < size)
    return offset;
}

// ------------------------------------------------------------------ transforms (medium space)
BPT_HD Pt3 ApplyInversePoint(const Mat4 &Minv, Pt3 p) {   // Transform::ApplyInverse(Point3f) (util/transform.h:387-398)
    const float (*m)[4] = Minv.m;
    float x = p.x, y = p.y, z = p.z;
    float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
    float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
    float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
    float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
    if (wp == 1) return Pt3(xp, yp, zp);
    return Pt3(xp, yp, zp) / wp;
}
BPT_HD Vec3 BoundsOffset(const float bmin[3], const float bmax[3], Pt3 p) {
    Vec3 o(p.x - bmin[0], p.y - bmin[1], p.z - bmin[2]);
    if (bmax[0] > bmin[0]) o.x /= bmax[0] - bmin[0];
    if (bmax[1] > bmin[1]) o.y /= bmax[1] - bmin[1];
    if (bmax[2] > bmin[2]) o.z /= bmax[2] - bmin[2];
    return o;
}
BPT_HD bool BoundsIntersectP(const float bmin[3], const float bmax[3], Pt3 o, Vec3 d, float tMax, float *hitt0, float *hitt1) {
    float t0 = 0, t1 = tMax;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int i = 0; i < 3; ++i) {
        float invRayDir = 1 / dd[i];
        float tNear = (bmin[i] - oo[i]) * invRayDir;
        float tFar = (bmax[i] - oo[i]) * invRayDir;
        if (tNear > tFar) { float t = tNear; tNear = tFar; tFar = t; }
        tFar *= 1 + 2 * gamma_n(3);
        t0 = tNear > t0 ? tNear : t0;
        t1 = tFar < t1 ? tFar : t1;
        if (t0 > t1) return false;
    }
    *hitt0 = t0; *hitt1 = t1;
    return true;
}

// ------------------------------------------------------------------ density grid
BPT_HD float GridFetch(const DeviceMedium &m, int x, int y, int z) {
    if (x < 0 || y < 0 || z < 0 || x >= m.nx || y >= m.ny || z >= m.nz) return 0.f;
    size_t i = ((size_t)z * m.ny + y) * m.nx + x;
#if defined(__CUDA_ARCH__)
    return __ldg(m.density + i);
#else
    return m.density[i];
#endif
}
BPT_HD float GridLookup(const DeviceMedium &m, Vec3 p) {
    float sx = p.x * m.nx - .5f, sy = p.y * m.ny - .5f, sz = p.z * m.nz - .5f;
    float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float dx = sx - (float)ix, dy = sy - (float)iy, dz = sz - (float)iz;
    float d00 = Lerp(dx, GridFetch(m, ix, iy, iz), GridFetch(m, ix + 1, iy, iz));
    float d10 = Lerp(dx, GridFetch(m, ix, iy + 1, iz), GridFetch(m, ix + 1, iy + 1, iz));
    float d01 = Lerp(dx, GridFetch(m, ix, iy, iz + 1), GridFetch(m, ix + 1, iy, iz + 1));
    float d11 = Lerp(dx, GridFetch(m, ix, iy + 1, iz + 1), GridFetch(m, ix + 1, iy + 1, iz + 1));
    return Lerp(dz, Lerp(dy, d00, d10), Lerp(dy, d01, d11));
}

// MajorantGrid of a GridMedium: majorantGrid.Set(x, y, z, densityGrid.MaxValue(majorantGrid.VoxelBounds(x, y, z)))
// (media.cpp:240-246, media.h:125-129, util/containers.h:837-858).  Host only (scene preparation).
inline void ComputeMajorantGrid(const float *density, int nx, int ny, int nz, int R, float *maj) {
    auto look = [&](int x, int y, int z) -> float {
        if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return 0.f;
        return density[((size_t)z * ny + y) * nx + x];
    };
    for (int z = 0; z < R; ++z) for (int y = 0; y < R; ++y) for (int x = 0; x < R; ++x) {
        float p0[3] = {float(x) / R, float(y) / R, float(z) / R}, p1[3] = {float(x + 1) / R, float(y + 1) / R, float(z + 1) / R};
        int n3[3] = {nx, ny, nz}, lo[3], hi[3];
        for (int k = 0; k < 3; ++k) {
            float a = p0[k] * n3[k] - .5f, b = p1[k] * n3[k] - .5f;
            int fa = (int)floorf(a), fb = (int)floorf(b);
            lo[k] = fa > 0 ? fa : 0;
            hi[k] = fb + 1 < n3[k] - 1 ? fb + 1 : n3[k] - 1;
        }
        float mv = look(lo[0], lo[1], lo[2]);
        for (int zz = lo[2]; zz <= hi[2]; ++zz) for (int yy = lo[1]; yy <= hi[1]; ++yy) for (int xx = lo[0]; xx <= hi[0]; ++xx) {
            float v = look(xx, yy, zz);
            mv = mv < v ? v : mv;
        }
        maj[(size_t)x + R * ((size_t)y + R * (size_t)z)] = mv;
    }
}

// ------------------------------------------------------------------ majorant iterators
struct MajorantSegment { float tMin, tMax; Spec sigma_maj; };

struct MajorantIter {
    int kind;                 // 0: exhausted / empty, 1: homogeneous (one segment), 2: DDA over the majorant grid
    Spec sigma;               // homogeneous: sigma_maj; DDA: sigma_t
    float tMin, tMax;
    const float *grid; int res[3];
    float nextCrossingT[3], deltaT[3];
    int step[3], voxelLimit[3], voxel[3];

    BPT_HD bool Next(MajorantSegment *seg) {
        if (kind == 1) {
            kind = 0;
            seg->tMin = tMin; seg->tMax = tMax; seg->sigma_maj = sigma;
            return true;
        }
        if (kind != 2) return false;
        if (tMin >= tMax) return false;
        int bits = ((nextCrossingT[0] < nextCrossingT[1]) << 2) + ((nextCrossingT[0] < nextCrossingT[2]) << 1) +
                   ((nextCrossingT[1] < nextCrossingT[2]));
        // cmpToAxis[8] = {2, 1, 2, 1, 2, 2, 0, 0}, packed 2 bits per entry
        int stepAxis = (int)((0x0A66u >> (2 * bits)) & 3u);
        float nct = stepAxis == 0 ? nextCrossingT[0] : (stepAxis == 1 ? nextCrossingT[1] : nextCrossingT[2]);
        float tVoxelExit = stdmin(tMax, nct);
        size_t vi = (size_t)voxel[0] + (size_t)res[0] * ((size_t)voxel[1] + (size_t)res[1] * (size_t)voxel[2]);
#if defined(__CUDA_ARCH__)
        float maxDensity = __ldg(grid + vi);
#else
        float maxDensity = grid[vi];
#endif
        seg->sigma_maj = sigma * maxDensity;
        seg->tMin = tMin; seg->tMax = tVoxelExit;
        tMin = tVoxelExit;
        if (nct > tMax) tMin = tMax;
        // (component updates written without dynamic indexing so that the arrays stay in registers)
        if (stepAxis == 0) { voxel[0] += step[0]; if (voxel[0] == voxelLimit[0]) tMin = tMax; nextCrossingT[0] += deltaT[0]; }
        else if (stepAxis == 1) { voxel[1] += step[1]; if (voxel[1] == voxelLimit[1]) tMin = tMax; nextCrossingT[1] += deltaT[1]; }
        else { voxel[2] += step[2]; if (voxel[2] == voxelLimit[2]) tMin = tMax; nextCrossingT[2] += deltaT[2]; }
        return true;
    }
};

// InstRay ApplyInverseRay(Minv, o, d) is Transform::ApplyInverse(Ray, &tMax) (scene.h)
BPT_HD MajorantIter MediumSampleRay(const DeviceMedium &m, const float *specPool, Pt3 o, Vec3 d, float raytMax, const Wavelengths &wl) {
    MajorantIter it;
    it.kind = 0;
    Spec sigma_a = DenseSample(specPool + m.sigmaA, wl), sigma_s = DenseSample(specPool + m.sigmaS, wl);
    if (m.type == MEDIUM_HOMOGENEOUS) {
        it.kind = 1; it.tMin = 0; it.tMax = raytMax; it.sigma = sigma_a + sigma_s;
        return it;
    }
    InstRay r = ApplyInverseRay(m.mInv, o, d);
    raytMax -= r.dt;
    float tMin, tMax;
    if (!BoundsIntersectP(m.bmin, m.bmax, r.o, r.d, raytMax, &tMin, &tMax)) return it;
    it.kind = 2;
    it.sigma = m.type == MEDIUM_RGBGRID ? Spec(1.f) : sigma_a + sigma_s;   // RGBGridMedium::SampleRay: sigma_t = 1 (media.h:413)
    it.tMin = tMin; it.tMax = tMax;
    it.grid = m.majorant;
    // DDAMajorantIterator constructor
    Vec3 diag(m.bmax[0] - m.bmin[0], m.bmax[1] - m.bmin[1], m.bmax[2] - m.bmin[2]);
    Vec3 og = BoundsOffset(m.bmin, m.bmax, r.o);
    float rd[3] = {r.d.x / diag.x, r.d.y / diag.y, r.d.z / diag.z};
    float gi[3] = {og.x + rd[0] * tMin, og.y + rd[1] * tMin, og.z + rd[2] * tMin};
    for (int axis = 0; axis < 3; ++axis) {
        int res = m.mres[axis];
        it.res[axis] = res;
        float v = gi[axis] * res;
        float lo = 0, hi = (float)(res - 1);
        v = v < lo ? lo : (v > hi ? hi : v);
        it.voxel[axis] = (int)v;
        it.deltaT[axis] = 1 / (fabsf(rd[axis]) * res);
        if (rd[axis] == -0.f) rd[axis] = 0.f;
        if (rd[axis] >= 0) {
            float nextVoxelPos = (float)(it.voxel[axis] + 1) / res;
            it.nextCrossingT[axis] = tMin + (nextVoxelPos - gi[axis]) / rd[axis];
            it.step[axis] = 1;
            it.voxelLimit[axis] = res;
        } else {
            float nextVoxelPos = (float)(it.voxel[axis]) / res;
            it.nextCrossingT[axis] = tMin + (nextVoxelPos - gi[axis]) / rd[axis];
            it.step[axis] = -1;
            it.voxelLimit[axis] = -1;
        }
    }
    return it;
}

struct MediumProps { Spec sigma_a, sigma_s, Le; };
// SampledGrid<Float>::Lookup (util/containers.h:803-835) over an arbitrary grid: trilinear, zero outside
BPT_HD float ScalarGridLookup(const float *g, int nx, int ny, int nz, Vec3 p) {
    auto fetch = [&](int x, int y, int z) -> float {
        if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return 0.f;
        size_t i = ((size_t)z * ny + y) * nx + x;
#if defined(__CUDA_ARCH__)
        return __ldg(g + i);
#else
        return g[i];
#endif
    };
    float sx = p.x * nx - .5f, sy = p.y * ny - .5f, sz = p.z * nz - .5f;
    float fx = floorf(sx), fy = floorf(sy), fz = floorf(sz);
    int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    float dx = sx - (float)ix, dy = sy - (float)iy, dz = sz - (float)iz;
    float d00 = Lerp(dx, fetch(ix, iy, iz), fetch(ix + 1, iy, iz));
    float d10 = Lerp(dx, fetch(ix, iy + 1, iz), fetch(ix + 1, iy + 1, iz));
    float d01 = Lerp(dx, fetch(ix, iy, iz + 1), fetch(ix + 1, iy, iz + 1));
    float d11 = Lerp(dx, fetch(ix, iy + 1, iz + 1), fetch(ix + 1, iy + 1, iz + 1));
    return Lerp(dz, Lerp(dy, d00, d10), Lerp(dy, d01, d11));
}
// SampledGrid<RGBUnboundedSpectrum / RGBIlluminantSpectrum>::Lookup(p, convert) (util/containers.h:803-835): trilinear over the
// spectra sampled at the wavelengths; a voxel outside the grid converts a default-constructed spectrum, i.e. zero
BPT_HD Spec RGBGridLookup(const float4 *g, int nx, int ny, int nz, Vec3 p, const Wavelengths &wl, const float *illum) {
    auto fetch = [&](int x, int y, int z) -> Spec {
        if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return Spec(0.f);
        size_t i = ((size_t)z * ny + y) * nx + x;
#if defined(__CUDA_ARCH__)
        float4 v = __ldg(g + i);
#else
        float4 v = g[i];
#endif
        Spec s;
        for (int k = 0; k < 4; ++k) s[k] = v.x * SigmoidPoly(v.y, v.z, v.w, wl.lambda[k]);
        if (illum) s = s * DenseSample(illum, wl);
        return s;
    };
    auto lerp = [](float t, const Spec &a, const Spec &b) { return (1 - t) * a + t * b; };
    float sx = p.x * nx - .5f, sy = p.y * ny - .5f, sz = p.z * nz - .5f;
    int ix = (int)floorf(sx), iy = (int)floorf(sy), iz = (int)floorf(sz);
    float dx = sx - (float)ix, dy = sy - (float)iy, dz = sz - (float)iz;
    Spec d00 = lerp(dx, fetch(ix, iy, iz), fetch(ix + 1, iy, iz));
    Spec d10 = lerp(dx, fetch(ix, iy + 1, iz), fetch(ix + 1, iy + 1, iz));
    Spec d01 = lerp(dx, fetch(ix, iy, iz + 1), fetch(ix + 1, iy, iz + 1));
    Spec d11 = lerp(dx, fetch(ix, iy + 1, iz + 1), fetch(ix + 1, iy + 1, iz + 1));
    return lerp(dz, lerp(dy, d00, d10), lerp(dy, d01, d11));
}
BPT_HD MediumProps MediumSamplePoint(const DeviceMedium &m, const float *specPool, Pt3 p, const Wavelengths &wl) {
    MediumProps mp;
    if (m.type == MEDIUM_RGBGRID) {   // RGBGridMedium::SamplePoint (media.h:377-401)
        Pt3 pm = ApplyInversePoint(m.mInv, p);
        Vec3 po = BoundsOffset(m.bmin, m.bmax, pm);
        mp.sigma_a = m.sigmaScale * (m.rgbA ? RGBGridLookup(m.rgbA, m.nx, m.ny, m.nz, po, wl, nullptr) : Spec(1.f));
        mp.sigma_s = m.sigmaScale * (m.rgbS ? RGBGridLookup(m.rgbS, m.nx, m.ny, m.nz, po, wl, nullptr) : Spec(1.f));
        mp.Le = Spec(0.f);
        if (m.rgbLe && m.rgbLeScale > 0) mp.Le = m.rgbLeScale * RGBGridLookup(m.rgbLe, m.nx, m.ny, m.nz, po, wl, specPool + m.illum);
        return mp;
    }
    mp.sigma_a = DenseSample(specPool + m.sigmaA, wl);
    mp.sigma_s = DenseSample(specPool + m.sigmaS, wl);
    mp.Le = Spec(0.f);
    if (m.type == MEDIUM_GRID) {
        Pt3 pm = ApplyInversePoint(m.mInv, p);
        Vec3 po = BoundsOffset(m.bmin, m.bmax, pm);
        float d = GridLookup(m, po);
        mp.sigma_a *= d;
        mp.sigma_s *= d;
        if (m.Le >= 0) {   // GridMedium::SamplePoint: scale = LeScale.Lookup(p); if (scale > 0) Le = scale * Le_spec.Sample(lambda)
            float scale = ScalarGridLookup(m.leScale, m.lnx, m.lny, m.lnz, po);
            if (scale > 0) mp.Le = DenseSample(specPool + m.Le, wl) * scale;
        }
    } else if (m.Le >= 0) mp.Le = DenseSample(specPool + m.Le, wl);   // HomogeneousMedium: Le_spec already scaled
    return mp;
}

// SampleT_maj (media.h:736-800) as a resumable tracker: each Step() either fetches the next majorant segment or makes one
// tentative collision, so that a warp can advance 32 independent rays in lock-step and refill the lanes that finish
// (k_sample_medium).  Same operations in the same order as the reference's doubly nested loop.
struct TMajTracker {
    enum { CONTINUE = 0, EVENT = 1, DONE = 2 };
    MajorantIter iter;
    MajorantSegment seg;
    Spec T_maj;
    Pt3 o; Vec3 d;          // ray with normalised direction
    float u, tMin, tEvent;
    bool haveSeg;
    BPT_HD void Init(const DeviceMedium &m, const float *specPool, Pt3 o_, Vec3 d_, float tMax, float u_, const Wavelengths &wl) {
        tMax *= Length(d_);
        d = Normalize(d_);
        o = o_;
        iter = MediumSampleRay(m, specPool, o, d, tMax, wl);
        T_maj = Spec(1.f);
        u = u_;
        haveSeg = false;
        tMin = tEvent = 0;
    }
    // EVENT: a tentative collision at *p (= ray(t)); T_maj holds the majorant transmittance since the previous event and
    // seg.sigma_maj the majorant.  The caller evaluates the medium there and then calls AfterEvent() to go on, or stops.
    // DONE: no more segments; T_maj is SampleT_maj's return value.
    BPT_HD int Step(RNG &rng, Pt3 *p) {
        if (!haveSeg) {
            if (!iter.Next(&seg)) return DONE;
            if (seg.sigma_maj[0] == 0) {
                float dt = seg.tMax - seg.tMin;
                if (IsInf_(dt)) dt = 3.402823466e+38f;
                T_maj *= FastExpSpec(seg.sigma_maj * (-dt));
                return CONTINUE;
            }
            tMin = seg.tMin;
            haveSeg = true;
            return CONTINUE;
        }
        float t = tMin + SampleExponential(u, seg.sigma_maj[0]);
        u = rng.UniformFloat();
        if (t < seg.tMax) {
            T_maj *= FastExpSpec(seg.sigma_maj * (-(t - tMin)));
            tEvent = t;
            *p = o + d * t;
            return EVENT;
        }
        float dt = seg.tMax - tMin;
        if (IsInf_(dt)) dt = 3.402823466e+38f;
        T_maj *= FastExpSpec(seg.sigma_maj * (-dt));
        haveSeg = false;
        return CONTINUE;
    }
    BPT_HD void AfterEvent() { T_maj = Spec(1.f); tMin = tEvent; }
};

// SampleT_maj (media.h:736-800).  callback(p, mp, sigma_maj, T_maj) -> bool (false stops the tracking and makes it return 1)
template <typename F>
BPT_HD Spec SampleT_maj(const DeviceMedium &m, const float *specPool, Pt3 o, Vec3 d, float tMax, float u, RNG &rng,
                        const Wavelengths &wl, F callback) {
    TMajTracker trk;
    trk.Init(m, specPool, o, d, tMax, u, wl);
    while (true) {
        Pt3 p;
        int st = trk.Step(rng, &p);
        if (st == TMajTracker::DONE) return trk.T_maj;
        if (st == TMajTracker::EVENT) {
            MediumProps mp = MediumSamplePoint(m, specPool, p, wl);
            if (!callback(p, mp, trk.seg.sigma_maj, trk.T_maj)) return Spec(1.f);
            trk.AfterEvent();
        }
    }
}

}  // namespace bpt
