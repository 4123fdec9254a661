// xform.h -- host-side 4x4 Transform with tracked inverse, restating util/transform.{h,cpp} and the matrix
// helpers of util/math.h (Inverse via Laplace expansion with InnerProduct :1571-1633, operator* :1476-1485,
// Determinant<3> :1419-1425) so that scene set-up produces the same floats as the reference's.
#pragma once
#include <cmath>
#include <cstring>
#include "pmath.h"

namespace bpt {

struct M4 {
    float m[4][4];
    M4() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = (i == j) ? 1.f : 0.f; }
    M4(float a00, float a01, float a02, float a03, float a10, float a11, float a12, float a13, float a20, float a21, float a22,
       float a23, float a30, float a31, float a32, float a33) {
        float t[16] = {a00, a01, a02, a03, a10, a11, a12, a13, a20, a21, a22, a23, a30, a31, a32, a33};
        std::memcpy(m, t, sizeof(t));
    }
    const float *operator[](int i) const { return m[i]; }
    float *operator[](int i) { return m[i]; }
    bool IsIdentity() const { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (m[i][j] != ((i == j) ? 1.f : 0.f)) return false; return true; }
    bool operator==(const M4 &o) const { return std::memcmp(m, o.m, sizeof(m)) == 0; }
};

// SquareMatrix<4> * SquareMatrix<4> as the reference EXECUTES it: the generic template (util/math.h:1497-1508), a chain of FMAs
// starting from 0.  The InnerProduct() version at util/math.h:1475-1484 is written as `template <>` BEFORE that primary template is
// declared, so it specialises the matrix * vector template of :1470-1473 instead and overload resolution never selects it
// (checked against oracle/_ref/lib/libpbrt.a: Translate * Scale * Translate differs in the last bit between the two, and the
// library gives the FMA-chain bits; tests/test_host_cpu.py pins it through the ABI recorder).
inline M4 Mul(const M4 &m1, const M4 &m2) {
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float acc = 0;
            for (int k = 0; k < 4; ++k) acc = FMA(m1[i][k], m2[k][j], acc);
            r[i][j] = acc;
        }
    return r;
}
inline M4 Transpose(const M4 &a) { M4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[i][j] = a[j][i]; return r; }

inline bool Inverse(const M4 &m, M4 *out) {
    float s0 = DifferenceOfProducts(m[0][0], m[1][1], m[1][0], m[0][1]);
    float s1 = DifferenceOfProducts(m[0][0], m[1][2], m[1][0], m[0][2]);
    float s2 = DifferenceOfProducts(m[0][0], m[1][3], m[1][0], m[0][3]);
    float s3 = DifferenceOfProducts(m[0][1], m[1][2], m[1][1], m[0][2]);
    float s4 = DifferenceOfProducts(m[0][1], m[1][3], m[1][1], m[0][3]);
    float s5 = DifferenceOfProducts(m[0][2], m[1][3], m[1][2], m[0][3]);
    float c0 = DifferenceOfProducts(m[2][0], m[3][1], m[3][0], m[2][1]);
    float c1 = DifferenceOfProducts(m[2][0], m[3][2], m[3][0], m[2][2]);
    float c2 = DifferenceOfProducts(m[2][0], m[3][3], m[3][0], m[2][3]);
    float c3 = DifferenceOfProducts(m[2][1], m[3][2], m[3][1], m[2][2]);
    float c4 = DifferenceOfProducts(m[2][1], m[3][3], m[3][1], m[2][3]);
    float c5 = DifferenceOfProducts(m[2][2], m[3][3], m[3][2], m[2][3]);
    float determinant = InnerProduct(s0, c5, -s1, c4, s2, c3, s3, c2, s5, c0, -s4, c1);
    if (determinant == 0) return false;
    float s = 1 / determinant;
    float inv[4][4] = {{s * InnerProduct(m[1][1], c5, m[1][3], c3, -m[1][2], c4), s * InnerProduct(-m[0][1], c5, m[0][2], c4, -m[0][3], c3),
                        s * InnerProduct(m[3][1], s5, m[3][3], s3, -m[3][2], s4), s * InnerProduct(-m[2][1], s5, m[2][2], s4, -m[2][3], s3)},
                       {s * InnerProduct(-m[1][0], c5, m[1][2], c2, -m[1][3], c1), s * InnerProduct(m[0][0], c5, m[0][3], c1, -m[0][2], c2),
                        s * InnerProduct(-m[3][0], s5, m[3][2], s2, -m[3][3], s1), s * InnerProduct(m[2][0], s5, m[2][3], s1, -m[2][2], s2)},
                       {s * InnerProduct(m[1][0], c4, m[1][3], c0, -m[1][1], c2), s * InnerProduct(-m[0][0], c4, m[0][1], c2, -m[0][3], c0),
                        s * InnerProduct(m[3][0], s4, m[3][3], s0, -m[3][1], s2), s * InnerProduct(-m[2][0], s4, m[2][1], s2, -m[2][3], s0)},
                       {s * InnerProduct(-m[1][0], c3, m[1][1], c1, -m[1][2], c0), s * InnerProduct(m[0][0], c3, m[0][2], c0, -m[0][1], c1),
                        s * InnerProduct(-m[3][0], s3, m[3][1], s1, -m[3][2], s0), s * InnerProduct(m[2][0], s3, m[2][2], s0, -m[2][1], s1)}};
    std::memcpy(out->m, inv, sizeof(inv));
    return true;
}

struct Xform {
    M4 m, mInv;
    Xform() {}
    explicit Xform(const M4 &mm) : m(mm) {
        if (!Inverse(mm, &mInv)) { float nan = std::nanf(""); for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) mInv[i][j] = nan; }
    }
    Xform(const M4 &mm, const M4 &mi) : m(mm), mInv(mi) {}
    Xform operator*(const Xform &t2) const { return Xform(Mul(m, t2.m), Mul(t2.mInv, mInv)); }
    bool IsIdentity() const { return m.IsIdentity(); }
    Pt3 operator()(Pt3 p) const {
        float xp = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
        float yp = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
        float zp = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
        float wp = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
        if (wp == 1) return Pt3(xp, yp, zp);
        return Pt3(xp, yp, zp) / wp;
    }
    Vec3 operator()(Vec3 v) const {
        return Vec3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
                    m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
    }
    Nrm3 operator()(Nrm3 n) const {
        float x = n.x, y = n.y, z = n.z;
        return Nrm3(mInv[0][0] * x + mInv[1][0] * y + mInv[2][0] * z, mInv[0][1] * x + mInv[1][1] * y + mInv[2][1] * z,
                    mInv[0][2] * x + mInv[1][2] * y + mInv[2][2] * z);
    }
    bool SwapsHandedness() const {
        float minor12 = DifferenceOfProducts(m[1][1], m[2][2], m[1][2], m[2][1]);
        float minor02 = DifferenceOfProducts(m[1][0], m[2][2], m[1][2], m[2][0]);
        float minor01 = DifferenceOfProducts(m[1][0], m[2][1], m[1][1], m[2][0]);
        float det = FMA(m[0][2], minor01, DifferenceOfProducts(m[0][0], minor12, m[0][1], minor02));
        return det < 0;
    }
};
inline Xform Inverse(const Xform &t) { return Xform(t.mInv, t.m); }

inline Xform Translate(Vec3 d) {
    return Xform(M4(1, 0, 0, d.x, 0, 1, 0, d.y, 0, 0, 1, d.z, 0, 0, 0, 1), M4(1, 0, 0, -d.x, 0, 1, 0, -d.y, 0, 0, 1, -d.z, 0, 0, 0, 1));
}
inline Xform Scale(float x, float y, float z) {
    return Xform(M4(x, 0, 0, 0, 0, y, 0, 0, 0, 0, z, 0, 0, 0, 0, 1), M4(1 / x, 0, 0, 0, 0, 1 / y, 0, 0, 0, 0, 1 / z, 0, 0, 0, 0, 1));
}
inline Xform RotateSC(float sinTheta, float cosTheta, Vec3 axis) {
    Vec3 a = Normalize(axis);
    M4 m;
    m[0][0] = a.x * a.x + (1 - a.x * a.x) * cosTheta;
    m[0][1] = a.x * a.y * (1 - cosTheta) - a.z * sinTheta;
    m[0][2] = a.x * a.z * (1 - cosTheta) + a.y * sinTheta;
    m[0][3] = 0;
    m[1][0] = a.x * a.y * (1 - cosTheta) + a.z * sinTheta;
    m[1][1] = a.y * a.y + (1 - a.y * a.y) * cosTheta;
    m[1][2] = a.y * a.z * (1 - cosTheta) - a.x * sinTheta;
    m[1][3] = 0;
    m[2][0] = a.x * a.z * (1 - cosTheta) - a.y * sinTheta;
    m[2][1] = a.y * a.z * (1 - cosTheta) + a.x * sinTheta;
    m[2][2] = a.z * a.z + (1 - a.z * a.z) * cosTheta;
    m[2][3] = 0;
    return Xform(m, Transpose(m));
}
inline Xform Rotate(float thetaDeg, Vec3 axis) {
    float sinTheta = std::sin(Radians(thetaDeg)), cosTheta = std::cos(Radians(thetaDeg));
    return RotateSC(sinTheta, cosTheta, axis);
}
inline Xform LookAt(Pt3 pos, Pt3 look, Vec3 up, bool *ok) {
    M4 wfc;
    wfc[0][3] = pos.x; wfc[1][3] = pos.y; wfc[2][3] = pos.z; wfc[3][3] = 1;
    Vec3 dir = Normalize(look - pos);
    *ok = Length(Cross(Normalize(up), dir)) != 0;
    Vec3 right = Normalize(Cross(Normalize(up), dir));
    Vec3 newUp = Cross(dir, right);
    wfc[0][0] = right.x; wfc[1][0] = right.y; wfc[2][0] = right.z; wfc[3][0] = 0.;
    wfc[0][1] = newUp.x; wfc[1][1] = newUp.y; wfc[2][1] = newUp.z; wfc[3][1] = 0.;
    wfc[0][2] = dir.x; wfc[1][2] = dir.y; wfc[2][2] = dir.z; wfc[3][2] = 0.;
    M4 cfw;
    Inverse(wfc, &cfw);
    return Xform(cfw, wfc);
}
inline Xform Perspective(float fov, float n, float f) {
    M4 persp(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, f / (f - n), -f * n / (f - n), 0, 0, 1, 0);
    float invTanAng = 1 / std::tan(Radians(fov) / 2);
    return Scale(invTanAng, invTanAng, 1) * Xform(persp);
}

struct Bounds3 {
    Pt3 pMin, pMax;
    Bounds3() : pMin(BPT_INFINITY, BPT_INFINITY, BPT_INFINITY), pMax(-BPT_INFINITY, -BPT_INFINITY, -BPT_INFINITY) {}
    Bounds3(Pt3 a, Pt3 b) : pMin(stdmin(a.x, b.x), stdmin(a.y, b.y), stdmin(a.z, b.z)), pMax(stdmax(a.x, b.x), stdmax(a.y, b.y), stdmax(a.z, b.z)) {}
    Vec3 Diagonal() const { return pMax - pMin; }
    float SurfaceArea() const { Vec3 d = Diagonal(); return 2 * (d.x * d.y + d.x * d.z + d.y * d.z); }
    Vec3 Offset(Pt3 p) const {
        Vec3 o = p - pMin;
        if (pMax.x > pMin.x) o.x /= pMax.x - pMin.x;
        if (pMax.y > pMin.y) o.y /= pMax.y - pMin.y;
        if (pMax.z > pMin.z) o.z /= pMax.z - pMin.z;
        return o;
    }
};
inline Bounds3 Union(const Bounds3 &b, Pt3 p) {
    Bounds3 r;
    r.pMin = Pt3(stdmin(b.pMin.x, p.x), stdmin(b.pMin.y, p.y), stdmin(b.pMin.z, p.z));
    r.pMax = Pt3(stdmax(b.pMax.x, p.x), stdmax(b.pMax.y, p.y), stdmax(b.pMax.z, p.z));
    return r;
}
inline Bounds3 Union(const Bounds3 &a, const Bounds3 &b) {
    Bounds3 r;
    r.pMin = Pt3(stdmin(a.pMin.x, b.pMin.x), stdmin(a.pMin.y, b.pMin.y), stdmin(a.pMin.z, b.pMin.z));
    r.pMax = Pt3(stdmax(a.pMax.x, b.pMax.x), stdmax(a.pMax.y, b.pMax.y), stdmax(a.pMax.z, b.pMax.z));
    return r;
}

}  // namespace bpt
