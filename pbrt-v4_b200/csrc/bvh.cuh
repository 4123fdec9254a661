// bvh.cuh -- 8-wide compressed BVH (80-byte nodes, quantised child boxes, octant-ordered slots) and its
// closest-hit / any-hit traversal.  No reference counterpart: the reference GPU path calls optixTrace
// (gpu/optix/optix.cu:65) and its CPU path walks a binary 32-byte-node BVH (cpu/aggregates.cpp:529-624).
// The only reference semantics kept here are the hit acceptance rules: the leaf test is the watertight
// IntersectTriangle (shapes.cpp:168-269) with a shrinking tMax, so the closest hit (t, b0, b1, b2) is the one
// BVHAggregate::Intersect would report (up to exact-t ties).
//
// Node layout (5 x 16 B, loaded as five 128-bit loads):
//   n0 = { p.x, p.y, p.z, ex | ey<<8 | ez<<16 | imask<<24 }      origin + biased exponents + internal mask
//   n1 = { childBase, triBase, meta[0..3], meta[4..7] }
//   n2 = { qlo.x[0..3], qlo.x[4..7], qhi.x[0..3], qhi.x[4..7] }
//   n3 = { qlo.y ... , qhi.y ... }      n4 = { qlo.z ... , qhi.z ... }
//   meta[i]: empty = 0; internal = 0b001_sssss with sssss = 24 + slot; leaf = unary tri count (1,3,7) << 5 | first
//   triangle offset relative to triBase (<= 23).
#pragma once
#include <cuda_runtime.h>
#include "pmath.h"
#include "shapes.h"
#include "quadrics.h"
#include "bilinear.h"
#ifndef BPT_FOLD_I2F
#define BPT_FOLD_I2F 0   // measured neutral (XU relief, but +2 % node visits): profiles/r01d_fold_prefetch_sweep.log
#endif

namespace bpt {

struct DeviceInstance;
struct BVHView {
    const float4 *nodes;      // 5 float4 per wide node; node 0 = root (instanced scenes: the top-level BVH over instances)
    const float4 *tris;       // 3 float4 per triangle in leaf order (p0.w = original triangle id; top level: instance id)
    const DeviceInstance *instances;   // non-null: two-level scene
    const float *meshAlpha;            // per mesh id: constant alpha of alpha-tested meshes (class bit 30 of the leaf record)
    const DeviceShape *shapes;         // quadrics (class bit 29 of the leaf record: the record holds the shape's bounds, not vertices)
    const float4 *patches;             // bilinear patches, 4 float4 each (same class bit)
    const int4 *meshShape;             // per mesh id: { quadric index or -1, first patch or -1, pool index of the mesh's first record relative to
                                       //   its BVH's first record (leaf records carry that relative index in p0.w), - }
};

#define BPT_STACK_SIZE 40

struct RayPre {              // per-ray constants of the watertight test
    Pt3 o; Vec3 d;
    int kx, ky, kz;
    float Sx, Sy, Sz;
};
BPT_D RayPre MakeRayPre(Pt3 o, Vec3 dir) {
    RayPre r;
    r.o = o; r.d = dir;
    r.kz = MaxComponentIndex(Abs(dir));
    r.kx = r.kz + 1; if (r.kx == 3) r.kx = 0;
    r.ky = r.kx + 1; if (r.ky == 3) r.ky = 0;
    Vec3 d = Permute(dir, r.kx, r.ky, r.kz);
    r.Sx = -d.x / d.z; r.Sy = -d.y / d.z; r.Sz = 1 / d.z;
    return r;
}
BPT_D Vec3 PermuteK(Vec3 v, int kz) {
    // (kx,ky,kz) is a cyclic rotation: kz=2 -> (x,y,z), kz=0 -> (y,z,x), kz=1 -> (z,x,y)
    return kz == 2 ? v : (kz == 0 ? Vec3(v.y, v.z, v.x) : Vec3(v.z, v.x, v.y));
}
// First stage of the watertight test (translate, permute, shear, edge functions incl. the FP64 fallback, sign and det tests):
// identical arithmetic to the opening of IntersectTrianglePre below, so `false` here means the full test returns false.
// ~85 % of the candidates end here; the survivors are parked (one per lane) and finished together with the full test, because
// running the long tail of the test (range checks, division, error bounds) right away would execute it with ~2 active lanes.
BPT_D bool IntersectTriangleHead(const RayPre &r, Pt3 p0, Pt3 p1, Pt3 p2) {
    Vec3 p0t = PermuteK(p0 - r.o, r.kz), p1t = PermuteK(p1 - r.o, r.kz), p2t = PermuteK(p2 - r.o, r.kz);
    p0t.x += r.Sx * p0t.z; p0t.y += r.Sy * p0t.z;
    p1t.x += r.Sx * p1t.z; p1t.y += r.Sy * p1t.z;
    p2t.x += r.Sx * p2t.z; p2t.y += r.Sy * p2t.z;
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
    return det != 0;
}
// Same arithmetic as IntersectTriangle (shapes.h) with the per-ray terms hoisted.
// The degenerate-triangle early-out of the reference (LengthSquared(Cross(p2 - p0, p1 - p0)) == 0) depends only on the
// triangle; the builder evaluates it once and stores the verdict in bit 31 of the leaf record's class word.
BPT_D bool IntersectTrianglePre(const RayPre &r, float tMax, Pt3 p0, Pt3 p1, Pt3 p2, TriHit *hit) {
    Vec3 p0t = PermuteK(p0 - r.o, r.kz), p1t = PermuteK(p1 - r.o, r.kz), p2t = PermuteK(p2 - r.o, r.kz);
    p0t.x += r.Sx * p0t.z; p0t.y += r.Sy * p0t.z;
    p1t.x += r.Sx * p1t.z; p1t.y += r.Sy * p1t.z;
    p2t.x += r.Sx * p2t.z; p2t.y += r.Sy * p2t.z;
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
    p0t.z *= r.Sz; p1t.z *= r.Sz; p2t.z *= r.Sz;
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

// GeometricPrimitive::Intersect's stochastic alpha test (cpu/primitive.cpp:56-75): the hit is ignored when HashFloat(ray.o, ray.d) > alpha
// -- a function of the ray the primitive sees (the instance-space ray inside an instance) and of the mesh, so a ray either passes
// through every triangle of an alpha mesh or hits them.  The reference then re-intersects the SAME shape from behind the hit
// (`Intersect(rNext, ...)` on this primitive), which a planar triangle cannot satisfy: the triangle simply does not count.
constexpr unsigned kClsAlpha = 0x40000000u;
// The hash is a real (outlined) function: scenes without alpha-tested meshes only pay the class-bit test in the traversal loops
// (inlined, it cost the closest-hit kernel 4 % through registers, r02l vs r02k).
static __device__ __noinline__ bool AlphaHashRejects(const float *meshAlpha, int mesh, float ox, float oy, float oz, float dx, float dy, float dz) {
    float a = __ldg(meshAlpha + mesh);
    if (a >= 1) return false;
    float u = 1.f;
    if (a > 0) {
        uint64_t w[3] = {(uint64_t)f2b(ox) | ((uint64_t)f2b(oy) << 32), (uint64_t)f2b(oz) | ((uint64_t)f2b(dx) << 32), (uint64_t)f2b(dy) | ((uint64_t)f2b(dz) << 32)};
        u = (float)(uint32_t)MurmurHash64A(w, 24, 0) * 0x1p-32f;
    }
    return u > a;
}
BPT_D bool AlphaRejects(const BVHView &bvh, const RayPre &r, unsigned cls, float meshIdBits) {
    if (!(cls & kClsAlpha)) return false;
    return AlphaHashRejects(bvh.meshAlpha, __float_as_int(meshIdBits), r.o.x, r.o.y, r.o.z, r.d.x, r.d.y, r.d.z);
}
// Shape leaves (Sphere / Disk / Cylinder / BilinearPatch as custom primitives, gpu/optix/optix.cu:301-469): the leaf record carries the shape's
// bounding box where a triangle has its vertices, is flagged degenerate (so the triangle test never runs on it) and has class bit 29
// set.  The test is Shape::IntersectP / BasicIntersect with the current tMax; a hit is reported as (pObj, tHit), from which the
// shading kernels rebuild the QuadricIntersection.  Outlined: scenes without quadrics only pay the class-bit test on degenerate records.
constexpr unsigned kClsShape = 0x20000000u;
static __device__ __noinline__ float4 IntersectShapeLeaf(const DeviceShape *shapes, const float4 *patches, const int4 *meshShape, int mesh, int orig, float ox, float oy,
                                                         float oz, float dx, float dy, float dz, float tMax) {
    const int4 ms = __ldg(meshShape + mesh);
    if (ms.x >= 0) {
        QuadricHit qh;
        if (!QuadricBasicIntersect(shapes[ms.x], Pt3(ox, oy, oz), Vec3(dx, dy, dz), tMax, &qh)) return make_float4(0.f, 0.f, 0.f, -1.f);
        return make_float4(qh.pObj.x, qh.pObj.y, qh.pObj.z, qh.t);
    }
    // bilinear patch (gpu/optix/optix.cu:379-469): hit = (u, v)
    const float4 *q = patches + 4 * (size_t)(ms.y + (orig - ms.z));
    float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    float u, v, t;
    if (!IntersectBilinearPatch(Pt3(ox, oy, oz), Vec3(dx, dy, dz), tMax, Pt3(a.x, a.y, a.z), Pt3(b.x, b.y, b.z), Pt3(c.x, c.y, c.z), Pt3(d.x, d.y, d.z), &u, &v, &t))
        return make_float4(0.f, 0.f, 0.f, -1.f);
    return make_float4(u, v, 0.f, t);
}
// One leaf record against the ray: the reference's Primitive::Intersect(ray, tMax) for whatever the record holds
// SHAPES = false compiles the triangle-only test (scenes without quadrics run kernels that do not contain the call at all: the
// outlined call costs the closest-hit kernel registers around the call site, wavefront_shapes.cu)
template <bool SHAPES>
BPT_D bool LeafIntersect(const BVHView &bvh, const RayPre &rp, float tMax, float4 a, float4 b, float4 c, TriHit *th) {
    const unsigned cls = __float_as_uint(b.w);
    if (cls & 0x80000000u) {
        if (!SHAPES || !(cls & kClsShape)) return false;   // degenerate triangle (shapes.cpp:172)
        float4 q = IntersectShapeLeaf(bvh.shapes, bvh.patches, bvh.meshShape, __float_as_int(c.w), __float_as_int(a.w), rp.o.x, rp.o.y, rp.o.z, rp.d.x, rp.d.y, rp.d.z, tMax);
        if (q.w < 0) return false;
        th->b0 = q.x; th->b1 = q.y; th->b2 = q.z; th->t = q.w;
        return true;
    }
    return IntersectTrianglePre(rp, tMax, Pt3(a.x, a.y, a.z), Pt3(b.x, b.y, b.z), Pt3(c.x, c.y, c.z), th) && !AlphaRejects(bvh, rp, cls, c.w);
}
BPT_D unsigned ExtractByte(unsigned x, unsigned i) { return (x >> (i * 8)) & 0xff; }
BPT_D unsigned SignExtendS8x4(unsigned x) {
    unsigned r;
    asm("prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"(r) : "r"(x));
    return r;
}
BPT_D float SafeInvDir(float d) { return 1.0f / (fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d)); }

BPT_D void PrefetchL1(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#endif
}
struct HitRec { int tri; float b0, b1, b2, t; int inst; };   // tri = leaf-order triangle index (of instance `inst`), -1 = miss

// Conservative slab test of the 8 quantised children; returns the hit mask (internal children in bits 24..31
// in octant-permuted order, triangles in bits 0..23).  The box test uses fast fused arithmetic: it only has to
// be conservative (every box whose triangle could be accepted is visited), which the per-axis error margin
// guarantees; it does not influence the reported hit.
BPT_D unsigned NodeIntersect(Vec3 o, Vec3 idir, unsigned octinv4, float tMax, float4 n0, float4 n1, float4 n2, float4 n3,
                             float4 n4) {
    unsigned e_imask = __float_as_uint(n0.w);
    float ex = __uint_as_float(ExtractByte(e_imask, 0) << 23);
    float ey = __uint_as_float(ExtractByte(e_imask, 1) << 23);
    float ez = __uint_as_float(ExtractByte(e_imask, 2) << 23);
    float ax = ex * idir.x, ay = ey * idir.y, az = ez * idir.z;
    float ox = (n0.x - o.x) * idir.x, oy = (n0.y - o.y) * idir.y, oz = (n0.z - o.z) * idir.z;
    // error margin: |org| and the largest |q*adj| each carry a few ulps
    const float kEps = 4.76837158203125e-7f;   // 2^-21
    float mx = kEps * __fmaf_rn(256.f, fabsf(ax), fabsf(ox));
    float my = kEps * __fmaf_rn(256.f, fabsf(ay), fabsf(oy));
    float mz = kEps * __fmaf_rn(256.f, fabsf(az), fabsf(oz));
#if BPT_FOLD_I2F
    // byte -> float without the XU pipe (ncu r01b: I2F is 7.5 % of the instructions but keeps the quarter-rate XU pipe 53 % busy):
    // PRMT builds 2^23 + q, and the 2^23 is folded into the plane offset, o' = o - 2^23 a.  o' carries an absolute rounding
    // error of at most |a| / 2 (half a quantisation cell), which the margin absorbs -- the test stays conservative.
    mx = __fmaf_rn(0.51f, fabsf(ax), mx); my = __fmaf_rn(0.51f, fabsf(ay), my); mz = __fmaf_rn(0.51f, fabsf(az), mz);
    float oxlo = __fmaf_rn(-8388608.f, ax, ox - mx), oxhi = __fmaf_rn(-8388608.f, ax, ox + mx);
    float oylo = __fmaf_rn(-8388608.f, ay, oy - my), oyhi = __fmaf_rn(-8388608.f, ay, oy + my);
    float ozlo = __fmaf_rn(-8388608.f, az, oz - mz), ozhi = __fmaf_rn(-8388608.f, az, oz + mz);
#define BPT_Q2F(word, j) __uint_as_float(__byte_perm((word), 0x4B000000u, 0x7650u + (j)))
#else
    float oxlo = ox - mx, oxhi = ox + mx, oylo = oy - my, oyhi = oy + my, ozlo = oz - mz, ozhi = oz + mz;
#define BPT_Q2F(word, j) ((float)ExtractByte((word), (j)))
#endif
    unsigned hitmask = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        unsigned meta4 = __float_as_uint(i == 0 ? n1.z : n1.w);
        unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010;
        unsigned inner_mask4 = SignExtendS8x4(is_inner4 << 3);
        unsigned bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1f;
        unsigned child_bits4 = (meta4 >> 5) & 0x07070707;
        unsigned qlox = __float_as_uint(i == 0 ? n2.x : n2.y), qhix = __float_as_uint(i == 0 ? n2.z : n2.w);
        unsigned qloy = __float_as_uint(i == 0 ? n3.x : n3.y), qhiy = __float_as_uint(i == 0 ? n3.z : n3.w);
        unsigned qloz = __float_as_uint(i == 0 ? n4.x : n4.y), qhiz = __float_as_uint(i == 0 ? n4.z : n4.w);
        unsigned xmin = idir.x < 0 ? qhix : qlox, xmax = idir.x < 0 ? qlox : qhix;
        unsigned ymin = idir.y < 0 ? qhiy : qloy, ymax = idir.y < 0 ? qloy : qhiy;
        unsigned zmin = idir.z < 0 ? qhiz : qloz, zmax = idir.z < 0 ? qloz : qhiz;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float tminx = __fmaf_rn(BPT_Q2F(xmin, j), ax, oxlo);
            float tminy = __fmaf_rn(BPT_Q2F(ymin, j), ay, oylo);
            float tminz = __fmaf_rn(BPT_Q2F(zmin, j), az, ozlo);
            float tmaxx = __fmaf_rn(BPT_Q2F(xmax, j), ax, oxhi);
            float tmaxy = __fmaf_rn(BPT_Q2F(ymax, j), ay, oyhi);
            float tmaxz = __fmaf_rn(BPT_Q2F(zmax, j), az, ozhi);
            float tmin = fmaxf(fmaxf(tminx, tminy), fmaxf(tminz, 0.f));
            float tmax = fminf(fminf(tmaxx, tmaxy), fminf(tmaxz, tMax));
            unsigned hb = ExtractByte(child_bits4, j) << ExtractByte(bit_index4, j);
            hitmask |= (tmin <= tmax) ? hb : 0u;   // branch-free: a branch here runs its body with ~3 lanes (ncu r01b)
        }
    }
#undef BPT_Q2F
    return hitmask;
}

// Traversal statistics (counting build only)
struct TraceCounters { unsigned long long nodes, tris, rays; };

template <bool ANYHIT, bool COUNT, bool SHAPES>
BPT_D void TraceRay(const BVHView &bvh, Pt3 o, Vec3 d, float tMax, HitRec *rec, unsigned *nNodes, unsigned *nTris) {
    RayPre rp = MakeRayPre(o, d);
    Vec3 idir(SafeInvDir(d.x), SafeInvDir(d.y), SafeInvDir(d.z));
    Vec3 ov(o.x, o.y, o.z);
    unsigned octinv = (d.x < 0 ? 0u : 4u) | (d.y < 0 ? 0u : 2u) | (d.z < 0 ? 0u : 1u);
    unsigned octinv4 = octinv * 0x01010101u;
    // box culling distance: the accepted-hit rule compares tScaled against tMax*det; give the boxes a little slack
    float tBox = tMax * (1 + 2 * gamma_n(3));

    uint2 stack[BPT_STACK_SIZE];
    int sp = 0;
    uint2 cur = make_uint2(0u, 0x80000000u);
    rec->tri = -1; rec->t = tMax;

    while (true) {
        uint2 triGroup;
        if (cur.y & 0xff000000u) {
            unsigned hits_imask = cur.y;
            unsigned child_bit = 31 - __clz(hits_imask);
            unsigned child_base = cur.x;
            cur.y &= ~(1u << child_bit);
            if (cur.y & 0xff000000u) stack[sp++] = cur;
            unsigned slot = (child_bit - 24) ^ (octinv4 & 0xff);
            unsigned rel = __popc(hits_imask & ~(0xffffffffu << slot));
            unsigned node = child_base + rel;
            const float4 *np = bvh.nodes + (size_t)node * 5;
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) ++*nNodes;
            unsigned hitmask = NodeIntersect(ov, idir, octinv4, tBox, n0, n1, n2, n3, n4);
            unsigned imask = ExtractByte(__float_as_uint(n0.w), 3);
            cur.x = __float_as_uint(n1.x);
            triGroup.x = __float_as_uint(n1.y);
            cur.y = (hitmask & 0xff000000u) | imask;
            triGroup.y = hitmask & 0x00ffffffu;
        } else {
            triGroup = cur;
            cur = make_uint2(0, 0);
        }
        while (triGroup.y != 0) {
            unsigned ti = 31 - __clz(triGroup.y);
            triGroup.y &= ~(1u << ti);
            unsigned tri = triGroup.x + ti;
            const float4 *tp = bvh.tris + (size_t)tri * 3;
            float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
            if (COUNT) ++*nTris;
            TriHit th;
            if (LeafIntersect<SHAPES>(bvh, rp, rec->t, a, b, c, &th)) {
                rec->tri = (int)tri; rec->b0 = th.b0; rec->b1 = th.b1; rec->b2 = th.b2; rec->t = th.t;
                if (ANYHIT) return;
                tBox = th.t * (1 + 2 * gamma_n(3));
            }
        }
        if ((cur.y & 0xff000000u) == 0) {
            if (sp == 0) break;
            cur = stack[--sp];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-synchronous persistent traversal (the production path).
//
// ncu on the first per-lane formulation (profiles/r01_*) showed 4.0 active threads per issued instruction: with
// independent thread scheduling the lanes of a warp drifted apart and every triangle test ran with one lane.
// Here the warp advances in explicit lock-step phases -- [fetch] [one node step] [triangle tests] [pop / commit] --
// separated by warp votes, so lanes doing the same kind of work execute together, and finished lanes are refilled
// from the ray queue (dynamic fetch) once enough of them are idle.
//
//   Load(i, &o, &d, &tMax)      reads ray i of the queue
//   Commit(i, rec)              stores the result of ray i (called exactly once per ray, by the lane that owns it)
// (tuning knobs are macros so that scripts/build_variants.py can sweep them on the GPU box)
#ifndef BPT_FETCH_THRESHOLD
#define BPT_FETCH_THRESHOLD 8
#endif
#ifndef BPT_FLUSH_LANES
#define BPT_FLUSH_LANES 16
#endif
#ifndef BPT_ROUND_LANES
#define BPT_ROUND_LANES 8
#endif
#ifndef BPT_TRI_BLOCK
#define BPT_TRI_BLOCK 5
#endif
#ifndef BPT_DEFER_COMMIT
#define BPT_DEFER_COMMIT 1
#endif
#ifndef BPT_DEFER_TAIL
#define BPT_DEFER_TAIL 0    // (measured +12 % slower, profiles/r01l_two_stage_triangle_test.log) two-stage triangle test: edge-function stage per round, full test batched at the end of the rounds
#endif
#ifndef BPT_FETCH_CHUNK
#define BPT_FETCH_CHUNK 0   // > 0: rays reserved from the queue cursor per atomic (single-level traversal)
#endif
#ifndef BPT_INST_GATE
#define BPT_INST_GATE 1     // > 1: gate instance entry / exit on a warp vote (experiment, see TraversePersistentInst)
#endif
#ifndef BPT_PREFETCH
#define BPT_PREFETCH 0      // (measured: +8 % slower, profiles/r01d_fold_prefetch_sweep.log) L1 prefetch of the next node / the first deferred triangle as soon as their addresses are known
#endif
constexpr int kFetchThreshold = BPT_FETCH_THRESHOLD;   // refill when at least this many lanes are idle
constexpr int kTriCap = 8;             // per-lane ring of deferred leaf groups (triBase, hit mask); one entry per node step
constexpr int kTriBlock = BPT_TRI_BLOCK;       // a lane stops taking node steps while it holds more than this many groups
constexpr int kFlushLanes = BPT_FLUSH_LANES;   // start triangle rounds when this many lanes have candidates
constexpr int kRoundLanes = BPT_ROUND_LANES;   // keep going while this many lanes still have candidates
constexpr bool kDeferCommit = BPT_DEFER_COMMIT != 0;   // finished lanes wait and commit together with the next refill

template <bool ANYHIT, bool COUNT, bool SHAPES, typename LoadF, typename CommitF>
BPT_D void TraversePersistent(const BVHView &bvh, int nRays, int *cursor, LoadF Load, CommitF Commit, unsigned *nNodes,
                              unsigned *nTris, unsigned *nRaysDone) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    __shared__ uint2 s_tri[kTriCap][128];
    uint2 (*triRing)[128] = s_tri;
    const int tid = threadIdx.x;
    int rayIdx = -1;
    bool exhausted = false;
    int wNext = 0, wEnd = 0; bool noMoreGlobal = false;   // (BPT_FETCH_CHUNK) warp-uniform
    (void)wNext; (void)wEnd; (void)noMoreGlobal;
    RayPre rp; Vec3 idir, ov; unsigned octinv4 = 0; float tBox = 0;
    uint2 stack[BPT_STACK_SIZE];
    int sp = 0;
    unsigned head = 0, tail = 0;   // deferred triangle ring [head, tail)
    uint2 cur = make_uint2(0, 0);
    HitRec rec; rec.tri = -1; rec.t = 0; rec.b0 = rec.b1 = rec.b2 = 0;
    bool done = false;             // traversal of rayIdx finished, result not yet committed
    int pend = -1;                 // (BPT_DEFER_TAIL) triangle that passed the edge-function stage, full test outstanding
    bool anyHitFound = false;
    auto finishPending = [&]() {
        const float4 *tp = bvh.tris + (size_t)pend * 3;
        float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
        TriHit th;
        if (IntersectTrianglePre(rp, rec.t, Pt3(a.x, a.y, a.z), Pt3(b.x, b.y, b.z), Pt3(c.x, c.y, c.z), &th)) {
            rec.tri = pend; rec.b0 = th.b0; rec.b1 = th.b1; rec.b2 = th.b2; rec.t = th.t;
            if (ANYHIT) anyHitFound = true;
            else tBox = th.t * (1 + 2 * gamma_n(3));
        }
        pend = -1;
    };
    (void)anyHitFound;
    while (true) {
        // ---- commit (batched: finished lanes wait until enough of them can commit and refill together)
        unsigned doneMask = __ballot_sync(FULL, done);
        unsigned idle = __ballot_sync(FULL, rayIdx < 0) | doneMask;
        if (doneMask != 0 && (!kDeferCommit || __popc(idle) >= kFetchThreshold || idle == FULL || exhausted)) {
            if (done) {
                Commit(rayIdx, rec);
                if (COUNT) ++*nRaysDone;
                rayIdx = -1; done = false;
            }
        } else {
            idle &= ~doneMask;
        }
        // ---- fetch
        if (idle != 0 && !exhausted && (__popc(idle) >= kFetchThreshold || idle == FULL)) {
            int nNeed = __popc(idle);
            int base = 0;
#if BPT_FETCH_CHUNK > 0
            // warp-private range [wNext, wEnd) of the queue, refilled BPT_FETCH_CHUNK rays at a time: one atomic on the
            // (single, contended) queue cursor per chunk instead of one per refill
            if (wNext >= wEnd) {
                if (lane == 0) base = atomicAdd(cursor, BPT_FETCH_CHUNK);
                base = __shfl_sync(FULL, base, 0);
                wNext = base; wEnd = min(base + BPT_FETCH_CHUNK, nRays);
                if (base + BPT_FETCH_CHUNK >= nRays) noMoreGlobal = true;
            }
            base = wNext;
            int take = min(nNeed, max(wEnd - wNext, 0));
            wNext += take;
            if (noMoreGlobal && wNext >= wEnd) exhausted = true;
            const int limit = base + take;
#else
            if (lane == 0) base = atomicAdd(cursor, nNeed);
            base = __shfl_sync(FULL, base, 0);
            if (base + nNeed >= nRays) exhausted = true;
            const int limit = nRays;
#endif
            if (rayIdx < 0) {
                int i = base + __popc(idle & ((1u << lane) - 1));
                if (i < limit) {
                    rayIdx = i;
                    Pt3 o; Vec3 d; float tMax;
                    Load(i, &o, &d, &tMax);
                    rp = MakeRayPre(o, d);
                    idir = Vec3(SafeInvDir(d.x), SafeInvDir(d.y), SafeInvDir(d.z));
                    ov = Vec3(o.x, o.y, o.z);
                    unsigned octinv = (d.x < 0 ? 0u : 4u) | (d.y < 0 ? 0u : 2u) | (d.z < 0 ? 0u : 1u);
                    octinv4 = octinv * 0x01010101u;
                    tBox = tMax * (1 + 2 * gamma_n(3));
                    sp = 0; head = tail = 0;
                    cur = make_uint2(0u, 0x80000000u);
                    rec.tri = -1; rec.t = tMax; anyHitFound = false;
                }
            }
        }
        bool act = rayIdx >= 0 && !done;
        if (!__any_sync(FULL, rayIdx >= 0)) break;
        // ---- one node step for every lane that has node work and room for its triangle candidates
        bool canNode = act && (cur.y & 0xff000000u) != 0 && (tail - head) <= (unsigned)kTriBlock;
        if (canNode) {
            unsigned hits_imask = cur.y;
            unsigned child_bit = 31 - __clz(hits_imask);
            unsigned child_base = cur.x;
            cur.y &= ~(1u << child_bit);
            if (cur.y & 0xff000000u) stack[sp++] = cur;
            unsigned slot = (child_bit - 24) ^ (octinv4 & 0xff);
            unsigned rel = __popc(hits_imask & ~(0xffffffffu << slot));
            const float4 *np = bvh.nodes + (size_t)(child_base + rel) * 5;
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) ++*nNodes;
            unsigned hitmask = NodeIntersect(ov, idir, octinv4, tBox, n0, n1, n2, n3, n4);
            cur.x = __float_as_uint(n1.x);
            cur.y = (hitmask & 0xff000000u) | ExtractByte(__float_as_uint(n0.w), 3);
            unsigned triBase = __float_as_uint(n1.y);
            unsigned tmask = hitmask & 0x00ffffffu;
            if (tmask) {
                triRing[tail & (kTriCap - 1)][tid] = make_uint2(triBase, tmask); ++tail;
                if (BPT_PREFETCH) PrefetchL1(bvh.tris + (size_t)(triBase + (31 - __clz(tmask))) * 3);
            }
            if ((cur.y & 0xff000000u) == 0 && sp > 0) cur = stack[--sp];
            if (BPT_PREFETCH && (cur.y & 0xff000000u)) {
                // the node the next step will visit (same arithmetic as at the top of the step)
                unsigned nb = 31 - __clz(cur.y);
                unsigned ns = (nb - 24) ^ (octinv4 & 0xff);
                const float4 *nx = bvh.nodes + (size_t)(cur.x + __popc(cur.y & ~(0xffffffffu << ns))) * 5;
                PrefetchL1(nx); PrefetchL1(nx + 4);
            }
        }
        // ---- triangle rounds (deferred so that many lanes test together)
        bool hasNode = act && (cur.y & 0xff000000u) != 0;
        bool hasTri = act && tail != head;
        unsigned triLanes = __ballot_sync(FULL, hasTri);
        bool anyNodeRunnable = __any_sync(FULL, hasNode && (tail - head) <= (unsigned)kTriBlock);
        bool finished = false;
        if (triLanes != 0 && (__popc(triLanes) >= kFlushLanes || !anyNodeRunnable)) {
            while (true) {
                if (hasTri) {
                    uint2 g = triRing[head & (kTriCap - 1)][tid];
                    unsigned ti = 31 - __clz(g.y);
                    g.y &= ~(1u << ti);
                    if (g.y == 0) ++head; else triRing[head & (kTriCap - 1)][tid].y = g.y;
                    unsigned tri = g.x + ti;
                    const float4 *tp = bvh.tris + (size_t)tri * 3;
                    float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                    if (COUNT) ++*nTris;
#if BPT_DEFER_TAIL
                    if (!(__float_as_uint(b.w) & 0x80000000u) &&
                        IntersectTriangleHead(rp, Pt3(a.x, a.y, a.z), Pt3(b.x, b.y, b.z), Pt3(c.x, c.y, c.z))) {
                        if (pend >= 0) finishPending();   // second survivor of this lane in one batch (rare): resolve the first now
                        pend = (int)tri;
                    }
#else
                    TriHit th;
                    if (LeafIntersect<SHAPES>(bvh, rp, rec.t, a, b, c, &th)) {
                        rec.tri = (int)tri; rec.b0 = th.b0; rec.b1 = th.b1; rec.b2 = th.b2; rec.t = th.t;
                        if (ANYHIT) { finished = true; head = tail; }
                        else tBox = th.t * (1 + 2 * gamma_n(3));
                    }
#endif
                    hasTri = tail != head;
                }
                unsigned m = __ballot_sync(FULL, hasTri);
                if (m == 0) break;
                if (__popc(m) < kRoundLanes && anyNodeRunnable) break;
            }
#if BPT_DEFER_TAIL
            // second stage for every lane that parked a survivor during this batch of rounds
            if (pend >= 0) finishPending();
            if (ANYHIT && anyHitFound) { finished = true; head = tail; }
#endif
        }
        // ---- mark finished rays (committed at the top of the loop)
        if (act) {
            if (!finished && (cur.y & 0xff000000u) == 0 && sp == 0 && tail == head) finished = true;
            if (finished) { done = true; cur = make_uint2(0, 0); sp = 0; head = tail = 0; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Two-level variant (object instancing): the top-level BVH's leaf records are instances; entering one restates
// TransformedPrimitive::Intersect (cpu/primitive.cpp:112-124): the ray is taken to instance space with
// Transform::ApplyInverse(Ray, &tMax) (util/transform.h:416-430, origin error offset included), the instance's BVH is
// traversed with the shifted tMax, and -- like the reference -- the reported tHit stays in the instance-space
// parameterisation.  Stack protocol: [saved top-level node group][sentinel] are pushed on entry; the sentinel is popped
// only after the lane's deferred triangle groups (which belong to the instance) are drained.
template <bool ANYHIT, bool COUNT, bool SHAPES, typename LoadF, typename CommitF>
BPT_D void TraversePersistentInst(const BVHView &bvh, const DeviceInstance *instances, int nRays, int *cursor, LoadF Load,
                                  CommitF Commit, unsigned *nNodes, unsigned *nTris, unsigned *nRaysDone) {
    const unsigned FULL = 0xffffffffu;
    const unsigned SENT = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    __shared__ uint2 s_tri[kTriCap][128];
    uint2 (*triRing)[128] = s_tri;
    const int tid = threadIdx.x;
    int rayIdx = -1;
    bool exhausted = false;
    RayPre rp; Vec3 idir, ov; unsigned octinv4 = 0; float tBox = 0;
    uint2 stack[BPT_STACK_SIZE];
    int sp = 0;
    unsigned head = 0, tail = 0;
    uint2 cur = make_uint2(0, 0), instGroup = make_uint2(0, 0);
    int inst = -1;
    float tSaved = 0; bool hitHere = false;
    const float4 *curNodes = bvh.nodes, *curTris = bvh.tris;
    HitRec rec; rec.tri = -1; rec.t = 0; rec.b0 = rec.b1 = rec.b2 = 0; rec.inst = -1;
    auto setRay = [&](Pt3 o, Vec3 d) {
        rp = MakeRayPre(o, d);
        idir = Vec3(SafeInvDir(d.x), SafeInvDir(d.y), SafeInvDir(d.z));
        ov = Vec3(o.x, o.y, o.z);
        unsigned octinv = (d.x < 0 ? 0u : 4u) | (d.y < 0 ? 0u : 2u) | (d.z < 0 ? 0u : 1u);
        octinv4 = octinv * 0x01010101u;
    };
    bool done = false;
    while (true) {
        unsigned doneMask = __ballot_sync(FULL, done);
        unsigned idle = __ballot_sync(FULL, rayIdx < 0) | doneMask;
        if (doneMask != 0 && (!kDeferCommit || __popc(idle) >= kFetchThreshold || idle == FULL || exhausted)) {
            if (done) {
                Commit(rayIdx, rec);
                if (COUNT) ++*nRaysDone;
                rayIdx = -1; done = false;
            }
        } else {
            idle &= ~doneMask;
        }
        if (idle != 0 && !exhausted && (__popc(idle) >= kFetchThreshold || idle == FULL)) {
            int nNeed = __popc(idle);
            int base = 0;
            if (lane == 0) base = atomicAdd(cursor, nNeed);
            base = __shfl_sync(FULL, base, 0);
            if (base + nNeed >= nRays) exhausted = true;
            if (rayIdx < 0) {
                int i = base + __popc(idle & ((1u << lane) - 1));
                if (i < nRays) {
                    rayIdx = i;
                    Pt3 o; Vec3 d; float tMax;
                    Load(i, &o, &d, &tMax);
                    setRay(o, d);
                    tBox = tMax * (1 + 1e-5f);
                    sp = 0; head = tail = 0; inst = -1; instGroup = make_uint2(0, 0);
                    curNodes = bvh.nodes; curTris = bvh.tris;
                    cur = make_uint2(0u, 0x80000000u);
                    rec.tri = -1; rec.t = tMax; rec.inst = -1;
                }
            }
        }
        bool act = rayIdx >= 0 && !done;
        if (!__any_sync(FULL, rayIdx >= 0)) break;
        // ---- enter the next instance of the pending top-level leaf group
#if BPT_INST_GATE > 1
        // EXPERIMENT, off by default and not yet validated on the GPU (DESIGN.md "Known gaps"): entering / leaving an instance is
        // ~150 instructions that ncu (profiles/r01o) shows running with 4.5 active lanes, so lanes wait at these two points until
        // BPT_INST_GATE of them are there -- or until no lane of the warp can take a node step.  Only node-runnable lanes count as
        // busy: they always make progress, so a waiting lane cannot wait forever (the first attempt also counted lanes that only
        // held deferred triangles, whose flush in turn waited for the lanes waiting here: a deadlock).
        const bool wantEnter = act && inst < 0 && instGroup.y != 0;
        const bool wantLeave = act && inst >= 0 && (cur.y & 0xff000000u) == 0 && tail == head && sp > 0 && stack[sp - 1].x == SENT;
        const bool nodeRunnable = act && !wantEnter && !wantLeave && (cur.y & 0xff000000u) != 0 && (tail - head) <= (unsigned)kTriBlock;
        const bool anyBusy = __any_sync(FULL, nodeRunnable);
        const bool enterNow = __popc(__ballot_sync(FULL, wantEnter)) >= BPT_INST_GATE || !anyBusy;
        const bool leaveNow = __popc(__ballot_sync(FULL, wantLeave)) >= BPT_INST_GATE || !anyBusy;
#else
        const bool enterNow = true, leaveNow = true;
#endif
        if (act && inst < 0 && instGroup.y != 0 && enterNow) {
            unsigned ti = 31 - __clz(instGroup.y);
            instGroup.y &= ~(1u << ti);
            int id = __float_as_int(__ldg(bvh.tris + (size_t)(instGroup.x + ti) * 3).w);
            const DeviceInstance &I = instances[id];
            stack[sp++] = cur;
            stack[sp++] = make_uint2(SENT, 0);
            inst = id; tSaved = rec.t; hitHere = false;
            curNodes = I.nodes; curTris = I.tris;
            if (!I.identity) {
                Pt3 o; Vec3 d; float tm;
                Load(rayIdx, &o, &d, &tm);
                InstRay r = ApplyInverseRay(I.mInv, o, d);
                setRay(r.o, r.d);
                rec.t = rec.t - r.dt;
            }
            tBox = rec.t * (1 + 1e-5f);
            cur = make_uint2(0u, 0x80000000u);
        }
        // ---- one node step
        bool canNode = act && (cur.y & 0xff000000u) != 0 && (tail - head) <= (unsigned)kTriBlock && (inst >= 0 || instGroup.y == 0);
        if (canNode) {
            unsigned hits_imask = cur.y;
            unsigned child_bit = 31 - __clz(hits_imask);
            unsigned child_base = cur.x;
            cur.y &= ~(1u << child_bit);
            if (cur.y & 0xff000000u) stack[sp++] = cur;
            unsigned slot = (child_bit - 24) ^ (octinv4 & 0xff);
            unsigned rel = __popc(hits_imask & ~(0xffffffffu << slot));
            const float4 *np = curNodes + (size_t)(child_base + rel) * 5;
            float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
            if (COUNT) ++*nNodes;
            unsigned hitmask = NodeIntersect(ov, idir, octinv4, tBox, n0, n1, n2, n3, n4);
            cur.x = __float_as_uint(n1.x);
            cur.y = (hitmask & 0xff000000u) | ExtractByte(__float_as_uint(n0.w), 3);
            unsigned triBase = __float_as_uint(n1.y);
            unsigned tmask = hitmask & 0x00ffffffu;
            if (tmask) {
                if (inst < 0) instGroup = make_uint2(triBase, tmask);
                else { triRing[tail & (kTriCap - 1)][tid] = make_uint2(triBase, tmask); ++tail; }
            }
            if ((cur.y & 0xff000000u) == 0 && sp > 0 && stack[sp - 1].x != SENT) cur = stack[--sp];
        } else if (act && (cur.y & 0xff000000u) == 0 && sp > 0 && stack[sp - 1].x != SENT && (inst >= 0 || instGroup.y == 0)) {
            cur = stack[--sp];
        }
        // ---- triangle rounds
        bool hasNode = act && (cur.y & 0xff000000u) != 0;
        bool hasTri = act && tail != head;
        unsigned triLanes = __ballot_sync(FULL, hasTri);
        bool anyNodeRunnable = __any_sync(FULL, (hasNode && (tail - head) <= (unsigned)kTriBlock) || (act && inst < 0 && instGroup.y != 0));
        bool finished = false;
        if (triLanes != 0 && (__popc(triLanes) >= kFlushLanes || !anyNodeRunnable)) {
            while (true) {
                if (hasTri) {
                    uint2 g = triRing[head & (kTriCap - 1)][tid];
                    unsigned ti = 31 - __clz(g.y);
                    g.y &= ~(1u << ti);
                    if (g.y == 0) ++head; else triRing[head & (kTriCap - 1)][tid].y = g.y;
                    unsigned tri = g.x + ti;
                    const float4 *tp = curTris + (size_t)tri * 3;
                    float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
                    if (COUNT) ++*nTris;
                    TriHit th;
                    if (LeafIntersect<SHAPES>(bvh, rp, rec.t, a, b, c, &th)) {
                        rec.tri = (int)tri; rec.b0 = th.b0; rec.b1 = th.b1; rec.b2 = th.b2; rec.t = th.t; rec.inst = inst;
                        hitHere = true;
                        if (ANYHIT) { finished = true; head = tail; }
                        else tBox = th.t * (1 + 1e-5f);
                    }
                    hasTri = tail != head;
                }
                unsigned m = __ballot_sync(FULL, hasTri);
                if (m == 0) break;
                if (__popc(m) < kRoundLanes && anyNodeRunnable) break;
            }
        }
        // ---- leave the instance / commit
        if (act && !finished) {
            bool nodeEmpty = (cur.y & 0xff000000u) == 0;
            if (leaveNow && inst >= 0 && nodeEmpty && tail == head && sp > 0 && stack[sp - 1].x == SENT) {
                --sp;                      // sentinel
                cur = stack[--sp];         // saved top-level node group
                if (!instances[inst].identity) {
                    Pt3 o; Vec3 d; float tm;
                    Load(rayIdx, &o, &d, &tm);
                    setRay(o, d);
                }
                if (!hitHere) rec.t = tSaved;
                inst = -1;
                curNodes = bvh.nodes; curTris = bvh.tris;
                tBox = rec.t * (1 + 1e-5f);
                nodeEmpty = (cur.y & 0xff000000u) == 0;
            }
            if (inst < 0 && nodeEmpty && sp == 0 && instGroup.y == 0 && tail == head) finished = true;
        }
        if (act && finished) {
            done = true;
            cur = make_uint2(0, 0); sp = 0; head = tail = 0; inst = -1; instGroup = make_uint2(0, 0);
        }
    }
}

}  // namespace bpt
