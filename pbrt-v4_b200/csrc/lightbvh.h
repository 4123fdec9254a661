// lightbvh.h -- host-side construction of the light BVH, restating BVHLightSampler's constructor and buildBVH
// (lightsamplers.cpp:  BVHLightSampler::BVHLightSampler / buildBVH), LightBounds Union (lights.h:137-155),
// DirectionCone Union (util/vecmath.cpp:56-86), CompactLightBounds quantisation (lightsamplers.h:104-225) and
// OctahedralVector (util/vecmath.h).  Node order, split decisions and quantised values must equal the reference's
// for the light-selection probabilities on the path to match.
#pragma once
#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>
#include "scene.h"
#include "xform.h"

namespace bpt {

struct LightBoundsH {
    Bounds3 bounds;
    float phi = 0;
    Vec3 w;
    float cosTheta_o = 0, cosTheta_e = 0;
    bool twoSided = false;
    bool valid = false;    // false: light has no bounds (infinite light)
    LightBoundsH() {}
    LightBoundsH(const Bounds3 &b, Vec3 w_, float phi_, float co, float ce, bool ts)
        : bounds(b), phi(phi_), w(Normalize(w_)), cosTheta_o(co), cosTheta_e(ce), twoSided(ts), valid(true) {}
    Pt3 Centroid() const { return (bounds.pMin + bounds.pMax) / 2; }
};

struct DirCone {
    Vec3 w;
    float cosTheta = BPT_INFINITY;
    DirCone() {}
    DirCone(Vec3 w_, float c) : w(Normalize(w_)), cosTheta(c) {}
    bool IsEmpty() const { return cosTheta == BPT_INFINITY; }
    static DirCone EntireSphere() { return DirCone(Vec3(0, 0, 1), -1); }
};
inline DirCone UnionCone(const DirCone &a, const DirCone &b) {
    if (a.IsEmpty()) return b;
    if (b.IsEmpty()) return a;
    float theta_a = SafeACos(a.cosTheta), theta_b = SafeACos(b.cosTheta);
    float theta_d = AngleBetween(a.w, b.w);
    if (stdmin(theta_d + theta_b, kPi) <= theta_a) return a;
    if (stdmin(theta_d + theta_a, kPi) <= theta_b) return b;
    float theta_o = (theta_a + theta_d + theta_b) / 2;
    if (theta_o >= kPi) return DirCone::EntireSphere();
    float theta_r = theta_o - theta_a;
    Vec3 wr = Cross(a.w, b.w);
    if (LengthSquared(wr) == 0) return DirCone::EntireSphere();
    Vec3 w = Rotate((180 / kPi) * theta_r, wr)(a.w);
    return DirCone(w, std::cos(theta_o));
}
inline LightBoundsH UnionLB(const LightBoundsH &a, const LightBoundsH &b) {
    if (a.phi == 0) return b;
    if (b.phi == 0) return a;
    DirCone cone = UnionCone(DirCone(a.w, a.cosTheta_o), DirCone(b.w, b.cosTheta_o));
    float cosTheta_o = cone.cosTheta;
    float cosTheta_e = stdmin(a.cosTheta_e, b.cosTheta_e);
    return LightBoundsH(Union(a.bounds, b.bounds), cone.w, a.phi + b.phi, cosTheta_o, cosTheta_e, a.twoSided | b.twoSided);
}

struct LightBVHBuilder {
    std::vector<DeviceLightBVHNode> nodes;
    std::vector<uint32_t> bitTrails;
    std::vector<int> lightInBVH;
    std::vector<int> infiniteLights;
    Bounds3 allLightBounds;

    static float SignF(float v) { return std::copysign(1.f, v); }
    static uint16_t OctEncode(float f) { return (uint16_t)std::round(Clampf((f + 1) / 2, 0, 1) * 65535.f); }
    static Vec3 OctRoundTrip(Vec3 v) {
        // OctahedralVector(Vector3f) then operator Vector3f()
        float s = std::abs(v.x) + std::abs(v.y) + std::abs(v.z);
        v = v / s;
        uint16_t x, y;
        if (v.z >= 0) { x = OctEncode(v.x); y = OctEncode(v.y); }
        else { x = OctEncode((1 - std::abs(v.y)) * SignF(v.x)); y = OctEncode((1 - std::abs(v.x)) * SignF(v.y)); }
        Vec3 r;
        r.x = -1 + 2 * (x / 65535.f);
        r.y = -1 + 2 * (y / 65535.f);
        r.z = 1 - (std::abs(r.x) + std::abs(r.y));
        if (r.z < 0) {
            float xo = r.x;
            r.x = (1 - std::abs(r.y)) * SignF(xo);
            r.y = (1 - std::abs(xo)) * SignF(r.y);
        }
        return Normalize(r);
    }
    static float QuantizeBounds(float c, float mn, float mx) {
        if (mn == mx) return 0;
        return 65535.f * Clampf((c - mn) / (mx - mn), 0, 1);
    }
    DeviceLightBVHNode Compact(const LightBoundsH &lb) const {
        DeviceLightBVHNode n;
        std::memset(&n, 0, sizeof(n));
        n.w = OctRoundTrip(Normalize(lb.w));
        n.phi = lb.phi;
        unsigned qo = (unsigned)std::floor(32767.f * ((lb.cosTheta_o + 1) / 2)) & 0x7fffu;
        unsigned qe = (unsigned)std::floor(32767.f * ((lb.cosTheta_e + 1) / 2)) & 0x7fffu;
        n.cosTheta_o = 2 * (qo / 32767.f) - 1;
        n.cosTheta_e = 2 * (qe / 32767.f) - 1;
        n.twoSided = lb.twoSided ? 1 : 0;
        uint16_t qb[2][3];
        const Bounds3 &allb = allLightBounds;
        for (int c = 0; c < 3; ++c) {
            qb[0][c] = (uint16_t)std::floor(QuantizeBounds(lb.bounds.pMin[c], allb.pMin[c], allb.pMax[c]));
            qb[1][c] = (uint16_t)std::ceil(QuantizeBounds(lb.bounds.pMax[c], allb.pMin[c], allb.pMax[c]));
        }
        n.bmin = Pt3(Lerp(qb[0][0] / 65535.f, allb.pMin.x, allb.pMax.x), Lerp(qb[0][1] / 65535.f, allb.pMin.y, allb.pMax.y),
                     Lerp(qb[0][2] / 65535.f, allb.pMin.z, allb.pMax.z));
        n.bmax = Pt3(Lerp(qb[1][0] / 65535.f, allb.pMin.x, allb.pMax.x), Lerp(qb[1][1] / 65535.f, allb.pMin.y, allb.pMax.y),
                     Lerp(qb[1][2] / 65535.f, allb.pMin.z, allb.pMax.z));
        return n;
    }
    static float EvaluateCost(const LightBoundsH &b, const Bounds3 &bounds, int dim) {
        float theta_o = std::acos(b.cosTheta_o), theta_e = std::acos(b.cosTheta_e);
        float theta_w = stdmin(theta_o + theta_e, kPi);
        float sinTheta_o = SafeSqrt(1 - Sqr(b.cosTheta_o));
        float M_omega = 2 * kPi * (1 - b.cosTheta_o) +
                        kPi / 2 * (2 * theta_w * sinTheta_o - std::cos(theta_o - 2 * theta_w) - 2 * theta_o * sinTheta_o + b.cosTheta_o);
        float Kr = MaxComponentValue(bounds.Diagonal()) / bounds.Diagonal()[dim];
        return b.phi * M_omega * Kr * b.bounds.SurfaceArea();
    }
    typedef std::pair<int, LightBoundsH> Item;
    std::pair<int, LightBoundsH> buildBVH(std::vector<Item> &bvhLights, int start, int end, uint32_t bitTrail, int depth) {
        if (end - start == 1) {
            int nodeIndex = (int)nodes.size();
            DeviceLightBVHNode cb = Compact(bvhLights[start].second);
            int lightIndex = bvhLights[start].first;
            cb.childOrLightIndex = lightIndex; cb.isLeaf = 1;
            nodes.push_back(cb);
            bitTrails[lightIndex] = bitTrail;
            lightInBVH[lightIndex] = 1;
            return {nodeIndex, bvhLights[start].second};
        }
        Bounds3 bounds, centroidBounds;
        for (int i = start; i < end; ++i) {
            const LightBoundsH &lb = bvhLights[i].second;
            bounds = Union(bounds, lb.bounds);
            centroidBounds = Union(centroidBounds, lb.Centroid());
        }
        float minCost = BPT_INFINITY;
        int minCostSplitBucket = -1, minCostSplitDim = -1;
        constexpr int nBuckets = 12;
        for (int dim = 0; dim < 3; ++dim) {
            if (centroidBounds.pMax[dim] == centroidBounds.pMin[dim]) continue;
            LightBoundsH bucketLightBounds[nBuckets];
            for (int i = start; i < end; ++i) {
                Pt3 pc = bvhLights[i].second.Centroid();
                int b = (int)(nBuckets * centroidBounds.Offset(pc)[dim]);
                if (b == nBuckets) b = nBuckets - 1;
                bucketLightBounds[b] = UnionLB(bucketLightBounds[b], bvhLights[i].second);
            }
            float cost[nBuckets - 1];
            for (int i = 0; i < nBuckets - 1; ++i) {
                LightBoundsH b0, b1;
                for (int j = 0; j <= i; ++j) b0 = UnionLB(b0, bucketLightBounds[j]);
                for (int j = i + 1; j < nBuckets; ++j) b1 = UnionLB(b1, bucketLightBounds[j]);
                cost[i] = EvaluateCost(b0, bounds, dim) + EvaluateCost(b1, bounds, dim);
            }
            for (int i = 1; i < nBuckets - 1; ++i) {
                if (cost[i] > 0 && cost[i] < minCost) { minCost = cost[i]; minCostSplitBucket = i; minCostSplitDim = dim; }
            }
        }
        int mid;
        if (minCostSplitDim == -1) mid = (start + end) / 2;
        else {
            Item *pmid = std::partition(&bvhLights[start], &bvhLights[end - 1] + 1, [=](const Item &l) {
                int b = (int)(nBuckets * centroidBounds.Offset(l.second.Centroid())[minCostSplitDim]);
                if (b == nBuckets) b = nBuckets - 1;
                return b <= minCostSplitBucket;
            });
            mid = (int)(pmid - &bvhLights[0]);
            if (mid == start || mid == end) mid = (start + end) / 2;
        }
        int nodeIndex = (int)nodes.size();
        nodes.push_back(DeviceLightBVHNode());
        std::pair<int, LightBoundsH> child0 = buildBVH(bvhLights, start, mid, bitTrail, depth + 1);
        std::pair<int, LightBoundsH> child1 = buildBVH(bvhLights, mid, end, bitTrail | (1u << depth), depth + 1);
        LightBoundsH lb = UnionLB(child0.second, child1.second);
        DeviceLightBVHNode cb = Compact(lb);
        cb.childOrLightIndex = child1.first; cb.isLeaf = 0;
        nodes[nodeIndex] = cb;
        return {nodeIndex, lb};
    }
    void Build(const std::vector<LightBoundsH> &lights) {
        nodes.clear(); infiniteLights.clear();
        bitTrails.assign(lights.size(), 0xffffffffu);
        lightInBVH.assign(lights.size(), 0);
        allLightBounds = Bounds3();
        std::vector<Item> bvhLights;
        for (size_t i = 0; i < lights.size(); ++i) {
            if (!lights[i].valid) infiniteLights.push_back((int)i);
            else if (lights[i].phi > 0) {
                bvhLights.push_back(std::make_pair((int)i, lights[i]));
                allLightBounds = Union(allLightBounds, lights[i].bounds);
            }
        }
        if (!bvhLights.empty()) buildBVH(bvhLights, 0, (int)bvhLights.size(), 0, 0);
    }
};

}  // namespace bpt
