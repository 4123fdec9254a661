// bvh_build_host.cpp -- host harness of the BVH builder: runs the SAME per-element code as the CUDA kernels
// (pbrt-v4_b200/csrc/bvh_build_core.h: Morton keys, Karras radix tree, treelet restructuring, 8-wide collapse) sequentially on
// the CPU, validates the result (every primitive in exactly one leaf, every child box inside its quantised parent slot, node
// and triangle counters consistent) and measures tree quality: SAH cost of the binary tree and wide nodes visited / triangles
// tested per ray for a path-traced ray distribution (camera rays + cosine-sampled bounces), with the traversal order of
// bvh.cuh's TraceRay.  Test infrastructure: nothing here is linked into the product.
//
//   bvh_build_host <scene> <size> <treelet rounds> [peraxis]      scene: heightfield | soup | boxes
// exit code 0 = valid tree; prints one JSON line.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <string>
#include <vector>
#include "bvh_build_core.h"

using namespace bpt;
using namespace bpt::bvhb;

struct V3 { float x, y, z; };
static V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
static V3 norm(V3 a) { float l = std::sqrt(dot(a, a)); return a * (1.f / l); }

struct Built {
    int n = 0;
    std::vector<float4> tris;          // input, 3 per primitive
    std::vector<int> left, right, parent, cnt, sortedVals;
    std::vector<float4> nmin, nmax;
    std::vector<float4> nodes, trisOut;
    int nNodes = 0, depth = 0;
    bool boxes = false;
};

static void MakeHeightfield(int n, std::vector<float4> &t) {
    std::mt19937 rng(1234);
    std::normal_distribution<float> N(0.f, 1.f);
    std::vector<V3> P((size_t)n * n);
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) {
            float x = -1.f + 2.f * i / (n - 1), z = -1.f + 2.f * j / (n - 1);
            P[(size_t)j * n + i] = {x, 0.05f * std::sin(9 * x) * std::cos(7 * z) + 0.01f * N(rng), z};
        }
    auto put = [&](V3 a, V3 b, V3 c) {
        t.push_back(make_float4(a.x, a.y, a.z, 0)); t.push_back(make_float4(b.x, b.y, b.z, 0)); t.push_back(make_float4(c.x, c.y, c.z, 0));
    };
    for (int pass = 0; pass < 2; ++pass)
        for (int j = 0; j < n - 1; ++j)
            for (int i = 0; i < n - 1; ++i) {
                V3 v00 = P[(size_t)j * n + i], v10 = P[(size_t)j * n + i + 1], v01 = P[(size_t)(j + 1) * n + i], v11 = P[(size_t)(j + 1) * n + i + 1];
                if (pass == 0) put(v00, v01, v11); else put(v00, v11, v10);
            }
}
static void MakeSoup(int n, std::vector<float4> &t) {
    std::mt19937 rng(7);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    for (int i = 0; i < n; ++i) {
        V3 c{U(rng), U(rng), U(rng)};
        float s = 0.02f + 0.05f * std::fabs(U(rng));
        V3 p[3];
        for (int k = 0; k < 3; ++k) p[k] = c + V3{U(rng), U(rng), U(rng)} * s;
        for (int k = 0; k < 3; ++k) t.push_back(make_float4(p[k].x, p[k].y, p[k].z, 0));
    }
}
// instance bounds of config 4: spheres of random radius in a slab, as the degenerate "box primitives" of the top-level BVH
static void MakeBoxes(int n, std::vector<float4> &t) {
    std::mt19937 rng(99);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    float ext = 50.f * std::cbrt(n / 100000.f);
    for (int i = 0; i < n; ++i) {
        V3 c{(2 * U(rng) - 1) * ext, (2 * U(rng) - 1) * 0.1f * ext, (2 * U(rng) - 1) * ext};
        float r = (0.35f + 0.35f * U(rng)) * (ext / 50.f) * 50.f / std::cbrt((float)n) * 0.9f;
        t.push_back(make_float4(c.x - r, c.y - r, c.z - r, 0)); t.push_back(make_float4(c.x + r, c.y + r, c.z + r, 0));
        t.push_back(make_float4(c.x - r, c.y - r, c.z - r, 0));
    }
}

static Tree View(Built &b) { return Tree{b.n, b.left.data(), b.right.data(), b.parent.data(), b.cnt.data(), b.nmin.data(), b.nmax.data()}; }

static void PostOrder(const Built &b, std::vector<int> &order) {
    order.clear();
    if (b.n < 2) return;
    std::vector<std::pair<int, int>> st;
    st.push_back({0, 0});
    while (!st.empty()) {
        auto &[v, state] = st.back();
        if (v >= b.n - 1) { st.pop_back(); continue; }
        if (state == 0) { state = 1; st.push_back({b.left[v], 0}); }
        else if (state == 1) { state = 2; st.push_back({b.right[v], 0}); }
        else { order.push_back(v); st.pop_back(); }
    }
}

static bool Build(Built &b, int rounds, bool perAxis, int leafMax) {
    const int n = b.n = (int)b.tris.size() / 3;
    std::vector<float4> lo(n), hi(n);
    float cb[6] = {3.4e38f, 3.4e38f, 3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    for (int i = 0; i < n; ++i) {
        TriBounds(b.tris[3 * i], b.tris[3 * i + 1], b.tris[3 * i + 2], &lo[i], &hi[i]);
        float c[3] = {0.5f * lo[i].x + 0.5f * hi[i].x, 0.5f * lo[i].y + 0.5f * hi[i].y, 0.5f * lo[i].z + 0.5f * hi[i].z};
        for (int k = 0; k < 3; ++k) { cb[k] = std::min(cb[k], c[k]); cb[3 + k] = std::max(cb[3 + k], c[k]); }
    }
    std::vector<uint64_t> keys(n);
    for (int i = 0; i < n; ++i) {
        if (!perAxis) keys[i] = MortonKey(lo[i], hi[i], cb);
        else {   // round-1 behaviour, for comparison: every axis normalised to its own extent
            float c[3] = {0.5f * lo[i].x + 0.5f * hi[i].x, 0.5f * lo[i].y + 0.5f * hi[i].y, 0.5f * lo[i].z + 0.5f * hi[i].z};
            uint64_t q[3];
            for (int k = 0; k < 3; ++k) {
                float e = cb[3 + k] - cb[k];
                float t = e > 0 ? (c[k] - cb[k]) / e : 0.f;
                t = std::min(std::max(t, 0.f), 1.f);
                q[k] = (uint64_t)std::min(t * 2097152.f, 2097151.f);
            }
            keys[i] = (Spread21(q[0]) << 2) | (Spread21(q[1]) << 1) | Spread21(q[2]);
        }
    }
    b.sortedVals.resize(n);
    std::iota(b.sortedVals.begin(), b.sortedVals.end(), 0);
    std::stable_sort(b.sortedVals.begin(), b.sortedVals.end(), [&](int a, int c) { return keys[a] < keys[c]; });
    std::vector<uint64_t> skeys(n);
    for (int i = 0; i < n; ++i) skeys[i] = keys[b.sortedVals[i]];
    b.left.assign(std::max(n - 1, 1), 0); b.right.assign(std::max(n - 1, 1), 0); b.parent.assign(2 * n, -1); b.cnt.assign(2 * n, 0);
    b.nmin.resize(2 * n); b.nmax.resize(2 * n);
    for (int i = 0; i < n - 1; ++i) {
        int lc, rc;
        RadixTreeNode(skeys.data(), n, i, &lc, &rc);
        b.left[i] = lc; b.right[i] = rc; b.parent[lc] = i; b.parent[rc] = i;
    }
    for (int j = 0; j < n; ++j) { int id = (n - 1) + j; b.nmin[id] = lo[b.sortedVals[j]]; b.nmax[id] = hi[b.sortedVals[j]]; b.cnt[id] = 1; }
    std::vector<int> order;
    PostOrder(b, order);
    for (int v : order) {
        b.nmin[v] = Min4(b.nmin[b.left[v]], b.nmin[b.right[v]]); b.nmax[v] = Max4(b.nmax[b.left[v]], b.nmax[b.right[v]]);
        b.cnt[v] = b.cnt[b.left[v]] + b.cnt[b.right[v]];
    }
    Tree T = View(b);
    int gamma = kTreeletLeaves;
    for (int r = 0; r < rounds; ++r) {
        PostOrder(b, order);
        long changed = 0;
        for (int v : order) if (b.cnt[v] >= gamma) changed += OptimizeTreelet(T, v) ? 1 : 0;
        fprintf(stderr, "round %d gamma %d: %ld treelets restructured\n", r, gamma, changed);
        gamma *= 2;
    }
    // collapse
    b.nodes.assign(5 * (size_t)n, make_float4(0, 0, 0, 0)); b.trisOut.assign(3 * (size_t)n, make_float4(0, 0, 0, 0));
    std::vector<CollapseItem> cur(n), nxt(n);
    int counters[3] = {1, 0, 0};
    CollapseCtx C{T, b.sortedVals.data(), b.tris.data(), b.nodes.data(), b.trisOut.data(), &counters[0], &counters[1], nxt.data(), &counters[2], leafMax};
    cur[0] = CollapseItem{n > 1 ? 0 : 0, 0};
    int nItems = 1;
    b.depth = 0;
    while (nItems > 0) {
        ++b.depth;
        counters[2] = 0;
        C.nextItems = nxt.data();
        for (int i = 0; i < nItems; ++i) CollapseOne(C, cur[i]);
        nItems = counters[2];
        std::swap(cur, nxt);
    }
    b.nNodes = counters[0];
    if (counters[1] != n) { fprintf(stderr, "emitted %d of %d primitives\n", counters[1], n); return false; }
    return true;
}

// ---- validation + traversal statistics over the wide nodes
struct WNode {
    float p[3], scale[3]; unsigned imask, childBase, triBase; unsigned char meta[8], qlo[3][8], qhi[3][8];
};
// internal child: 0b001_sssss with sssss = 24 + slot; a 1-triangle leaf has the same top bits but an offset <= 23
static inline bool IsInner(unsigned char m) { return (m & 0xe0) == 0x20 && (m & 0x18) == 0x18; }
static WNode Decode(const Built &b, int node) {
    const float4 *np = b.nodes.data() + (size_t)node * 5;
    WNode w;
    w.p[0] = np[0].x; w.p[1] = np[0].y; w.p[2] = np[0].z;
    unsigned e = f2b(np[0].w);
    for (int a = 0; a < 3; ++a) w.scale[a] = b2f(((e >> (8 * a)) & 0xff) << 23);
    w.imask = e >> 24;
    w.childBase = f2b(np[1].x); w.triBase = f2b(np[1].y);
    unsigned m0 = f2b(np[1].z), m1 = f2b(np[1].w);
    for (int s = 0; s < 4; ++s) { w.meta[s] = (m0 >> (8 * s)) & 0xff; w.meta[4 + s] = (m1 >> (8 * s)) & 0xff; }
    const float4 q[3] = {np[2], np[3], np[4]};
    for (int a = 0; a < 3; ++a) {
        unsigned l0 = f2b(q[a].x), l1 = f2b(q[a].y), h0 = f2b(q[a].z), h1 = f2b(q[a].w);
        for (int s = 0; s < 4; ++s) {
            w.qlo[a][s] = (l0 >> (8 * s)) & 0xff; w.qlo[a][4 + s] = (l1 >> (8 * s)) & 0xff;
            w.qhi[a][s] = (h0 >> (8 * s)) & 0xff; w.qhi[a][4 + s] = (h1 >> (8 * s)) & 0xff;
        }
    }
    return w;
}

static bool Validate(const Built &b) {
    std::vector<int> seen(b.n, 0);
    std::vector<int> stack{0};
    long nodes = 0;
    while (!stack.empty()) {
        int v = stack.back(); stack.pop_back();
        ++nodes;
        WNode w = Decode(b, v);
        int rel = 0;
        for (int s = 0; s < 8; ++s) {
            if (!w.meta[s]) continue;
            float lo[3], hi[3];
            for (int a = 0; a < 3; ++a) { lo[a] = w.p[a] + w.qlo[a][s] * w.scale[a]; hi[a] = w.p[a] + w.qhi[a][s] * w.scale[a]; }
            if (IsInner(w.meta[s])) {
                if (!(w.imask >> s & 1)) { fprintf(stderr, "imask mismatch\n"); return false; }
                int child = w.childBase + rel++;
                // the child's own frame must lie inside this slot's box (its origin is the child's exact minimum)
                WNode c = Decode(b, child);
                for (int a = 0; a < 3; ++a) if (c.p[a] < lo[a] - 1e-6f * std::fabs(lo[a]) - 1e-30f) { fprintf(stderr, "child %d origin outside slot box\n", child); return false; }
                stack.push_back(child);
            } else {
                int count = (w.meta[s] >> 5) == 1 ? 1 : ((w.meta[s] >> 5) == 3 ? 2 : 3);
                int off = w.meta[s] & 31;
                for (int t = 0; t < count; ++t) {
                    size_t o = 3 * (size_t)(w.triBase + off + t);
                    int orig = (int)f2b(b.trisOut[o].w);
                    if (orig < 0 || orig >= b.n) { fprintf(stderr, "bad primitive id\n"); return false; }
                    ++seen[orig];
                    for (int k = 0; k < 3; ++k) {
                        float4 p = b.trisOut[o + k];
                        float pv[3] = {p.x, p.y, p.z};
                        for (int a = 0; a < 3; ++a) if (pv[a] < lo[a] || pv[a] > hi[a]) { fprintf(stderr, "vertex outside leaf box (node %d slot %d axis %d: %g not in [%g, %g])\n", v, s, a, pv[a], lo[a], hi[a]); return false; }
                    }
                }
            }
        }
    }
    if (nodes != b.nNodes) { fprintf(stderr, "reached %ld of %d nodes\n", nodes, b.nNodes); return false; }
    for (int i = 0; i < b.n; ++i) if (seen[i] != 1) { fprintf(stderr, "primitive %d in %d leaves\n", i, seen[i]); return false; }
    return true;
}

struct RayStats { long nodes = 0, tris = 0, rays = 0; };
struct Hit { float t; int tri; };

static bool TriHitMT(const Built &b, size_t o, V3 ro, V3 rd, float tMax, float *tOut) {
    V3 p0{b.trisOut[o].x, b.trisOut[o].y, b.trisOut[o].z}, p1{b.trisOut[o + 1].x, b.trisOut[o + 1].y, b.trisOut[o + 1].z}, p2{b.trisOut[o + 2].x, b.trisOut[o + 2].y, b.trisOut[o + 2].z};
    V3 e1 = p1 - p0, e2 = p2 - p0, pv = cross(rd, e2);
    float det = dot(e1, pv);
    if (std::fabs(det) < 1e-20f) return false;
    float inv = 1.f / det;
    V3 tv = ro - p0;
    float u = dot(tv, pv) * inv;
    if (u < 0 || u > 1) return false;
    V3 qv = cross(tv, e1);
    float v = dot(rd, qv) * inv;
    if (v < 0 || u + v > 1) return false;
    float t = dot(e2, qv) * inv;
    if (t <= 1e-5f || t >= tMax) return false;
    *tOut = t;
    return true;
}

static void Visit(const Built &b, int node, V3 ro, V3 rd, V3 inv, unsigned octinv, Hit &hit, RayStats &st) {
    ++st.nodes;
    WNode w = Decode(b, node);
    int order[8]; int nh = 0;
    unsigned hitSlots = 0;
    for (int s = 0; s < 8; ++s) {
        if (!w.meta[s]) continue;
        float t0 = 0.f, t1 = hit.t;
        const float o3[3] = {ro.x, ro.y, ro.z}, i3[3] = {inv.x, inv.y, inv.z};
        for (int a = 0; a < 3; ++a) {
            float lo = w.p[a] + w.qlo[a][s] * w.scale[a], hi = w.p[a] + w.qhi[a][s] * w.scale[a];
            float ta = (lo - o3[a]) * i3[a], tb = (hi - o3[a]) * i3[a];
            if (ta > tb) std::swap(ta, tb);
            t0 = std::max(t0, ta); t1 = std::min(t1, tb * 1.0000003f);
        }
        if (t0 <= t1) hitSlots |= 1u << s;
    }
    // leaf children: tested right away (TraceRay order)
    for (int s = 0; s < 8; ++s) {
        if (!(hitSlots >> s & 1) || IsInner(w.meta[s])) continue;
        int count = (w.meta[s] >> 5) == 1 ? 1 : ((w.meta[s] >> 5) == 3 ? 2 : 3);
        int off = w.meta[s] & 31;
        for (int t = 0; t < count; ++t) {
            ++st.tris;
            if (b.boxes) continue;
            float th;
            size_t o = 3 * (size_t)(w.triBase + off + t);
            if (TriHitMT(b, o, ro, rd, hit.t, &th)) { hit.t = th; hit.tri = (int)(w.triBase + off + t); }
        }
    }
    // internal children in descending (slot ^ octinv)
    for (int key = 7; key >= 0; --key) {
        int s = key ^ (int)octinv;
        if ((hitSlots >> s & 1) && IsInner(w.meta[s])) order[nh++] = s;
    }
    for (int k = 0; k < nh; ++k) {
        int s = order[k];
        int child = w.childBase + __builtin_popcount(w.imask & ((1u << s) - 1));
        Visit(b, child, ro, rd, inv, octinv, hit, st);
    }
}
static Hit Trace(const Built &b, V3 ro, V3 rd, RayStats &st) {
    auto safe = [](float d) { return 1.f / (std::fabs(d) > 1e-20f ? d : std::copysign(1e-20f, d)); };
    V3 inv{safe(rd.x), safe(rd.y), safe(rd.z)};
    unsigned octinv = (rd.x < 0 ? 0u : 4u) | (rd.y < 0 ? 0u : 2u) | (rd.z < 0 ? 0u : 1u);
    Hit h{3.4e38f, -1};
    ++st.rays;
    Visit(b, 0, ro, rd, inv, octinv, h, st);
    return h;
}

static double SAHCost(const Built &b) {
    double root = HalfArea(b.nmin[0], b.nmax[0]), sum = 0;
    for (int v = 0; v < b.n - 1; ++v) sum += HalfArea(b.nmin[v], b.nmax[v]);
    return sum / root;
}

int main(int argc, char **argv) {
    if (argc < 4) { fprintf(stderr, "usage: bvh_build_host heightfield|soup|boxes size rounds [peraxis|cubic] [leafMax]\n"); return 2; }
    std::string scene = argv[1];
    int size = atoi(argv[2]), rounds = atoi(argv[3]);
    bool perAxis = argc > 4 && !strcmp(argv[4], "peraxis");
    int leafMaxArg = argc > 5 ? atoi(argv[5]) : 3;
    Built b;
    if (scene == "heightfield") MakeHeightfield(size, b.tris);
    else if (scene == "soup") MakeSoup(size, b.tris);
    else { MakeBoxes(size, b.tris); b.boxes = true; }
    if (!Build(b, rounds, perAxis, b.boxes ? 1 : leafMaxArg)) return 1;
    if (!Validate(b)) return 1;
    // ray distribution: camera rays of the bench scenes, then cosine-sampled bounces off the hit points (5 bounces)
    RayStats cam, bounce;
    std::mt19937 rng(5);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    const int R = b.boxes ? 96 : 160;
    float ext = b.boxes ? 50.f * std::cbrt(b.n / 100000.f) : 1.f;
    V3 eye = b.boxes ? V3{0, 0.9f * ext, 1.6f * ext} : V3{0, 1.6f, 2.6f};
    V3 fw = norm(V3{0, 0, 0} - eye), rt = norm(cross(fw, V3{0, 1, 0})), up = cross(rt, fw);
    float tanH = std::tan(0.5f * 45.f * 3.14159265f / 180.f);
    for (int y = 0; y < R; ++y)
        for (int x = 0; x < R; ++x) {
            float sx = (2 * (x + U(rng)) / R - 1) * tanH, sy = (1 - 2 * (y + U(rng)) / R) * tanH;
            V3 ro = eye, rd = norm(fw + rt * sx + up * sy);
            for (int depth = 0; depth <= 5; ++depth) {
                Hit h = Trace(b, ro, rd, depth == 0 ? cam : bounce);
                if (h.tri < 0 || b.boxes) break;
                size_t o = 3 * (size_t)h.tri;
                V3 p0{b.trisOut[o].x, b.trisOut[o].y, b.trisOut[o].z}, p1{b.trisOut[o + 1].x, b.trisOut[o + 1].y, b.trisOut[o + 1].z}, p2{b.trisOut[o + 2].x, b.trisOut[o + 2].y, b.trisOut[o + 2].z};
                V3 nrm = norm(cross(p1 - p0, p2 - p0));
                if (dot(nrm, rd) > 0) nrm = nrm * -1.f;
                V3 P = ro + rd * h.t + nrm * 1e-4f;
                float u1 = U(rng), u2 = U(rng), r = std::sqrt(u1), ph = 6.2831853f * u2;
                V3 t1 = norm(std::fabs(nrm.x) > 0.5f ? cross(nrm, V3{0, 1, 0}) : cross(nrm, V3{1, 0, 0})), t2 = cross(nrm, t1);
                rd = norm(t1 * (r * std::cos(ph)) + t2 * (r * std::sin(ph)) + nrm * std::sqrt(std::max(0.f, 1 - u1)));
                ro = P;
            }
        }
    printf("{\"scene\": \"%s\", \"prims\": %d, \"rounds\": %d, \"morton\": \"%s\", \"wide_nodes\": %d, \"depth\": %d, \"sah_internal_area_sum\": %.3f, "
           "\"camera\": {\"rays\": %ld, \"nodes_per_ray\": %.3f, \"tris_per_ray\": %.3f}, \"bounce\": {\"rays\": %ld, \"nodes_per_ray\": %.3f, \"tris_per_ray\": %.3f}}\n",
           scene.c_str(), b.n, rounds, perAxis ? "peraxis" : "cubic", b.nNodes, b.depth, SAHCost(b), cam.rays, (double)cam.nodes / cam.rays,
           (double)cam.tris / cam.rays, bounce.rays, bounce.rays ? (double)bounce.nodes / bounce.rays : 0., bounce.rays ? (double)bounce.tris / bounce.rays : 0.);
    return 0;
}
