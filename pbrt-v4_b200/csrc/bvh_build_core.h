// bvh_build_core.h -- per-element logic of the BVH builder, shared by the CUDA kernels (bvh_build.cu) and by the host harness
// that measures tree quality without a GPU (tests/tools/bvh_build_host.cpp).
//
// Reference counterparts (cpu/aggregates.cpp): Morton codes + radix sort of buildHLBVH (:389-417, 42-82), the SAH objective of
// buildRecursive (:192-387; cost = sum of child areas relative to the parent, leaves of <= maxPrimsInNode primitives).  The
// reference minimises that objective top-down with 12 buckets per split; here the same objective (sum of the surface areas of
// the internal nodes) is minimised bottom-up: a Karras radix tree over 63-bit Morton keys gives the initial topology, and
// treelets of 7 subtrees are restructured to their SAH-optimal binary topology by exhaustive dynamic programming over the 2^7
// subsets (Karras & Aila 2013, "Fast parallel construction of high-quality bounding volume hierarchies"), several rounds with a
// growing minimum subtree size.  The hierarchy only decides which triangles are tested; the hit record does not depend on it.
#pragma once
#include <cuda_runtime.h>
#include "pmath.h"

namespace bpt {
namespace bvhb {

#if defined(__CUDA_ARCH__)
// tree arrays are read and written by different thread blocks inside one bottom-up kernel: bypass the (incoherent) L1
#define BVHB_LD(p) __ldcg(p)
#define BVHB_ST(p, v) __stcg((p), (v))
#define BVHB_ATOMIC_ADD(p, v) atomicAdd((p), (v))
BPT_HD int Clz64(uint64_t x) { return __clzll((long long)x); }
BPT_HD int Clz32(uint32_t x) { return __clz((int)x); }
BPT_HD int Ctz32(uint32_t x) { return __ffs((int)x) - 1; }
BPT_HD int Popc32(uint32_t x) { return __popc(x); }
#else
#define BVHB_LD(p) (*(p))
#define BVHB_ST(p, v) (*(p) = (v))
inline int HostFetchAdd(int *p, int v) { int o = *p; *p += v; return o; }
#define BVHB_ATOMIC_ADD(p, v) ::bpt::bvhb::HostFetchAdd((p), (v))
BPT_HD int Clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }
BPT_HD int Clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
BPT_HD int Ctz32(uint32_t x) { return __builtin_ctz(x); }
BPT_HD int Popc32(uint32_t x) { return __builtin_popcount(x); }
#endif

// Binary tree over n leaves.  Node ids: internal 0 .. n-2 (root = 0), leaf j (position j of the sorted order) = (n-1) + j.
struct Tree {
    int n;
    int *left, *right, *parent;   // left/right: n-1 entries; parent: 2n-1
    int *cnt;                     // leaves below each node, 2n-1 entries
    float4 *nmin, *nmax;          // 2n-1 boxes
};

BPT_HD float HalfArea(float4 mn, float4 mx) {
    float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
    return dx * dy + dy * dz + dz * dx;
}
BPT_HD float4 Min4(float4 a, float4 b) { return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), 0.f); }
BPT_HD float4 Max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), 0.f); }

BPT_HD void TriBounds(float4 a, float4 b, float4 c, float4 *lo, float4 *hi) {
    *lo = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0.f);
    *hi = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0.f);
}

BPT_HD uint64_t Spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
// 63-bit Morton key of a primitive centroid on a grid of CUBIC cells (edge = largest centroid extent / 2^21).  Cubic cells matter:
// with per-axis normalisation (as the reference's HLBVH does, aggregates.cpp:404-410) a flat scene -- a height field -- gets
// its thin axis split as early and as often as the long ones, which produces large, heavily overlapping sibling boxes.
BPT_HD uint64_t MortonKey(float4 lo, float4 hi, const float *cb /* centroid bounds: min xyz, max xyz */) {
    float ext = fmaxf(cb[3] - cb[0], fmaxf(cb[4] - cb[1], cb[5] - cb[2]));
    float scale = ext > 0 ? 2097152.f / ext : 0.f;
    float ctr[3] = {0.5f * lo.x + 0.5f * hi.x, 0.5f * lo.y + 0.5f * hi.y, 0.5f * lo.z + 0.5f * hi.z};
    uint64_t q[3];
    for (int k = 0; k < 3; ++k) {
        float t = (ctr[k] - cb[k]) * scale;
        t = fminf(fmaxf(t, 0.f), 2097151.f);
        q[k] = (uint64_t)t;
    }
    return (Spread21(q[0]) << 2) | (Spread21(q[1]) << 1) | Spread21(q[2]);
}

// ---- Karras 2012 binary radix tree, one call per internal node i
BPT_HD int Delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + Clz32((uint32_t)(i ^ j));
    return Clz64(a ^ b);
}
BPT_HD void RadixTreeNode(const uint64_t *keys, int n, int i, int *lc, int *rc) {
    int d = (Delta(keys, n, i, i + 1) - Delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = Delta(keys, n, i, i - d);
    int lmax = 2;
    while (Delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (Delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = Delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (Delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? d : 0);
    int lo = i < j ? i : j, hi = i < j ? j : i;
    *lc = (lo == gamma) ? (n - 1) + gamma : gamma;
    *rc = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
}

// ---- treelet restructuring.  The treelet rooted at `root` is grown to (up to) 7 leaves by repeatedly opening the leaf with the
// largest surface area; the optimal binary topology over those leaves -- the one minimising the summed surface area of its
// internal nodes -- is found by dynamic programming over all subsets, and the treelet's internal nodes are rewired in place
// when that lowers the cost.  Returns true if the tree was changed.  Called for a node only after both its subtrees are final.
constexpr int kTreeletLeaves = 7;
// copt / popt: the dynamic-programming tables over the 2^7 leaf subsets, element s at [s * STRIDE].  The CUDA kernel keeps them in
// shared memory, one column per thread (STRIDE = block size: conflict-free, and no 1.1 KB-per-thread local-memory frame -- with the
// tables in local memory the restructuring pass was memory-bound: 3.2 ms per round for 1 M triangles, profiles/r02g); the host
// harness passes plain arrays (STRIDE = 1).  Subset areas are recomputed from the leaf boxes (registers) where they are needed.
struct Treelet {
    int leaf[kTreeletLeaves], inner[kTreeletLeaves - 1];
    float4 lmn[kTreeletLeaves], lmx[kTreeletLeaves];
    int nl;
    float curCost;
};
// grow the treelet rooted at `root` to (up to) 7 leaves by repeatedly opening the leaf with the largest surface area
BPT_HD void FormTreelet(const Tree &T, int root, Treelet &t) {
    const int nInternalTotal = T.n - 1;
    float larea[kTreeletLeaves];
    int nl = 0, ni = 0;
    auto addLeaf = [&](int id) {
        t.leaf[nl] = id; t.lmn[nl] = BVHB_LD(T.nmin + id); t.lmx[nl] = BVHB_LD(T.nmax + id);
        larea[nl] = id < nInternalTotal ? HalfArea(t.lmn[nl], t.lmx[nl]) : -1.f;   // real leaves cannot be opened
        ++nl;
    };
    t.inner[ni++] = root;
    addLeaf(BVHB_LD(T.left + root));
    addLeaf(BVHB_LD(T.right + root));
    float curCost = HalfArea(BVHB_LD(T.nmin + root), BVHB_LD(T.nmax + root));
    while (nl < kTreeletLeaves) {
        int best = -1; float bestA = -1.f;
        for (int k = 0; k < nl; ++k) if (larea[k] > bestA) { bestA = larea[k]; best = k; }
        if (best < 0) break;
        int b = t.leaf[best];
        curCost += larea[best];
        t.inner[ni++] = b;
        int l = BVHB_LD(T.left + b), r = BVHB_LD(T.right + b);
        t.leaf[best] = l; t.lmn[best] = BVHB_LD(T.nmin + l); t.lmx[best] = BVHB_LD(T.nmax + l);
        larea[best] = l < nInternalTotal ? HalfArea(t.lmn[best], t.lmx[best]) : -1.f;
        addLeaf(r);
    }
    t.nl = nl; t.curCost = curCost;
}
BPT_HD float TreeletSubsetArea(const Treelet &t, unsigned sub) {
    float4 mn = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 0), mx = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, 0);
#pragma unroll
    for (int k = 0; k < kTreeletLeaves; ++k) if (k < t.nl && (sub >> k & 1)) { mn = Min4(mn, t.lmn[k]); mx = Max4(mx, t.lmx[k]); }
    return HalfArea(mn, mx);
}
// best split of subset s given the optimal costs of all smaller subsets: copt[s] = area(s) + min over partitions
template <int STRIDE>
BPT_HD void TreeletSolveSubset(const Treelet &t, unsigned s, float *copt, unsigned char *popt) {
    float best = 3.4e38f; unsigned bp = 0;
    unsigned delta = (s - 1) & s;          // s without its lowest bit: partitions enumerated once (lowest bit stays in s ^ p)
    unsigned p = (0u - delta) & s;
    do {
        float c = copt[p * STRIDE] + copt[(s ^ p) * STRIDE];
        if (c < best) { best = c; bp = p; }
        p = (p - delta) & s;
    } while (p != 0);
    copt[s * STRIDE] = TreeletSubsetArea(t, s) + best;
    popt[s * STRIDE] = (unsigned char)bp;
}
// rewire: internal ids are reused, root keeps its id (and its box and count)
template <int STRIDE>
BPT_HD void RewireTreelet(const Tree &T, int root, const Treelet &t, const unsigned char *popt) {
    const int nl = t.nl;
    const unsigned full = (1u << nl) - 1u;
    unsigned stackS[kTreeletLeaves]; int stackId[kTreeletLeaves];
    int sp = 0, nextI = 1;
    stackS[sp] = full; stackId[sp] = root; ++sp;
    while (sp > 0) {
        --sp;
        unsigned s = stackS[sp]; int id = stackId[sp];
        unsigned part[2] = {popt[s * STRIDE], s ^ (unsigned)popt[s * STRIDE]};
        int cid[2];
        for (int c = 0; c < 2; ++c) {
            unsigned ps = part[c];
            if ((ps & (ps - 1)) == 0) cid[c] = t.leaf[Ctz32(ps)];
            else {
                cid[c] = t.inner[nextI++];
                float4 mn = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 0), mx = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, 0);
                int count = 0;
                for (int k = 0; k < nl; ++k) if (ps >> k & 1) { mn = Min4(mn, t.lmn[k]); mx = Max4(mx, t.lmx[k]); count += BVHB_LD(T.cnt + t.leaf[k]); }
                BVHB_ST(T.nmin + cid[c], mn); BVHB_ST(T.nmax + cid[c], mx); BVHB_ST(T.cnt + cid[c], count);
                stackS[sp] = ps; stackId[sp] = cid[c]; ++sp;
            }
            BVHB_ST(T.parent + cid[c], id);
        }
        BVHB_ST(T.left + id, cid[0]); BVHB_ST(T.right + id, cid[1]);
    }
}
template <int STRIDE>
BPT_HD bool OptimizeTreelet(const Tree &T, int root, float *copt, unsigned char *popt) {
    Treelet t;
    FormTreelet(T, root, t);
    const int nl = t.nl;
    if (nl < 4) return false;   // 3 leaves or fewer: nothing a different topology could gain that the next level up does not see
    const unsigned full = (1u << nl) - 1u;
    for (int k = 0; k < nl; ++k) copt[(1u << k) * STRIDE] = 0.f;
    for (int sz = 2; sz <= nl; ++sz)
        for (unsigned s = 1; s <= full; ++s)
            if (Popc32(s) == sz) TreeletSolveSubset<STRIDE>(t, s, copt, popt);
    if (!(copt[full * STRIDE] < t.curCost * 0.9999f)) return false;
    RewireTreelet<STRIDE>(T, root, t, popt);
    return true;
}
// plain-array form (host harness)
inline bool OptimizeTreelet(const Tree &T, int root) {
    float copt[1 << kTreeletLeaves];
    unsigned char popt[1 << kTreeletLeaves];
    return OptimizeTreelet<1>(T, root, copt, popt);
}

// ---- collapse of the binary tree into the 8-wide compressed nodes of bvh.cuh, one call per wide node
struct CollapseItem { int bnode; int wnode; };

struct CollapseCtx {
    Tree T;
    const int *sortedVals;        // leaf position -> original primitive index
    const float4 *trisIn;         // 3 float4 per primitive, original order
    float4 *nodesOut, *trisOut;
    int *nodeCounter, *triCounter;
    CollapseItem *nextItems; int *nextCount;
    int leafMax;
};

// leaves of the (at most 3-leaf) subtree e, in tree order
BPT_HD int GatherLeaves(const Tree &T, int e, int *out) {
    int stack[8]; int sp = 0, no = 0;
    stack[sp++] = e;
    while (sp > 0) {
        int v = stack[--sp];
        if (v >= T.n - 1) out[no++] = v - (T.n - 1);
        else { stack[sp++] = T.right[v]; stack[sp++] = T.left[v]; }
    }
    return no;
}

BPT_HD void CollapseOne(const CollapseCtx &C, CollapseItem it) {
    const Tree &T = C.T;
    const int n = T.n;
    int ent[8];
    int ne = 1;
    ent[0] = it.bnode;
    // greedy expansion by surface area
    while (ne < 8) {
        int best = -1;
        float bestA = -1.f;
        for (int k = 0; k < ne; ++k) {
            if (ent[k] >= n - 1) continue;   // leaf
            float a = HalfArea(T.nmin[ent[k]], T.nmax[ent[k]]);
            if (a > bestA) { bestA = a; best = k; }
        }
        if (best < 0) break;
        int b = ent[best];
        ent[best] = T.left[b];
        ent[ne++] = T.right[b];
    }
    // a lone root leaf (single-triangle scene) stays a single leaf entry
    float4 wmn = T.nmin[it.bnode], wmx = T.nmax[it.bnode];
    float ctrx = 0.5f * (wmn.x + wmx.x), ctry = 0.5f * (wmn.y + wmx.y), ctrz = 0.5f * (wmn.z + wmx.z);
    int cnt[8];
    float4 emn[8], emx[8];
    for (int k = 0; k < ne; ++k) {
        int e = ent[k];
        cnt[k] = T.cnt[e];
        emn[k] = T.nmin[e]; emx[k] = T.nmax[e];
    }
    // slot assignment: greedy maximisation of dot(centroid offset, slot direction)
    int slotOf[8];
    bool slotUsed[8] = {false, false, false, false, false, false, false, false};
    bool entDone[8] = {false, false, false, false, false, false, false, false};
    for (int round = 0; round < ne; ++round) {
        float bestC = -3.4e38f; int bk = -1, bs = -1;
        for (int k = 0; k < ne; ++k) {
            if (entDone[k]) continue;
            float cx = 0.5f * (emn[k].x + emx[k].x) - ctrx, cy = 0.5f * (emn[k].y + emx[k].y) - ctry,
                  cz = 0.5f * (emn[k].z + emx[k].z) - ctrz;
            for (int s = 0; s < 8; ++s) {
                if (slotUsed[s]) continue;
                float c = ((s & 4) ? cx : -cx) + ((s & 2) ? cy : -cy) + ((s & 1) ? cz : -cz);
                if (c > bestC) { bestC = c; bk = k; bs = s; }
            }
        }
        slotOf[bk] = bs; slotUsed[bs] = true; entDone[bk] = true;
    }
    // quantisation frame
    float ext[3] = {wmx.x - wmn.x, wmx.y - wmn.y, wmx.z - wmn.z};
    int ex[3];
    float scale[3];
    for (int a = 0; a < 3; ++a) {
        int e = -100;
        if (ext[a] > 0) {
            e = (int)ceilf(log2f(ext[a] / 255.f));
            if (e < -100) e = -100;
            while (ldexpf(255.f, e) < ext[a]) ++e;
        }
        ex[a] = e;
        scale[a] = ldexpf(1.f, e);
    }
    int nInternal = 0, nTris = 0;
    for (int k = 0; k < ne; ++k) { if (cnt[k] > C.leafMax) ++nInternal; else nTris += cnt[k]; }
    int childBase = nInternal ? BVHB_ATOMIC_ADD(C.nodeCounter, nInternal) : 0;
    int triBase = nTris ? BVHB_ATOMIC_ADD(C.triCounter, nTris) : 0;
    unsigned imask = 0;
    unsigned char meta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned char qlo[3][8], qhi[3][8];
    for (int s = 0; s < 8; ++s) for (int a = 0; a < 3; ++a) { qlo[a][s] = 0; qhi[a][s] = 0; }
    // children in slot order
    int triOff = 0, rel = 0;
    for (int s = 0; s < 8; ++s) {
        int k = -1;
        for (int kk = 0; kk < ne; ++kk) if (slotOf[kk] == s) k = kk;
        if (k < 0) continue;
        float lo3[3] = {emn[k].x, emn[k].y, emn[k].z}, hi3[3] = {emx[k].x, emx[k].y, emx[k].z};
        float org[3] = {wmn.x, wmn.y, wmn.z};
        for (int a = 0; a < 3; ++a) {
            float ql = floorf((lo3[a] - org[a]) / scale[a]);
            float qh = ceilf((hi3[a] - org[a]) / scale[a]);
            ql = fminf(fmaxf(ql, 0.f), 255.f); qh = fminf(fmaxf(qh, 0.f), 255.f);
            while (ql > 0.f && org[a] + ql * scale[a] > lo3[a]) ql -= 1.f;
            while (qh < 255.f && org[a] + qh * scale[a] < hi3[a]) qh += 1.f;
            qlo[a][s] = (unsigned char)ql; qhi[a][s] = (unsigned char)qh;
        }
        if (cnt[k] > C.leafMax) {
            imask |= 1u << s;
            meta[s] = (unsigned char)((1u << 5) | (24 + s));
            int pos = BVHB_ATOMIC_ADD(C.nextCount, 1);
            C.nextItems[pos].bnode = ent[k];
            C.nextItems[pos].wnode = childBase + rel;
            ++rel;
        } else {
            unsigned unary = cnt[k] == 1 ? 1u : (cnt[k] == 2 ? 3u : 7u);
            meta[s] = (unsigned char)((unary << 5) | (unsigned)triOff);
            int leaves[4];
            int nLeaves = GatherLeaves(T, ent[k], leaves);
            for (int t = 0; t < nLeaves; ++t) {
                int orig = C.sortedVals[leaves[t]];
                float4 a = C.trisIn[3 * (size_t)orig], b = C.trisIn[3 * (size_t)orig + 1], c = C.trisIn[3 * (size_t)orig + 2];
                // leaf-order record: p0.w = original triangle id, p1.w = class bits (material tag | emissive << 8),
                // p2.w = mesh id  (input pool: p0.w = mesh id, p1.w = triangle index in mesh, p2.w = class bits)
                float meshId = a.w;
                a.w = b2f((uint32_t)orig);
                // degenerate test of IntersectTriangle (shapes.cpp:172), same float expression, evaluated once per triangle
                unsigned cls = f2b(c.w);
                {
                    Vec3 e20(c.x - a.x, c.y - a.y, c.z - a.z), e10(b.x - a.x, b.y - a.y, b.z - a.z);
                    if (LengthSquared(Cross(e20, e10)) == 0) cls |= 0x80000000u;
                }
                b.w = b2f(cls);
                c.w = meshId;
                size_t o = 3 * (size_t)(triBase + triOff + t);
                C.trisOut[o] = a; C.trisOut[o + 1] = b; C.trisOut[o + 2] = c;
            }
            triOff += cnt[k];
        }
    }
    // NOTE: `rel` order above follows ascending slot order, which is what traversal's popc(imask & below) expects.
    auto pack4 = [](const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24); };
    float4 n0, n1, n2, n3, n4;
    n0.x = wmn.x; n0.y = wmn.y; n0.z = wmn.z;
    n0.w = b2f((unsigned)(ex[0] + 127) | ((unsigned)(ex[1] + 127) << 8) | ((unsigned)(ex[2] + 127) << 16) | (imask << 24));
    n1.x = b2f((unsigned)childBase); n1.y = b2f((unsigned)triBase);
    n1.z = b2f(pack4(meta)); n1.w = b2f(pack4(meta + 4));
    n2.x = b2f(pack4(qlo[0])); n2.y = b2f(pack4(qlo[0] + 4));
    n2.z = b2f(pack4(qhi[0])); n2.w = b2f(pack4(qhi[0] + 4));
    n3.x = b2f(pack4(qlo[1])); n3.y = b2f(pack4(qlo[1] + 4));
    n3.z = b2f(pack4(qhi[1])); n3.w = b2f(pack4(qhi[1] + 4));
    n4.x = b2f(pack4(qlo[2])); n4.y = b2f(pack4(qlo[2] + 4));
    n4.z = b2f(pack4(qhi[2])); n4.w = b2f(pack4(qhi[2] + 4));
    float4 *np = C.nodesOut + (size_t)it.wnode * 5;
    np[0] = n0; np[1] = n1; np[2] = n2; np[3] = n3; np[4] = n4;
}

}  // namespace bvhb
}  // namespace bpt
