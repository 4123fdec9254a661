// bvh_build.cu -- GPU LBVH build and collapse into the 8-wide compressed BVH of bvh.cuh.
//
// Pipeline (all on the device, one stream):
//   1. triangle bounds + centroid bounds            (reference analogue: BVHAggregate ctor, cpu/aggregates.cpp:140-190)
//   2. 63-bit Morton codes of centroids             (reference HLBVH uses 30-bit codes, cpu/aggregates.cpp:389-417)
//   3. radix sort (radix_sort.cuh)                  (cpu/aggregates.cpp:42-82)
//   4. Karras-style binary radix tree over the sorted keys, one thread per internal node
//   5. bottom-up AABB refit with per-node arrival counters
//   6. top-down, level-synchronous collapse into 8-wide nodes: greedy surface-area expansion to <= 8 children,
//      subtrees of <= 3 triangles become leaf children, octant slot assignment, 8-bit box quantisation
// The hierarchy only decides which triangles are tested; the hit record does not depend on it.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>
#include "bvh_build.h"
#include "radix_sort.cuh"
#include "pmath.h"

#ifndef BPT_LEAF_MAX
#define BPT_LEAF_MAX 3   // subtrees of at most this many triangles become one leaf child (<= 3: unary count in 3 meta bits)
#endif
namespace bpt {

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, 256, "%s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return false; } } while (0)

__device__ __forceinline__ void atomicMinF(float *addr, float v) {
    // works for any sign via ordered-int trick
    if (v >= 0) atomicMin((int *)addr, __float_as_int(v)); else atomicMax((unsigned *)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxF(float *addr, float v) {
    if (v >= 0) atomicMax((int *)addr, __float_as_int(v)); else atomicMin((unsigned *)addr, __float_as_uint(v));
}

__global__ void centroid_bounds_kernel(const float4 *__restrict__ tris, int n, float *cb /* 6 floats: min xyz, max xyz */) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    if (i < n) {
        float4 a = tris[3 * (size_t)i], b = tris[3 * (size_t)i + 1], c = tris[3 * (size_t)i + 2];
        float lo[3] = {fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z))};
        float hi[3] = {fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z))};
        for (int k = 0; k < 3; ++k) { float ctr = 0.5f * lo[k] + 0.5f * hi[k]; mn[k] = ctr; mx[k] = ctr; }
    }
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 3; ++k) { atomicMinF(&cb[k], mn[k]); atomicMaxF(&cb[3 + k], mx[k]); }
    }
}

__device__ __forceinline__ uint64_t Spread21(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void morton_kernel(const float4 *__restrict__ tris, int n, const float *__restrict__ cb, uint64_t *keys, int *vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = tris[3 * (size_t)i], b = tris[3 * (size_t)i + 1], c = tris[3 * (size_t)i + 2];
    float ctr[3] = {0.5f * fminf(a.x, fminf(b.x, c.x)) + 0.5f * fmaxf(a.x, fmaxf(b.x, c.x)),
                    0.5f * fminf(a.y, fminf(b.y, c.y)) + 0.5f * fmaxf(a.y, fmaxf(b.y, c.y)),
                    0.5f * fminf(a.z, fminf(b.z, c.z)) + 0.5f * fmaxf(a.z, fmaxf(b.z, c.z))};
    uint64_t q[3];
    for (int k = 0; k < 3; ++k) {
        float ext = cb[3 + k] - cb[k];
        float t = ext > 0 ? (ctr[k] - cb[k]) / ext : 0.f;
        t = fminf(fmaxf(t, 0.f), 1.f);
        q[k] = (uint64_t)fminf(t * 2097152.f, 2097151.f);
    }
    keys[i] = (Spread21(q[0]) << 2) | (Spread21(q[1]) << 1) | Spread21(q[2]);
    vals[i] = i;
}

// ---- binary radix tree.  Node ids: internal 0..n-2, leaf j -> (n-1)+j.
__device__ __forceinline__ int Delta(const uint64_t *keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void hierarchy_kernel(const uint64_t *__restrict__ keys, int n, int *left, int *right, int *parent, int *first, int *last) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
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
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int lc = (lo == gamma) ? (n - 1) + gamma : gamma;
    int rc = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
    left[i] = lc; right[i] = rc;
    parent[lc] = i; parent[rc] = i;
    first[i] = lo; last[i] = hi;
    if (i == 0) parent[0] = -1;
}

__global__ void fit_kernel(const float4 *__restrict__ tris, const int *__restrict__ sortedVals, int n, const int *__restrict__ left,
                           const int *__restrict__ right, const int *__restrict__ parent, float4 *nmin, float4 *nmax, int *flags) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    int tri = sortedVals[j];
    float4 a = tris[3 * (size_t)tri], b = tris[3 * (size_t)tri + 1], c = tris[3 * (size_t)tri + 2];
    float4 lo = make_float4(fminf(a.x, fminf(b.x, c.x)), fminf(a.y, fminf(b.y, c.y)), fminf(a.z, fminf(b.z, c.z)), 0);
    float4 hi = make_float4(fmaxf(a.x, fmaxf(b.x, c.x)), fmaxf(a.y, fmaxf(b.y, c.y)), fmaxf(a.z, fmaxf(b.z, c.z)), 0);
    int id = (n - 1) + j;
    nmin[id] = lo; nmax[id] = hi;
    if (n == 1) return;
    __threadfence();
    int p = parent[id];
    while (p >= 0) {
        if (atomicAdd(&flags[p], 1) == 0) return;   // first arrival: the sibling's thread finishes this node
        int l = left[p], r = right[p];
        volatile float4 *vmin = nmin; volatile float4 *vmax = nmax;
        float4 lmn, lmx, rmn, rmx;
        lmn.x = vmin[l].x; lmn.y = vmin[l].y; lmn.z = vmin[l].z;
        lmx.x = vmax[l].x; lmx.y = vmax[l].y; lmx.z = vmax[l].z;
        rmn.x = vmin[r].x; rmn.y = vmin[r].y; rmn.z = vmin[r].z;
        rmx.x = vmax[r].x; rmx.y = vmax[r].y; rmx.z = vmax[r].z;
        nmin[p] = make_float4(fminf(lmn.x, rmn.x), fminf(lmn.y, rmn.y), fminf(lmn.z, rmn.z), 0);
        nmax[p] = make_float4(fmaxf(lmx.x, rmx.x), fmaxf(lmx.y, rmx.y), fmaxf(lmx.z, rmx.z), 0);
        __threadfence();
        p = parent[p];
    }
}

// ---- collapse
struct CollapseItem { int bnode; int wnode; };

__device__ __forceinline__ float Area(float4 mn, float4 mx) {
    float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
    return dx * dy + dy * dz + dz * dx;
}

__global__ void collapse_kernel(const CollapseItem *__restrict__ items, int nItems, int n, const int *__restrict__ left,
                                const int *__restrict__ right, const int *__restrict__ first, const int *__restrict__ last,
                                const float4 *__restrict__ nmin, const float4 *__restrict__ nmax, const int *__restrict__ sortedVals,
                                const float4 *__restrict__ trisIn, float4 *nodesOut, float4 *trisOut, int *nodeCounter,
                                int *triCounter, CollapseItem *nextItems, int *nextCount, int leafMax) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nItems) return;
    CollapseItem it = items[idx];
    int ent[8];
    int ne = 1;
    ent[0] = it.bnode;
    // greedy expansion by surface area
    while (ne < 8) {
        int best = -1;
        float bestA = -1.f;
        for (int k = 0; k < ne; ++k) {
            if (ent[k] >= n - 1) continue;   // leaf
            float a = Area(nmin[ent[k]], nmax[ent[k]]);
            if (a > bestA) { bestA = a; best = k; }
        }
        if (best < 0) break;
        int b = ent[best];
        ent[best] = left[b];
        ent[ne++] = right[b];
    }
    // a lone root leaf (single-triangle scene) stays a single leaf entry
    float4 wmn = nmin[it.bnode], wmx = nmax[it.bnode];
    float ctrx = 0.5f * (wmn.x + wmx.x), ctry = 0.5f * (wmn.y + wmx.y), ctrz = 0.5f * (wmn.z + wmx.z);
    // entry info
    int cnt[8], fst[8];
    float4 emn[8], emx[8];
    for (int k = 0; k < ne; ++k) {
        int e = ent[k];
        if (e >= n - 1) { fst[k] = e - (n - 1); cnt[k] = 1; }
        else { fst[k] = first[e]; cnt[k] = last[e] - first[e] + 1; }
        emn[k] = nmin[e]; emx[k] = nmax[e];
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
    for (int k = 0; k < ne; ++k) { if (cnt[k] > leafMax) ++nInternal; else nTris += cnt[k]; }
    int childBase = nInternal ? atomicAdd(nodeCounter, nInternal) : 0;
    int triBase = nTris ? atomicAdd(triCounter, nTris) : 0;
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
        if (cnt[k] > leafMax) {
            imask |= 1u << s;
            meta[s] = (unsigned char)((1u << 5) | (24 + s));
            int pos = atomicAdd(nextCount, 1);
            nextItems[pos].bnode = ent[k];
            nextItems[pos].wnode = childBase + rel;
            ++rel;
        } else {
            unsigned unary = cnt[k] == 1 ? 1u : (cnt[k] == 2 ? 3u : 7u);
            meta[s] = (unsigned char)((unary << 5) | (unsigned)triOff);
            for (int t = 0; t < cnt[k]; ++t) {
                int orig = sortedVals[fst[k] + t];
                float4 a = trisIn[3 * (size_t)orig], b = trisIn[3 * (size_t)orig + 1], c = trisIn[3 * (size_t)orig + 2];
                // leaf-order record: p0.w = original triangle id, p1.w = class bits (material tag | emissive << 8),
                // p2.w = mesh id  (input pool: p0.w = mesh id, p1.w = triangle index in mesh, p2.w = class bits)
                float meshId = a.w;
                a.w = __int_as_float(orig);
                // degenerate test of IntersectTriangle (shapes.cpp:172), same float expression, evaluated once per triangle
                unsigned cls = __float_as_uint(c.w);
                {
                    Vec3 e20(c.x - a.x, c.y - a.y, c.z - a.z), e10(b.x - a.x, b.y - a.y, b.z - a.z);
                    if (LengthSquared(Cross(e20, e10)) == 0) cls |= 0x80000000u;
                }
                b.w = __uint_as_float(cls);
                c.w = meshId;
                size_t o = 3 * (size_t)(triBase + triOff + t);
                trisOut[o] = a; trisOut[o + 1] = b; trisOut[o + 2] = c;
            }
            triOff += cnt[k];
        }
    }
    // NOTE: `rel` order above follows ascending slot order, which is what traversal's popc(imask & below) expects.
    auto pack4 = [](const unsigned char *p) { return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24); };
    float4 n0, n1, n2, n3, n4;
    n0.x = wmn.x; n0.y = wmn.y; n0.z = wmn.z;
    n0.w = __uint_as_float((unsigned)(ex[0] + 127) | ((unsigned)(ex[1] + 127) << 8) | ((unsigned)(ex[2] + 127) << 16) | (imask << 24));
    n1.x = __uint_as_float((unsigned)childBase); n1.y = __uint_as_float((unsigned)triBase);
    n1.z = __uint_as_float(pack4(meta)); n1.w = __uint_as_float(pack4(meta + 4));
    n2.x = __uint_as_float(pack4(qlo[0])); n2.y = __uint_as_float(pack4(qlo[0] + 4));
    n2.z = __uint_as_float(pack4(qhi[0])); n2.w = __uint_as_float(pack4(qhi[0] + 4));
    n3.x = __uint_as_float(pack4(qlo[1])); n3.y = __uint_as_float(pack4(qlo[1] + 4));
    n3.z = __uint_as_float(pack4(qhi[1])); n3.w = __uint_as_float(pack4(qhi[1] + 4));
    n4.x = __uint_as_float(pack4(qlo[2])); n4.y = __uint_as_float(pack4(qlo[2] + 4));
    n4.z = __uint_as_float(pack4(qhi[2])); n4.w = __uint_as_float(pack4(qhi[2] + 4));
    float4 *np = nodesOut + (size_t)it.wnode * 5;
    np[0] = n0; np[1] = n1; np[2] = n2; np[3] = n3; np[4] = n4;
}

template <typename T>
static bool dalloc(T **p, size_t n, char *err, cudaStream_t st = nullptr) {
    // stream-ordered allocation: the device pool (release threshold raised in bpt_create) recycles the builder's scratch and
    // output buffers across re-commits instead of paying cudaMalloc / cudaFree (device-wide syncs) every time
    cudaError_t e = cudaMallocAsync((void **)p, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e != cudaSuccess) { snprintf(err, 256, "cudaMallocAsync(%zu): %s", n * sizeof(T), cudaGetErrorString(e)); return false; }
    return true;
}

bool BuildWideBVH(const float4 *trisDev, int n, WideBVH *out, cudaStream_t stream, char *err, int leafMax) {
    if (leafMax <= 0 || leafMax > 3) leafMax = BPT_LEAF_MAX;
    out->nodes = nullptr; out->tris = nullptr; out->nNodes = 0; out->nTris = n; out->depth = 0;
    if (n <= 0) { snprintf(err, 256, "BuildWideBVH: empty triangle set"); return false; }
    const int B = 256;
    float *cb = nullptr;
    uint64_t *keys = nullptr, *keysTmp = nullptr;
    int *vals = nullptr, *valsTmp = nullptr;
    unsigned *hist = nullptr;
    int *left = nullptr, *right = nullptr, *parent = nullptr, *first = nullptr, *last = nullptr, *flags = nullptr;
    float4 *nmin = nullptr, *nmax = nullptr;
    CollapseItem *q0 = nullptr, *q1 = nullptr;
    int *counters = nullptr;   // [0]=nodeCounter [1]=triCounter [2]=nextCount
    int nTiles = (n + RS_TILE - 1) / RS_TILE;
    bool ok = dalloc(&cb, 6, err, stream) && dalloc(&keys, n, err, stream) && dalloc(&keysTmp, n, err, stream) && dalloc(&vals, n, err, stream) &&
              dalloc(&valsTmp, n, err, stream) && dalloc(&hist, 256 * (size_t)nTiles, err, stream) && dalloc(&left, n, err, stream) &&
              dalloc(&right, n, err, stream) && dalloc(&parent, 2 * (size_t)n, err, stream) && dalloc(&first, n, err, stream) && dalloc(&last, n, err, stream) &&
              dalloc(&flags, n, err, stream) && dalloc(&nmin, 2 * (size_t)n, err, stream) && dalloc(&nmax, 2 * (size_t)n, err, stream) &&
              dalloc(&q0, n, err, stream) && dalloc(&q1, n, err, stream) && dalloc(&counters, 4, err, stream) &&
              dalloc(&out->nodes, 5 * (size_t)n, err, stream) && dalloc(&out->tris, 3 * (size_t)n, err, stream);
    auto cleanup = [&]() {
        void *scratch[] = {cb, keys, keysTmp, vals, valsTmp, hist, left, right, parent, first, last, flags, nmin, nmax, q0, q1, counters};
        for (void *q : scratch) if (q) cudaFreeAsync(q, stream);
    };
    if (!ok) { cleanup(); return false; }
    float cbInit[6] = {3.4e38f, 3.4e38f, 3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f};
    CK(cudaMemcpyAsync(cb, cbInit, sizeof(cbInit), cudaMemcpyHostToDevice, stream));
    centroid_bounds_kernel<<<(n + B - 1) / B, B, 0, stream>>>(trisDev, n, cb);
    morton_kernel<<<(n + B - 1) / B, B, 0, stream>>>(trisDev, n, cb, keys, vals);
    bool flipped = RadixSortPairs<uint64_t, int>(keys, keysTmp, vals, valsTmp, n, 0, 64, hist, stream);
    uint64_t *skeys = flipped ? keysTmp : keys;
    int *svals = flipped ? valsTmp : vals;
    CK(cudaMemsetAsync(flags, 0, sizeof(int) * n, stream));
    if (n > 1) hierarchy_kernel<<<(n - 1 + B - 1) / B, B, 0, stream>>>(skeys, n, left, right, parent, first, last);
    fit_kernel<<<(n + B - 1) / B, B, 0, stream>>>(trisDev, svals, n, left, right, parent, nmin, nmax, flags);
    // collapse, level by level
    int hostCounters[4] = {1, 0, 0, 0};   // root wide node is node 0
    CK(cudaMemcpyAsync(counters, hostCounters, sizeof(hostCounters), cudaMemcpyHostToDevice, stream));
    CollapseItem root{n > 1 ? 0 : 0 /* leaf id (n-1)+0 == 0 when n == 1 */, 0};
    CK(cudaMemcpyAsync(q0, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
    int nItems = 1, depth = 0;
    CollapseItem *cur = q0, *nxt = q1;
    while (nItems > 0) {
        ++depth;
        CK(cudaMemsetAsync(counters + 2, 0, sizeof(int), stream));
        collapse_kernel<<<(nItems + 63) / 64, 64, 0, stream>>>(cur, nItems, n, left, right, first, last, nmin, nmax, svals, trisDev,
                                                               out->nodes, out->tris, counters, counters + 1, nxt, counters + 2, leafMax);
        CK(cudaMemcpyAsync(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        nItems = hostCounters[2];
        std::swap(cur, nxt);
        if (depth > 200) { snprintf(err, 256, "BuildWideBVH: runaway depth"); cleanup(); return false; }
    }
    out->nNodes = hostCounters[0];
    out->depth = depth;
    float hcb[6];
    // scene bounds = root binary node bounds
    float4 rmn, rmx;
    int rootId = 0;
    CK(cudaMemcpy(&rmn, nmin + rootId, sizeof(float4), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&rmx, nmax + rootId, sizeof(float4), cudaMemcpyDeviceToHost));
    (void)hcb;
    out->bmin[0] = rmn.x; out->bmin[1] = rmn.y; out->bmin[2] = rmn.z;
    out->bmax[0] = rmx.x; out->bmax[1] = rmx.y; out->bmax[2] = rmx.z;
    if (hostCounters[1] != n) { snprintf(err, 256, "BuildWideBVH: emitted %d of %d triangles", hostCounters[1], n); cleanup(); return false; }
    if (depth >= BPT_MAX_BVH_DEPTH) { snprintf(err, 256, "BuildWideBVH: depth %d exceeds traversal stack", depth); cleanup(); return false; }
    cleanup();
    CK(cudaGetLastError());
    return true;
}

void FreeWideBVH(WideBVH *b) {
    if (b->nodes) cudaFreeAsync(b->nodes, nullptr);
    if (b->tris) cudaFreeAsync(b->tris, nullptr);
    b->nodes = nullptr; b->tris = nullptr;
}

}  // namespace bpt
