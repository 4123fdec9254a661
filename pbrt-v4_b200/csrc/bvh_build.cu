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
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "bvh_build.h"
#include "radix_sort.cuh"
#include "pmath.h"
#include "bvh_build_core.h"

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

using namespace bvhb;

#ifndef BPT_TREELET_ROUNDS
#define BPT_TREELET_ROUNDS 2   // treelet-restructuring rounds after the radix tree (0 = plain LBVH); BPT_TREELET_ROUNDS env overrides
#endif

__global__ void morton_kernel(const float4 *__restrict__ tris, int n, const float *__restrict__ cb, uint64_t *keys, int *vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 lo, hi;
    TriBounds(tris[3 * (size_t)i], tris[3 * (size_t)i + 1], tris[3 * (size_t)i + 2], &lo, &hi);
    keys[i] = MortonKey(lo, hi, cb);
    vals[i] = i;
}

// ---- binary radix tree.  Node ids: internal 0..n-2, leaf j -> (n-1)+j.
__global__ void hierarchy_kernel(const uint64_t *__restrict__ keys, int n, int *left, int *right, int *parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int lc, rc;
    RadixTreeNode(keys, n, i, &lc, &rc);
    left[i] = lc; right[i] = rc;
    parent[lc] = i; parent[rc] = i;
    if (i == 0) parent[0] = -1;
}

// Bottom-up pass with per-node arrival counters: the second thread to arrive at an internal node owns it (both subtrees are final).
// FIT: computes leaf and internal boxes and leaf counts (first pass).  Every pass labels each internal node with its height
// (leaves = 0), which schedules the treelet restructuring below.
template <bool FIT>
__global__ void bottom_up_kernel(Tree T, const float4 *__restrict__ tris, const int *__restrict__ sortedVals, int *flags, int *height) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = T.n;
    if (j >= n) return;
    int id = (n - 1) + j;
    if (FIT) {
        int tri = sortedVals[j];
        float4 lo, hi;
        TriBounds(tris[3 * (size_t)tri], tris[3 * (size_t)tri + 1], tris[3 * (size_t)tri + 2], &lo, &hi);
        BVHB_ST(T.nmin + id, lo); BVHB_ST(T.nmax + id, hi); BVHB_ST(T.cnt + id, 1);
    }
    if (n == 1) return;
    __threadfence();
    int p = BVHB_LD(T.parent + id);
    while (p >= 0) {
        if (atomicAdd(&flags[p], 1) == 0) return;   // first arrival: the sibling's thread finishes this node
        __threadfence();
        int l = BVHB_LD(T.left + p), r = BVHB_LD(T.right + p);
        if (FIT) {
            BVHB_ST(T.nmin + p, Min4(BVHB_LD(T.nmin + l), BVHB_LD(T.nmin + r)));
            BVHB_ST(T.nmax + p, Max4(BVHB_LD(T.nmax + l), BVHB_LD(T.nmax + r)));
            BVHB_ST(T.cnt + p, BVHB_LD(T.cnt + l) + BVHB_LD(T.cnt + r));
        }
        int hl = l >= n - 1 ? 0 : BVHB_LD(height + l), hr = r >= n - 1 ? 0 : BVHB_LD(height + r);
        BVHB_ST(height + p, 1 + max(hl, hr));
        __threadfence();
        p = BVHB_LD(T.parent + p);
    }
}

// Treelet restructuring, level-synchronous: nodes of equal height root disjoint subtrees, and restructuring the treelet of a node
// only rewires nodes strictly inside its subtree, so all treelet roots of one height are processed together (one thread per
// treelet, every lane busy), heights ascending (children before parents).  The first version climbed the tree inside the
// bottom-up kernel -- one live lane per warp near the top: 32 ms for 1 M triangles, 290 ms for 10 M (profiles/r02a sweep).
constexpr int kMaxLevels = 128;
constexpr int kTreeletBlock = 64;
// One cooperative launch per restructuring round: histogram of the treelet roots by height, offsets, scatter into per-level lists,
// then the levels in ascending order with a grid-wide barrier between them.  No host round trip and no launch per level: two
// rounds used to be ~200 launches + 2 host syncs, which is what a 1 k-triangle object definition or a 20 k-instance top level paid
// (20 ms per re-commit of the 20 k-instance scene, profiles/r02g).  The DP tables of OptimizeTreelet live in shared memory.
__global__ void __launch_bounds__(kTreeletBlock) treelet_round_kernel(Tree T, const int *height, int gamma,
                                                                     int *levels /* count | offset | cursor, kMaxLevels each */, int *list) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_copt[(1 << kTreeletLeaves) * kTreeletBlock];
    __shared__ unsigned char s_popt[(1 << kTreeletLeaves) * kTreeletBlock];
    __shared__ int s_h[kMaxLevels], s_base[kMaxLevels];
    const int nInternal = T.n - 1;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nThreads = gridDim.x * blockDim.x;
    int *levelCount = levels, *levelOffset = levels + kMaxLevels, *levelCursor = levels + 2 * kMaxLevels;
    // (levels[] is zeroed by the host with a memset before the launch)
    // ---- histogram
    for (int k = threadIdx.x; k < kMaxLevels; k += blockDim.x) s_h[k] = 0;
    __syncthreads();
    for (int i = tid; i < nInternal; i += nThreads)
        if (__ldcg(T.cnt + i) >= gamma) atomicAdd(&s_h[min(__ldcg(height + i), kMaxLevels - 1)], 1);
    __syncthreads();
    for (int k = threadIdx.x; k < kMaxLevels; k += blockDim.x) if (s_h[k]) atomicAdd(&levelCount[k], s_h[k]);
    grid.sync();
    // ---- offsets (every block computes them for itself: 128 values)
    if (threadIdx.x == 0) { int acc = 0; for (int k = 0; k < kMaxLevels; ++k) { s_base[k] = acc; acc += __ldcg(levelCount + k); } }
    __syncthreads();
    if (blockIdx.x == 0) for (int k = threadIdx.x; k < kMaxLevels; k += blockDim.x) levelOffset[k] = s_base[k];
    // ---- scatter
    for (int i = tid; i < nInternal; i += nThreads)
        if (__ldcg(T.cnt + i) >= gamma) {
            int lvl = min(__ldcg(height + i), kMaxLevels - 1);
            int pos = atomicAdd(&levelCursor[lvl], 1);
            list[s_base[lvl] + pos] = i;
        }
    grid.sync();
    // ---- levels, ascending
    float *copt = s_copt + threadIdx.x;
    unsigned char *popt = s_popt + threadIdx.x;
    for (int lvl = 0; lvl < kMaxLevels; ++lvl) {
        const int n = __ldcg(levelCount + lvl);
        if (n == 0) continue;                       // (uniform over the grid)
        const int off = s_base[lvl];
        if (n >= nThreads / 32) {
            for (int i = tid; i < n; i += nThreads) OptimizeTreelet<kTreeletBlock>(T, __ldcg(list + off + i), copt, popt);
        } else {
            // few treelets at this height (the upper levels: ~40 heights with a handful of nodes each): one WARP per treelet, the
            // subsets of each size spread over the lanes -- the sequential DP of one thread (~60 us) was the critical path of the
            // round (40 levels x 60 us of a 3 ms round for 1 M triangles)
            const int lane = threadIdx.x & 31, warp = tid >> 5, nWarps = nThreads >> 5;
            float *wc = s_copt + (threadIdx.x >> 5) * (1 << kTreeletLeaves);
            unsigned char *wp = s_popt + (threadIdx.x >> 5) * (1 << kTreeletLeaves);
            for (int i = warp; i < n; i += nWarps) {
                const int root = __ldcg(list + off + i);
                Treelet t;
                FormTreelet(T, root, t);          // (all lanes: same addresses, one transaction each)
                const int nl = t.nl;
                if (nl >= 4) {
                    const unsigned full = (1u << nl) - 1u;
                    if (lane < nl) wc[1u << lane] = 0.f;
                    __syncwarp();
                    for (int sz = 2; sz <= nl; ++sz) {
                        for (unsigned sset = lane + 1; sset <= full; sset += 32)
                            if (__popc(sset) == sz) TreeletSolveSubset<1>(t, sset, wc, wp);
                        __syncwarp();
                    }
                    if (lane == 0 && wc[full] < t.curCost * 0.9999f) RewireTreelet<1>(T, root, t, wp);
                }
                __syncwarp();
            }
        }
        grid.sync();
    }
}

// ---- collapse (bvh_build_core.h CollapseOne), one thread per wide node of the current level
__global__ void collapse_kernel(CollapseCtx C, const CollapseItem *__restrict__ items, int nItems) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nItems) return;
    CollapseOne(C, items[idx]);
}

template <typename T>
static bool dalloc(T **p, size_t n, char *err, cudaStream_t st = nullptr) {
    // stream-ordered allocation: the device pool (release threshold raised in bpt_create) recycles the builder's scratch and
    // output buffers across re-commits instead of paying cudaMalloc / cudaFree (device-wide syncs) every time
    cudaError_t e = cudaMallocAsync((void **)p, std::max<size_t>(n, 1) * sizeof(T), st);
    if (e != cudaSuccess) { snprintf(err, 256, "cudaMallocAsync(%zu): %s", n * sizeof(T), cudaGetErrorString(e)); return false; }
    return true;
}

bool BuildWideBVH(const float4 *trisDev, int n, WideBVH *out, cudaStream_t stream, char *err, int leafMax, int treeletRounds) {
    if (leafMax <= 0 || leafMax > 3) {
        leafMax = BPT_LEAF_MAX;
        if (const char *lm = std::getenv("BPT_LEAF_MAX")) leafMax = std::max(1, std::min(3, std::atoi(lm)));   // tuning / invariance tests
    }
    out->nodes = nullptr; out->tris = nullptr; out->nNodes = 0; out->nTris = n; out->depth = 0;
    if (n <= 0) { snprintf(err, 256, "BuildWideBVH: empty triangle set"); return false; }
    int rounds = treeletRounds >= 0 ? std::min(8, treeletRounds) : BPT_TREELET_ROUNDS;
    if (const char *tr = std::getenv("BPT_TREELET_ROUNDS")) rounds = std::max(0, std::min(8, std::atoi(tr)));
    const int B = 256;
    float *cb = nullptr;
    uint64_t *keys = nullptr, *keysTmp = nullptr;
    int *vals = nullptr, *valsTmp = nullptr;
    unsigned *hist = nullptr;
    int *left = nullptr, *right = nullptr, *parent = nullptr, *cnt = nullptr, *flags = nullptr, *height = nullptr, *levels = nullptr, *levelList = nullptr;
    float4 *nmin = nullptr, *nmax = nullptr;
    CollapseItem *q0 = nullptr, *q1 = nullptr;
    int *counters = nullptr;   // [0]=nodeCounter [1]=triCounter [2]=nextCount
    int nTiles = (n + RS_TILE - 1) / RS_TILE;
    bool ok = dalloc(&cb, 6, err, stream) && dalloc(&keys, n, err, stream) && dalloc(&keysTmp, n, err, stream) && dalloc(&vals, n, err, stream) &&
              dalloc(&valsTmp, n, err, stream) && dalloc(&hist, 256 * (size_t)nTiles, err, stream) && dalloc(&left, n, err, stream) &&
              dalloc(&right, n, err, stream) && dalloc(&parent, 2 * (size_t)n, err, stream) && dalloc(&cnt, 2 * (size_t)n, err, stream) &&
              dalloc(&flags, n, err, stream) && dalloc(&height, n, err, stream) && dalloc(&levels, 3 * kMaxLevels, err, stream) && dalloc(&levelList, n, err, stream) && dalloc(&nmin, 2 * (size_t)n, err, stream) && dalloc(&nmax, 2 * (size_t)n, err, stream) &&
              dalloc(&q0, n, err, stream) && dalloc(&q1, n, err, stream) && dalloc(&counters, 4, err, stream) &&
              dalloc(&out->nodes, 5 * (size_t)n, err, stream) && dalloc(&out->tris, 3 * (size_t)n, err, stream);
    auto cleanup = [&]() {
        void *scratch[] = {cb, keys, keysTmp, vals, valsTmp, hist, left, right, parent, cnt, flags, height, levels, levelList, nmin, nmax, q0, q1, counters};
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
    Tree T{n, left, right, parent, cnt, nmin, nmax};
    if (n > 1) hierarchy_kernel<<<(n - 1 + B - 1) / B, B, 0, stream>>>(skeys, n, left, right, parent);
    else CK(cudaMemsetAsync(parent, 0xff, sizeof(int), stream));
    // fit pass (boxes, counts, heights), then the restructuring rounds with a doubled minimum subtree size each
    int gamma = kTreeletLeaves;
    CK(cudaMemsetAsync(flags, 0, sizeof(int) * n, stream));
    bottom_up_kernel<true><<<(n + 127) / 128, 128, 0, stream>>>(T, trisDev, svals, flags, height);
    if (rounds > 0 && n > kTreeletLeaves) {
        int dev = 0, sms = 0, perSM = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, treelet_round_kernel, kTreeletBlock, 0));
        // cooperative grid: all blocks resident; small trees get a small grid (the barriers are cheaper)
        int blocks = std::max(1, std::min(sms * std::max(perSM, 1), (n / 4 + kTreeletBlock - 1) / kTreeletBlock));
        for (int r = 0; r < rounds; ++r, gamma *= 2) {
            if (r > 0) {   // heights of the restructured tree
                CK(cudaMemsetAsync(flags, 0, sizeof(int) * n, stream));
                bottom_up_kernel<false><<<(n + 127) / 128, 128, 0, stream>>>(T, trisDev, svals, flags, height);
            }
            CK(cudaMemsetAsync(levels, 0, sizeof(int) * 3 * kMaxLevels, stream));
            int gammaArg = gamma;
            void *args[] = {(void *)&T, (void *)&height, (void *)&gammaArg, (void *)&levels, (void *)&levelList};
            CK(cudaLaunchCooperativeKernel((const void *)treelet_round_kernel, dim3(blocks), dim3(kTreeletBlock), args, 0, stream));
        }
    }
    // collapse, level by level
    int hostCounters[4] = {1, 0, 0, 0};   // root wide node is node 0
    CK(cudaMemcpyAsync(counters, hostCounters, sizeof(hostCounters), cudaMemcpyHostToDevice, stream));
    CollapseItem root{0 /* internal node 0, or leaf id (n-1)+0 == 0 when n == 1 */, 0};
    CK(cudaMemcpyAsync(q0, &root, sizeof(root), cudaMemcpyHostToDevice, stream));
    int nItems = 1, depth = 0;
    CollapseItem *cur = q0, *nxt = q1;
    while (nItems > 0) {
        ++depth;
        CK(cudaMemsetAsync(counters + 2, 0, sizeof(int), stream));
        CollapseCtx C{T, svals, trisDev, out->nodes, out->tris, counters, counters + 1, nxt, counters + 2, leafMax};
        collapse_kernel<<<(nItems + 63) / 64, 64, 0, stream>>>(C, cur, nItems);
        CK(cudaMemcpyAsync(hostCounters, counters, sizeof(hostCounters), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        nItems = hostCounters[2];
        std::swap(cur, nxt);
        if (depth > 200) { snprintf(err, 256, "BuildWideBVH: runaway depth"); cleanup(); return false; }
    }
    out->nNodes = hostCounters[0];
    out->depth = depth;
    // scene bounds = root binary node bounds
    float4 rmn, rmx;
    CK(cudaMemcpyAsync(&rmn, nmin, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&rmx, nmax, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
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
