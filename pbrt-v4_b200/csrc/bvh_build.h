// bvh_build.h -- host entry points of the GPU BVH builder (bvh_build.cu).
#pragma once
#include <cuda_runtime.h>

namespace bpt {

#define BPT_MAX_BVH_DEPTH 40   // == BPT_STACK_SIZE in bvh.cuh

struct WideBVH {
    float4 *nodes;   // 5 float4 per node (device)
    float4 *tris;    // 3 float4 per triangle, leaf order (device); p0.w = original triangle id
    int nNodes, nTris, depth;
    float bmin[3], bmax[3];
};

// trisDev: 3 float4 per triangle in original order.  Returns false and fills err (256 bytes) on failure.
// leafMax: subtrees of at most this many primitives become one leaf child (1..3; 0 = default 3).  The top-level BVH of an
// instanced scene uses 1: entering an instance (ray transform + a second traversal) is far costlier than a triangle test, so
// every instance gets its own child box.
// leafMax <= 0 / treeletRounds < 0: the defaults (BPT_LEAF_MAX / BPT_TREELET_ROUNDS, overridable by the environment variables of the same names)
bool BuildWideBVH(const float4 *trisDev, int n, WideBVH *out, cudaStream_t stream, char *err, int leafMax = 0, int treeletRounds = -1);
void FreeWideBVH(WideBVH *b);

}  // namespace bpt
