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
bool BuildWideBVH(const float4 *trisDev, int n, WideBVH *out, cudaStream_t stream, char *err);
void FreeWideBVH(WideBVH *b);

}  // namespace bpt
