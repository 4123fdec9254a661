// wavefront.cuh -- device-side state of the wavefront path tracer: SOA path state, index queues, render context.
//
// B200-first restructuring of the reference's work queues (wavefront/workitems.soa, workqueue.h:41-172):
// the reference copies a 184..452-byte work item from queue to queue at every stage; here the path state
// stays in place in float4-packed SOA arrays indexed by the path slot and the queues
// (RayQueue / EscapedRayQueue / HitAreaLightQueue / per-material MaterialEvalQueue) carry only 4-byte slot
// indices, compacted with warp-aggregated atomics.  Only the ShadowRayQueue carries a payload (48 B).
#pragma once
#include <cuda_runtime.h>
#include "scene.h"
#include "lights.h"
#include "bvh.cuh"
#include "media.h"

namespace bpt {

// queue ids inside RenderCtx::counts
enum { Q_RAY0 = 0, Q_RAY1 = 1, Q_SHADOW = 2, Q_ESCAPED = 3, Q_HITLIGHT = 4, Q_MEDIUM = 5 /* MediumSampleQueue */,
       Q_MSCATTER = 6 /* MediumScatterQueue */, Q_MAT0 = 7 /* + MaterialTag */,
       Q_CURSOR_CLOSEST = Q_MAT0 + MAT_NTAGS, Q_CURSOR_SHADOW, Q_CURSOR_MEDIUM, Q_NCOUNTERS };

struct PathState {
    float4 *rayO;        // o.xyz, (unused)
    float4 *rayD;        // d.xyz, time
    float4 *hit;         // original triangle id (int bits), b0, b1, b2
    int *hitInst;        // instance id of the hit (two-level scenes only)
    float *hitT;         // tHit of the closest hit (scenes with media only; +inf on a miss)
    int *medium;         // Ray::medium of the current segment, -1 = vacuum (scenes with media only)
    float4 *beta, *r_u, *r_l, *L;
    float4 *lambda, *lambdaPdf;
    float4 *ctx0;        // prevIntrCtx: pi.x.low, pi.x.high, pi.y.low, pi.y.high
    float4 *ctx1;        //              pi.z.low, pi.z.high, n.x, n.y
    float4 *ctx2;        //              n.z, ns.x, ns.y, ns.z
    float4 *vs;          // GBufferFilm only: VisibleSurface of the path, 6 float4 per slot in arrays of maxPaths:
                         //   [0] p.xyz | set  [1] n.xyz | u  [2] ns.xyz | v  [3] dpdx  [4] dpdy  [5] albedo
    float4 *rdiff;       // scenes with textures only: the ray's differentials rxOrigin, ryOrigin, rxDirection, ryDirection (4 arrays of
                         //   maxPaths) while misc.w says they are stored (shade.cuh: kPathStoredDiffs)
    uint4 *misc;         // px | py << 16, sampleIndex, sampler dimension, depth | specularBounce << 8 | anyNonSpecular << 9 | diff state
    float2 *aux;         // etaScale, filterWeight
};

struct ShadowItem { float4 o_slot; float4 d; float4 Ld; };   // o.xyz + slot bits ; d.xyz + tMax ; Ld
// Shadow ray of a scene with media or "interface" surfaces (VolPathIntegrator::SampleLd's transmittance loop,
// cpu/integrators.cpp:1331-1389 / wavefront TraceTransmittance): the ray may cross several interface surfaces, so the item
// carries the light-side end point (interval + normal) for SpawnRayTo and the factors of the final estimator separately.
struct ShadowTrItem {
    float4 o_slot;     // o.xyz, slot
    float4 d_med;      // d.xyz, medium id (int bits)
    float4 bf;         // beta * f_hat
    float4 Ls;         // ls->L
    float4 rl, ru;     // r_p * p_l, r_p * scatterPDF
    float4 pl0, pl1;   // pLight: x.low x.high y.low y.high | z.low z.high n.x n.y
    float4 pl2;        // n.z, -, -, -
};

struct RenderCtx {
    PathState ps;
    int *queues[Q_CURSOR_CLOSEST];     // item arrays for Q_RAY0.. (Q_SHADOW unused: shadow items below)
    ShadowItem *shadowItems;
    ShadowTrItem *shadowTr;            // used instead of shadowItems when useTr
    const DeviceMedium *media;
    int nMedia;
    int cameraMedium;
    int useTr;                         // scene has media or interface surfaces: shadow rays run the transmittance loop
    int storeHitT;                     // GBufferFilm + interface surfaces: tHit of every closest hit is kept in ps.hitT
    int *counts;                       // Q_NCOUNTERS ints
    unsigned long long *stats;         // [0] camera rays [1] indirect rays [2] shadow rays [3] nodes [4] tris [5] counted rays
    BVHView bvh;
    SceneView sv;
    DeviceCamera camera;
    DeviceFilm film;
    DeviceFilter filter;
    SamplerParams sampler;
    IntegratorParams integ;
    double4 *filmPixels;               // xres*yres {rgbSum[3], weightSum}
    GBufferPixel *gbuf;                // xres*yres, GBufferFilm only
    int nInfinite;
    int maxPaths;
    // interleaved row bands (multi-GPU tiling inside one render call): local row lr of this part maps to film row
    // bandY0 + ((lr / bandRows) * bandParts + bandPart) * bandRows + lr % bandRows; bandRows == 0: identity
    int bandRows, bandParts, bandPart, bandY0;
};

#if defined(__CUDACC__)
__device__ __forceinline__ int MapRow(const RenderCtx &rc, int lr) {
    if (rc.bandRows == 0) return lr;
    int b = lr / rc.bandRows;
    return rc.bandY0 + (b * rc.bandParts + rc.bandPart) * rc.bandRows + (lr - b * rc.bandRows);
}
__device__ __forceinline__ int WarpAggPush(int *counter) {
    // Warp-aggregated atomic slot allocation.  Lanes are grouped by the counter they push to (__match_any_sync), not
    // just by __activemask(): the compiler may merge the bodies of `if (a) push(&x); else push(&y);` into one converged
    // call, in which case lanes with different counters execute this together (seen as lost queue entries on B200).
    unsigned m = __activemask();
    unsigned peers = __match_any_sync(m, (unsigned long long)counter);
    int lane = threadIdx.x & 31;
    int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    return base + __popc(peers & ((1u << lane) - 1));
}
#endif

}  // namespace bpt
