// wavefront.cu -- the wavefront path-tracing kernels (sm_100a).
//
// Kernel                      reference counterpart (what it must reproduce, CPU semantics)
//   k_generate_camera_rays    RayIntegrator::EvaluatePixelSample cpu/integrators.cpp:227-289 (+ wavefront/camera.cpp:31-80)
//   k_trace_closest           BVHAggregate::Intersect cpu/aggregates.cpp:529-579 + EnqueueWorkAfterIntersection/Miss
//                             wavefront/intersect.h:16-156   (persistent threads, dynamic ray fetch, wide BVH)
//   k_handle_escaped          VolPathIntegrator::Li miss branch cpu/integrators.cpp:1081-1098 (wavefront/integrator.cpp:495-537)
//   k_handle_emissive         Li emission branch cpu/integrators.cpp:1100-1112 (wavefront/integrator.cpp:539-573)
//   k_shade<Material>         Li body cpu/integrators.cpp:1115-1268 + SampleLd :1273-1390 (wavefront/surfscatter.cpp:57-328)
//   k_trace_shadow            SampleLd shadow loop cpu/integrators.cpp:1331-1382 (wavefront/intersect.h:31-46)
//   k_update_film             EvaluatePixelSample tail + RGBFilm::AddSample film.h:239-255 (wavefront/film.cpp:13-39)
// Sampler dimensions are consumed sequentially exactly as VolPathIntegrator::Li does: the path
// carries its own dimension counter.
// This file is compiled twice: as it is (BPT_WITH_SHAPES = 0: triangle-only scenes, the benchmark configurations) and through
// wavefront_shapes.cu (BPT_WITH_SHAPES = 1: the same kernels with the quadric leaf test and the quadric interactions / area lights
// compiled in, every name inside namespace bpt::tu_shapes).  The engine picks the set per scene, so the quadric code costs scenes
// without quadrics nothing (with it in, the closest-hit kernel gained local-memory traffic around the outlined call and the diffuse
// shading kernel went from 96 to 109 registers).
#ifndef BPT_WITH_SHAPES
#define BPT_WITH_SHAPES 0
#endif
#define BPT_SOBOL_DEVICE_TABLE
#include "wavefront.cuh"
#include "radix_sort.cuh"
#include "bxdfs.h"
#include "materials.cuh"
#include "kernels.h"
#include "camera.h"
#include "shade.cuh"

namespace bpt {
#if BPT_WITH_SHAPES
namespace tu_shapes {
#endif
constexpr bool kShapes = BPT_WITH_SHAPES != 0;

uint32_t h_sobol32[2 * 52];

void UploadSobolTable(const uint32_t *t) {
    for (int i = 0; i < 104; ++i) h_sobol32[i] = t[i];
    cudaMemcpyToSymbol(c_sobol32, t, sizeof(uint32_t) * 104);
    static uint32_t lut[2 * 4 * 256];
    for (int d = 0; d < 2; ++d)
        for (int b = 0; b < 4; ++b)
            for (int x = 0; x < 256; ++x) {
                uint32_t v = 0;
                for (int i = 0; i < 8; ++i) if (x >> i & 1) v ^= t[d * 52 + b * 8 + i];
                lut[(d * 4 + b) * 256 + x] = v;
            }
    cudaMemcpyToSymbol(g_sobolByteLUT, lut, sizeof(lut));
}

// Per-render sampler constants of the outlined ZSobol function (sampling.h ZSobolOutlined): stream-ordered, so engines that share a
// device may use different samplers
void UploadSamplerConstants(const SamplerParams &sp, cudaStream_t st) {
    static thread_local uint64_t hashes[kDimHashTable];
    for (int d = 0; d < kDimHashTable; ++d) hashes[d] = Hash_i32x2(d, sp.seed);
    cudaMemcpyToSymbolAsync(c_sampler, &sp, sizeof(SamplerParams), 0, cudaMemcpyHostToDevice, st);
    cudaMemcpyToSymbolAsync(c_dimHash, hashes, sizeof(hashes), 0, cudaMemcpyHostToDevice, st);
}

__global__ void k_reset_counts(int *counts, int first, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) counts[first + i] = 0;
}
__global__ void k_reset_depth(int *counts, int nextRayQ) {
    int i = threadIdx.x;
    if (i == nextRayQ || i == Q_SHADOW || i == Q_ESCAPED || i == Q_HITLIGHT || i == Q_MEDIUM || i == Q_MSCATTER || (i >= Q_MAT0 && i < Q_NCOUNTERS)) counts[i] = 0;
}

// ---------------------------------------------------------------- ZSobol per-pixel digit prefix (see SamplerParams::prefix)
__global__ void k_zsobol_prefix(SamplerParams sp, uint32_t *table) {
    const size_t nPix = (size_t)sp.prefixW * sp.prefixH;
    const size_t total = nPix * sp.prefixDims;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int dim = (int)(i / nPix);
        int pix = (int)(i - (size_t)dim * nPix);
        SamplerParams plain = sp;
        plain.prefix = nullptr;
        Sampler s;
        s.StartPixelSample(&plain, pix % sp.prefixW, pix / sp.prefixW, 0, dim);
        bool pow2Samples = sp.log2SamplesPerPixel & 1;
        table[i] = (uint32_t)s.FinishSampleIndex(0, sp.nBase4Digits - 1, sp.nLowDigits, pow2Samples, false);
    }
}
void LaunchZSobolPrefix(const SamplerParams &sp, uint32_t *table, int sms, cudaStream_t st) {
    k_zsobol_prefix<<<sms * 8, 256, 0, st>>>(sp, table);
}

// ---------------------------------------------------------------- camera rays
__global__ void __launch_bounds__(256) k_generate_camera_rays(RenderCtx rc, int x0, int x1, int y0, int y1, int firstSample, int nSamples) {
    int nPix = (x1 - x0) * (y1 - y0);
    int total = nPix * nSamples;
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < total; slot += gridDim.x * blockDim.x) {
        int s = slot / nPix, pix = slot - s * nPix;
        int py = MapRow(rc, y0 + pix / (x1 - x0)), px = x0 + pix % (x1 - x0);
        int sampleIndex = firstSample + s;
        Sampler sampler;
        CamSample cs = GetCameraSample(rc.sampler, rc.filter, sampler, px, py, sampleIndex);
        Wavelengths wl = Wavelengths::SampleVisible(cs.lu);
        Vec2 pFilm = cs.pFilm, pLens = cs.pLens;
        float fweight = cs.fweight, time = cs.time;
        // PerspectiveCamera::GenerateRayDifferential (cameras.cpp:429-447), main ray only: the statements of GenerateCameraRay()
        // (camera.h), kept inline here -- calling the helper reorders this kernel's instructions, and a changed binary needs a GPU run
        Pt3 pCamera = XformPoint(rc.camera.cameraFromRaster, Pt3(pFilm.x, pFilm.y, 0));
        Vec3 dir = Normalize(Vec3(pCamera.x, pCamera.y, pCamera.z));
        Pt3 o(0, 0, 0);
        if (rc.camera.lensRadius > 0) {
            Vec2 dsk = SampleUniformDiskConcentric(pLens);
            Vec2 pl(rc.camera.lensRadius * dsk.x, rc.camera.lensRadius * dsk.y);
            float ft = rc.camera.focalDistance / dir.z;
            Pt3 pFocus = o + dir * ft;
            o = Pt3(pl.x, pl.y, 0);
            dir = Normalize(pFocus - o);
        }
        XformRay(rc.camera.renderFromCamera, o, dir);
        if (rc.camera.type != 0) GenerateCameraRay(rc.camera, pFilm, pLens, &o, &dir);   // orthographic: the helper (uniform branch, rare)
        const PathState &ps = rc.ps;
        ps.rayO[slot] = make_float4(o.x, o.y, o.z, 0.f);
        ps.rayD[slot] = make_float4(dir.x, dir.y, dir.z, time);
        ps.beta[slot] = make_float4(1, 1, 1, 1);
        ps.r_u[slot] = make_float4(1, 1, 1, 1);
        ps.r_l[slot] = make_float4(1, 1, 1, 1);
        ps.L[slot] = make_float4(0, 0, 0, 0);
        ps.lambda[slot] = make_float4(wl.lambda[0], wl.lambda[1], wl.lambda[2], wl.lambda[3]);
        ps.lambdaPdf[slot] = make_float4(wl.pdf[0], wl.pdf[1], wl.pdf[2], wl.pdf[3]);
        ps.misc[slot] = make_uint4((unsigned)px | ((unsigned)py << 16), (unsigned)sampleIndex, (unsigned)sampler.dimension, 0u);
        ps.aux[slot] = make_float2(1.f, fweight);
        if (rc.nMedia > 0) ps.medium[slot] = rc.cameraMedium;   // CameraBase::medium (cameras.h)
        if (rc.film.isGBuffer) {
            ps.vs[slot] = make_float4(0, 0, 0, 0);   // VisibleSurface::set = false
            if (rc.storeHitT) { ps.vs[3 * (size_t)rc.maxPaths + slot] = make_float4(0, 0, 0, 0); ps.vs[4 * (size_t)rc.maxPaths + slot] = make_float4(0, 0, 0, 0); }
        }
        rc.queues[Q_RAY0][slot] = slot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        rc.counts[Q_RAY0] = total;
        atomicAdd(&rc.stats[0], (unsigned long long)total);
    }
}

// ---------------------------------------------------------------- closest hit (persistent threads + dynamic fetch)
#ifndef BPT_TRACE_CTAS
#define BPT_TRACE_CTAS 8      // resident CTAs per SM the persistent traversal kernels are sized (and register-limited) for
#endif
// INST selects the two-level traversal at compile time: one traversal body (and one triangle ring in shared memory) per kernel
template <bool COUNT, bool INST>
__global__ void __launch_bounds__(128, BPT_TRACE_CTAS) k_trace_closest(RenderCtx rc, int curQ) {
    const int nRays = rc.counts[curQ];
    const int *queue = rc.queues[curQ];
    unsigned locNodes = 0, locTris = 0, locRays = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&rc.stats[1], (unsigned long long)nRays);
    auto load = [&](int i, Pt3 *o, Vec3 *d, float *tMax) {
        int slot = queue[i];
        float4 o4 = rc.ps.rayO[slot], d4 = rc.ps.rayD[slot];
        *o = Pt3(o4.x, o4.y, o4.z); *d = Vec3(d4.x, d4.y, d4.z); *tMax = BPT_INFINITY;
    };
    auto commit = [&](int i, const HitRec &rec) {
        int slot = queue[i];
        if (rc.nMedia > 0 && rc.ps.medium[slot] >= 0) {
            // ray.medium set: the segment goes through SampleMediumInteraction first (EnqueueWorkAfterMiss/AfterIntersection,
            // wavefront/intersect.h:16-29,48-70); the medium kernel forwards unscattered paths to the surface queues
            int orig = -1;
            if (rec.tri >= 0) orig = __float_as_int(__ldg(rc.bvh.tris + (size_t)rec.tri * 3).w);
            rc.ps.hit[slot] = make_float4(__int_as_float(orig), rec.b0, rec.b1, rec.b2);
            rc.ps.hitT[slot] = rec.tri >= 0 ? rec.t : BPT_INFINITY;
            int p = WarpAggPush(&rc.counts[Q_MEDIUM]);
            rc.queues[Q_MEDIUM][p] = slot;
            return;
        }
        if (rec.tri < 0) {
            if (rc.nInfinite > 0) { int p = WarpAggPush(&rc.counts[Q_ESCAPED]); rc.queues[Q_ESCAPED][p] = slot; }
        } else {
            const float4 *tp = rc.bvh.tris + (size_t)rec.tri * 3;
            int orig = __float_as_int(__ldg(tp).w);
            unsigned cls = __float_as_uint(__ldg(tp + 1).w);
            rc.ps.hit[slot] = make_float4(__int_as_float(orig), rec.b0, rec.b1, rec.b2);
            if (rc.storeHitT) rc.ps.hitT[slot] = rec.t;
            if (cls & 0x100u) { int p = WarpAggPush(&rc.counts[Q_HITLIGHT]); rc.queues[Q_HITLIGHT][p] = slot; }
            unsigned mt = cls & 0xffu;
            int p = WarpAggPush(&rc.counts[Q_MAT0 + mt]);
            rc.queues[Q_MAT0 + mt][p] = slot;
        }
    };
    auto commitInst = [&](int i, const HitRec &rec) {
        int slot = queue[i];
        if (rec.tri < 0) {
            if (rc.nInfinite > 0) { int p = WarpAggPush(&rc.counts[Q_ESCAPED]); rc.queues[Q_ESCAPED][p] = slot; }
        } else {
            const DeviceInstance &I = rc.bvh.instances[rec.inst];
            const float4 *tp = I.tris + (size_t)rec.tri * 3;
            int orig = __float_as_int(__ldg(tp).w) + I.triBase;
            unsigned cls = __float_as_uint(__ldg(tp + 1).w);
            rc.ps.hit[slot] = make_float4(__int_as_float(orig), rec.b0, rec.b1, rec.b2);
            rc.ps.hitInst[slot] = rec.inst;
            if (cls & 0x100u) { int p = WarpAggPush(&rc.counts[Q_HITLIGHT]); rc.queues[Q_HITLIGHT][p] = slot; }
            unsigned mt = cls & 0xffu;
            int p = WarpAggPush(&rc.counts[Q_MAT0 + mt]);
            rc.queues[Q_MAT0 + mt][p] = slot;
        }
    };
    if (INST)
        TraversePersistentInst<false, COUNT, kShapes>(rc.bvh, rc.bvh.instances, nRays, &rc.counts[Q_CURSOR_CLOSEST], load, commitInst, &locNodes, &locTris, &locRays);
    else
        TraversePersistent<false, COUNT, kShapes>(rc.bvh, nRays, &rc.counts[Q_CURSOR_CLOSEST], load, commit, &locNodes, &locTris, &locRays);
    if (COUNT) {
        atomicAdd(&rc.stats[3], (unsigned long long)locNodes);
        atomicAdd(&rc.stats[4], (unsigned long long)locTris);
        atomicAdd(&rc.stats[5], (unsigned long long)locRays);
    }
}

// ---------------------------------------------------------------- shadow rays
template <bool COUNT, bool INST>
__global__ void __launch_bounds__(128, BPT_TRACE_CTAS) k_trace_shadow(RenderCtx rc) {
    const int nRays = rc.counts[Q_SHADOW];
    unsigned locNodes = 0, locTris = 0, locRays = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&rc.stats[2], (unsigned long long)nRays);
    auto load = [&](int i, Pt3 *o, Vec3 *d, float *tMax) {
        float4 a = rc.shadowItems[i].o_slot, b = rc.shadowItems[i].d;
        *o = Pt3(a.x, a.y, a.z); *d = Vec3(b.x, b.y, b.z); *tMax = b.w;
    };
    auto commit = [&](int i, const HitRec &rec) {
        if (rec.tri < 0) {
            int slot = __float_as_int(rc.shadowItems[i].o_slot.w);
            float4 Ld = rc.shadowItems[i].Ld;
            float4 L = rc.ps.L[slot];
            L.x += Ld.x; L.y += Ld.y; L.z += Ld.z; L.w += Ld.w;
            rc.ps.L[slot] = L;
        }
    };
    if (INST)
        TraversePersistentInst<true, COUNT, kShapes>(rc.bvh, rc.bvh.instances, nRays, &rc.counts[Q_CURSOR_SHADOW], load, commit, &locNodes, &locTris, &locRays);
    else
        TraversePersistent<true, COUNT, kShapes>(rc.bvh, nRays, &rc.counts[Q_CURSOR_SHADOW], load, commit, &locNodes, &locTris, &locRays);
    if (COUNT) {
        atomicAdd(&rc.stats[6], (unsigned long long)locNodes);
        atomicAdd(&rc.stats[7], (unsigned long long)locTris);
        atomicAdd(&rc.stats[8], (unsigned long long)locRays);
    }
}

// ---------------------------------------------------------------- emission at hit vertices
__global__ void __launch_bounds__(128) k_handle_emissive(RenderCtx rc) {
    const int n = rc.counts[Q_HITLIGHT];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int slot = rc.queues[Q_HITLIGHT][i];
        const PathState &ps = rc.ps;
        float4 h = ps.hit[slot];
        int tri = __float_as_int(h.x);
        TriData td = LoadTriangle(rc.sv, tri);
        SurfaceIntr isect = HitInteraction<kShapes>(rc.sv, td, h.y, h.z, h.w);
        const DeviceMesh &mesh = rc.sv.meshes[td.mesh];
        int lightIndex = mesh.areaLightBase + td.triIndex;
        const DeviceLight &light = rc.sv.lights[lightIndex];
        float4 d4 = ps.rayD[slot];
        Vec3 rd(d4.x, d4.y, d4.z);
        // (area lights are not allowed on instanced shapes -- scene.cpp -- so emissive hits are always in the identity instance)
        Wavelengths wl = LoadLambda(ps, slot);
        Spec Le = AreaLightL(rc.sv, light, isect.n, -rd, wl);
        if (!Le.IsNonZero()) continue;
        uint4 misc = ps.misc[slot];
        int depth = misc.w & 0xff;
        bool specularBounce = (misc.w >> 8) & 1;
        Spec beta = ld4(&ps.beta[slot]), r_u = ld4(&ps.r_u[slot]);
        Spec L = ld4(&ps.L[slot]);
        if (depth == 0 || specularBounce)
            L += beta * Le / r_u.Average();
        else {
            LightSampleContext prev = LoadCtx(ps, slot);
            float p_l = LightSamplerPMF(rc.sv, prev, lightIndex) * AreaLightPDF_Li<kShapes>(rc.sv, light, prev, rd);
            Spec r_l = ld4(&ps.r_l[slot]);
            r_l *= p_l;
            L += beta * Le / (r_u + r_l).Average();
        }
        st4(&ps.L[slot], L);
    }
}

// ---------------------------------------------------------------- escaped rays (infinite lights)
__global__ void __launch_bounds__(128) k_handle_escaped(RenderCtx rc) {
    const int n = rc.counts[Q_ESCAPED];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int slot = rc.queues[Q_ESCAPED][i];
        const PathState &ps = rc.ps;
        float4 d4 = ps.rayD[slot];
        Vec3 rd(d4.x, d4.y, d4.z);
        Wavelengths wl = LoadLambda(ps, slot);
        uint4 misc = ps.misc[slot];
        int depth = misc.w & 0xff;
        bool specularBounce = (misc.w >> 8) & 1;
        Spec beta = ld4(&ps.beta[slot]), r_u = ld4(&ps.r_u[slot]), r_l = ld4(&ps.r_l[slot]);
        Spec L = ld4(&ps.L[slot]);
        LightSampleContext prev = LoadCtx(ps, slot);
        for (int k = 0; k < rc.sv.ls.nInfinite; ++k) {
            int lightIndex = rc.sv.ls.infiniteLights[k];
            const DeviceLight &light = rc.sv.lights[lightIndex];
            Spec Le = InfiniteLightLe(rc.sv, light, rd, wl);
            if (!Le.IsNonZero()) continue;
            if (depth == 0 || specularBounce)
                L += beta * Le / r_u.Average();
            else {
                float p_l = LightSamplerPMF(rc.sv, prev, lightIndex) * InfiniteLightPDF_Li(rc.sv, light, prev, rd);
                r_l *= p_l;
                L += beta * Le / (r_u + r_l).Average();
            }
        }
        st4(&ps.L[slot], L);
    }
}

// ---------------------------------------------------------------- participating media
// VolPathIntegrator::Li, "Sample the participating medium" (cpu/integrators.cpp:970-1078) == wavefront SampleMediumInteraction
// (wavefront/media.cpp:22-257) with VolPath's sampler consumption (3 x Get1D per segment).  Persistent and warp-synchronous:
// every lane runs one TMajTracker (one majorant-segment fetch or one tentative collision per iteration) and lanes whose
// segment ends -- absorbed, scattered, or through -- are refilled from the queue, because the number of tentative collisions per
// segment varies by an order of magnitude between neighbouring paths.  Real scattering events go to the MediumScatterQueue.
#ifndef BPT_MEDIUM_CTAS
#define BPT_MEDIUM_CTAS 8     // resident 128-thread CTAs per SM the medium kernels are register-limited for (r01k: -34 % vs 3)
#endif
__global__ void __launch_bounds__(128, BPT_MEDIUM_CTAS) k_sample_medium(RenderCtx rc) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int n = rc.counts[Q_MEDIUM];
    int *cursor = &rc.counts[Q_CURSOR_MEDIUM];
    const PathState &ps = rc.ps;
    const int maxDepth = rc.integ.maxDepth;
    int slot = -1;
    bool exhausted = false;
    TMajTracker trk;
    RNG rng;
    Wavelengths wl;
    Spec beta, r_u, r_l;
    uint4 misc = make_uint4(0, 0, 0, 0);
    int depth = 0, dimension = 0, mediumId = 0;
    while (true) {
        unsigned idle = __ballot_sync(FULL, slot < 0);
        if (idle != 0 && !exhausted && (__popc(idle) >= 8 || idle == FULL)) {
            int nNeed = __popc(idle);
            int base = 0;
            if (lane == 0) base = atomicAdd(cursor, nNeed);
            base = __shfl_sync(FULL, base, 0);
            if (base + nNeed >= n) exhausted = true;
            if (slot < 0) {
                int i = base + __popc(idle & ((1u << lane) - 1));
                if (i < n) {
                    slot = rc.queues[Q_MEDIUM][i];
                    float4 o4 = ps.rayO[slot], d4 = ps.rayD[slot];
                    mediumId = ps.medium[slot];
                    wl = LoadLambda(ps, slot);
                    misc = ps.misc[slot];
                    depth = misc.w & 0xff;
                    beta = ld4(&ps.beta[slot]); r_u = ld4(&ps.r_u[slot]); r_l = ld4(&ps.r_l[slot]);
                    Sampler sampler;
                    sampler.StartPixelSample(&rc.sampler, misc.x & 0xffff, misc.x >> 16, (int)misc.y, (int)misc.z);
                    uint64_t hash0 = Hash_f32(sampler.Get1D());
                    uint64_t hash1 = Hash_f32(sampler.Get1D());
                    rng = RNG(hash0, hash1);
                    float uT = sampler.Get1D();
                    dimension = sampler.dimension;
                    trk.Init(rc.media[mediumId], rc.sv.specPool, Pt3(o4.x, o4.y, o4.z), Vec3(d4.x, d4.y, d4.z), ps.hitT[slot], uT, wl);
                }
            }
        }
        if (!__any_sync(FULL, slot >= 0)) break;
        if (slot < 0) continue;
        Pt3 p;
        int st = trk.Step(rng, &p);
        if (st == TMajTracker::CONTINUE) continue;
        if (st == TMajTracker::EVENT) {
            if (!beta.IsNonZero()) { slot = -1; continue; }   // terminated
            MediumProps mp = MediumSamplePoint(rc.media[mediumId], rc.sv.specPool, p, wl);
            Spec sigma_maj = trk.seg.sigma_maj, T_maj = trk.T_maj;
            // Handle medium emission (cpu/integrators.cpp:989-1000): added before the absorption / scattering / null decision
            if (depth < maxDepth && mp.Le.IsNonZero()) {
                float pdfE = sigma_maj[0] * T_maj[0];
                Spec betap = beta * T_maj / pdfE;
                Spec r_e = r_u * sigma_maj * T_maj / pdfE;
                if (r_e.IsNonZero()) {
                    Spec Lm = ld4(&ps.L[slot]);
                    Lm += betap * mp.sigma_a * mp.Le / r_e.Average();
                    st4(&ps.L[slot], Lm);
                }
            }
            float pAbsorb = mp.sigma_a[0] / sigma_maj[0];
            float pScatter = mp.sigma_s[0] / sigma_maj[0];
            float pNull = stdmax(0.f, 1 - pAbsorb - pScatter);
            float um = rng.UniformFloat();
            int mode = SampleDiscrete3(pAbsorb, pScatter, pNull, um);
            if (mode == 0) { slot = -1; continue; }            // absorbed
            if (mode == 1) {
                if (depth++ < maxDepth) {
                    float pdf = T_maj[0] * mp.sigma_s[0];
                    beta *= T_maj * mp.sigma_s / pdf;
                    r_u *= T_maj * mp.sigma_s / pdf;
                    if (beta.IsNonZero() && r_u.IsNonZero()) {
                        // real scattering: SampleLd + phase sampling run in k_medium_scatter (MediumScatterQueue)
                        st4(&ps.beta[slot], beta); st4(&ps.r_u[slot], r_u);
                        StoreCtx(ps, slot, Pt3i(p), Nrm3(0, 0, 0), Nrm3(0, 0, 0));
                        ps.misc[slot] = make_uint4(misc.x, misc.y, (unsigned)dimension, (unsigned)depth | (misc.w & 0xf00u));
                        int q = WarpAggPush(&rc.counts[Q_MSCATTER]);
                        rc.queues[Q_MSCATTER][q] = slot;
                    }
                }
                slot = -1;
                continue;
            }
            // null scattering
            Spec sigma_n = ClampZero(sigma_maj - mp.sigma_a - mp.sigma_s);
            float pdf = T_maj[0] * sigma_n[0];
            beta *= T_maj * sigma_n / pdf;
            if (pdf == 0) beta = Spec(0.f);
            r_u *= T_maj * sigma_n / pdf;
            r_l *= T_maj * sigma_maj / pdf;
            if (!(beta.IsNonZero() && r_u.IsNonZero())) { slot = -1; continue; }
            trk.AfterEvent();
            continue;
        }
        // DONE: the ray crossed the medium segment without a real collision
        if (beta.IsNonZero() && r_u.IsNonZero()) {
            Spec T_maj = trk.T_maj;
            beta *= T_maj / T_maj[0];
            r_u *= T_maj / T_maj[0];
            r_l *= T_maj / T_maj[0];
            st4(&ps.beta[slot], beta); st4(&ps.r_u[slot], r_u); st4(&ps.r_l[slot], r_l);
            ps.misc[slot] = make_uint4(misc.x, misc.y, (unsigned)dimension, (unsigned)depth | (misc.w & 0xf00u));
            // surviving unscattered ray: hand over to the surface queues (what the closest-hit commit does for vacuum rays)
            int tri = __float_as_int(ps.hit[slot].x);
            if (tri < 0) {
                if (rc.nInfinite > 0) { int q = WarpAggPush(&rc.counts[Q_ESCAPED]); rc.queues[Q_ESCAPED][q] = slot; }
            } else {
                unsigned cls = __float_as_uint(__ldg(rc.sv.triPool + (size_t)tri * 3 + 2).w);
                if (cls & 0x100u) { int q = WarpAggPush(&rc.counts[Q_HITLIGHT]); rc.queues[Q_HITLIGHT][q] = slot; }
                unsigned mt = cls & 0xffu;
                int q = WarpAggPush(&rc.counts[Q_MAT0 + mt]);
                rc.queues[Q_MAT0 + mt][q] = slot;
            }
        }
        slot = -1;
    }
}

// Real scattering events (wavefront SampleMediumScattering, media.cpp:259-352 == cpu/integrators.cpp:1021-1055): direct lighting at
// the MediumInteraction (SampleLd with bsdf == nullptr: 1D + 2D sampler dimensions) and the Henyey-Greenstein direction (2D).
__global__ void __launch_bounds__(128) k_medium_scatter(RenderCtx rc, int nextQ) {
    const int n = rc.counts[Q_MSCATTER];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int slot = rc.queues[Q_MSCATTER][i];
        const PathState &ps = rc.ps;
        float4 d4 = ps.rayD[slot];
        Vec3 wo = -Vec3(d4.x, d4.y, d4.z);
        const int mediumId = ps.medium[slot];
        const DeviceMedium &med = rc.media[mediumId];
        Wavelengths wl = LoadLambda(ps, slot);
        uint4 misc = ps.misc[slot];
        Sampler sampler;
        sampler.StartPixelSample(&rc.sampler, misc.x & 0xffff, misc.x >> 16, (int)misc.y, (int)misc.z);
        Spec beta = ld4(&ps.beta[slot]), r_u = ld4(&ps.r_u[slot]);
        LightSampleContext ctx = LoadCtx(ps, slot);   // LightSampleContext(MediumInteraction): pi = p, n = ns = 0
        Pt3 p = ctx.pi.p();
        float u = sampler.Get1D();
        float lightPMF = 0;
        int lightIndex = LightSamplerSample(rc.sv, ctx, u, &lightPMF);
        Vec2 uLight = sampler.Get2D();
        if (lightIndex >= 0) {
            const DeviceLight &light = rc.sv.lights[lightIndex];
            LightLiSample ls;
            if (light.type == LIGHT_DIFFUSE_AREA) ls = AreaLightSampleLi<kShapes>(rc.sv, light, ctx, uLight, wl);
            else if (IsDeltaLight(light.type)) ls = DeltaLightSampleLi(rc.sv.specPool, light, ctx.p(), wl);
            else ls = InfiniteLightSampleLi(rc.sv, light, ctx, uLight, wl);
            if (ls.valid && ls.L.IsNonZero() && ls.pdf != 0) {
                float p_l = lightPMF * ls.pdf;
                float ph = PhaseHG_p(med.g, wo, ls.wi);
                Spec f_hat(ph);
                // delta lights: the estimate divides by r_l.Average() alone (cpu/integrators.cpp:1385-1388); a zero scattering pdf
                // makes r_u vanish in (r_l + r_u) and gives the same bits
                float scatterPDF = IsDeltaLight(light.type) ? 0.f : ph;
                if (f_hat.IsNonZero()) {
                    Nrm3 nL = light.type == LIGHT_DIFFUSE_AREA ? ls.nLight : Nrm3(0, 0, 0);
                    Pt3 pf = OffsetRayOrigin(ctx.pi, ctx.n, ls.pLight.p() - ctx.pi.p());
                    Pt3 pt = OffsetRayOrigin(ls.pLight, nL, pf - ls.pLight.p());
                    PushShadowTr(rc, slot, pf, pt - pf, mediumId, beta * f_hat, ls.L, r_u * p_l, r_u * scatterPDF, ls.pLight, nL);
                }
            }
        }
        Vec2 u2 = sampler.Get2D();
        float ppdf;
        Vec3 wi = SampleHenyeyGreenstein(wo, med.g, u2, &ppdf);
        if (ppdf == 0) continue;   // terminated
        beta *= ppdf / ppdf;       // ps->p / ps->pdf with p == pdf
        Spec r_l = r_u / ppdf;
        ps.rayO[slot] = make_float4(p.x, p.y, p.z, 0.f);
        ps.rayD[slot] = make_float4(wi.x, wi.y, wi.z, d4.w);
        st4(&ps.beta[slot], beta); st4(&ps.r_l[slot], r_l);
        // specularBounce = false, anyNonSpecularBounces = true
        ps.misc[slot] = make_uint4(misc.x, misc.y, (unsigned)sampler.dimension, (misc.w & 0xcffu) | 0x200u);   // ray.o = p; ray.d = wi (cpu/integrators.cpp:1048-1049): the RayDifferential keeps whatever differentials it had
        int q = WarpAggPush(&rc.counts[nextQ]);
        rc.queues[nextQ][q] = slot;
    }
}

// Shadow rays through media / interface surfaces: the loop of VolPathIntegrator::SampleLd (cpu/integrators.cpp:1331-1389), i.e.
// the wavefront TraceTransmittance (wavefront/intersect.h:164-274): closest hit, stop at the first surface with a material,
// ratio-track the transmittance through the current medium, continue behind interface surfaces.  One thread per shadow ray: a
// persistent lock-step variant like k_sample_medium was measured 2.2x slower here (few collisions per ray, phases misalign).
__global__ void __launch_bounds__(128, BPT_MEDIUM_CTAS) k_trace_shadow_tr(RenderCtx rc) {
    const int n = rc.counts[Q_SHADOW];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&rc.stats[2], (unsigned long long)n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const ShadowTrItem &it = rc.shadowTr[i];
        float4 a = it.o_slot, b = it.d_med;
        int slot = __float_as_int(a.w);
        Pt3 o(a.x, a.y, a.z); Vec3 d(b.x, b.y, b.z);
        int medium = __float_as_int(b.w);
        float4 q0 = it.pl0, q1 = it.pl1, q2 = it.pl2;
        Pt3i pLight; pLight.x = Interval(q0.x, q0.y); pLight.y = Interval(q0.z, q0.w); pLight.z = Interval(q1.x, q1.y);
        Nrm3 nLight(q1.z, q1.w, q2.x);
        Wavelengths wl = LoadLambda(rc.ps, slot);
        Spec T_ray(1.f), r_l(1.f), r_u(1.f);
        RNG rng(Hash_f32x3(o.x, o.y, o.z), Hash_f32x3(d.x, d.y, d.z));
        bool blocked = false;
        while (d != Vec3(0, 0, 0)) {
            HitRec rec;
            unsigned nn = 0, nt = 0;
            TraceRay<false, false, kShapes>(rc.bvh, o, d, 1 - kShadowEpsilon, &rec, &nn, &nt);
            int tri = -1; unsigned cls = 0;
            if (rec.tri >= 0) {
                const float4 *tp = rc.bvh.tris + (size_t)rec.tri * 3;
                tri = __float_as_int(__ldg(tp).w);
                cls = __float_as_uint(__ldg(tp + 1).w);
                if ((cls & 0xffu) != MAT_INTERFACE) { blocked = true; break; }
            }
            if (medium >= 0) {
                float tMax = rec.tri >= 0 ? rec.t : (1 - kShadowEpsilon);
                float u = rng.UniformFloat();
                Spec T_maj = SampleT_maj(rc.media[medium], rc.sv.specPool, o, d, tMax, u, rng, wl,
                    [&](Pt3 p, const MediumProps &mp, Spec sigma_maj, Spec T_maj) -> bool {
                        Spec sigma_n = ClampZero(sigma_maj - mp.sigma_a - mp.sigma_s);
                        float pdf = T_maj[0] * sigma_maj[0];
                        T_ray *= T_maj * sigma_n / pdf;
                        r_l *= T_maj * sigma_maj / pdf;
                        r_u *= T_maj * sigma_n / pdf;
                        Spec Tr = T_ray / (r_l + r_u).Average();
                        if (Tr.MaxComponentValue() < 0.05f) {
                            float q = 0.75f;
                            if (rng.UniformFloat() < q) T_ray = Spec(0.f);
                            else T_ray /= 1 - q;
                        }
                        return T_ray.IsNonZero();
                    });
                T_ray *= T_maj / T_maj[0];
                r_l *= T_maj / T_maj[0];
                r_u *= T_maj / T_maj[0];
            }
            if (!T_ray.IsNonZero()) { blocked = true; break; }
            if (rec.tri < 0) break;
            // lightRay = si->intr.SpawnRayTo(ls->pLight)
            TriData td = LoadTriangle(rc.sv, tri);
            SurfaceIntr isect = HitInteraction<kShapes>(rc.sv, td, rec.b0, rec.b1, rec.b2);
            Pt3 pf = OffsetRayOrigin(isect.pi, isect.n, pLight.p() - isect.pi.p());
            Pt3 pt = OffsetRayOrigin(pLight, nLight, pf - pLight.p());
            o = pf; d = pt - pf;
            medium = SurfaceGetMedium(rc.sv.meshes[td.mesh], d, isect.n, medium);
        }
        if (blocked) continue;
        r_l *= ld4(&it.rl);
        r_u *= ld4(&it.ru);
        Spec Ld = ld4(&it.bf) * T_ray * ld4(&it.Ls) / (r_l + r_u).Average();
        float4 L = rc.ps.L[slot];
        L.x += Ld[0]; L.y += Ld[1]; L.z += Ld[2]; L.w += Ld[3];
        rc.ps.L[slot] = L;
    }
}

// ---------------------------------------------------------------- ray sort (no reference counterpart)
// 8-bit key per queued ray: direction octant (3 bits) and a coarse cell of the origin inside the scene bounds (2 + 1 + 2 bits).
// One stable LSD pass (radix_sort.cuh) over the slot indices groups rays that take the same child order and start in the same
// region, while the pixel order inside a bucket is preserved.
__global__ void k_ray_sort_keys(RenderCtx rc, int q, int n, float3 bmin, float3 binv, unsigned *keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int slot = rc.queues[q][i];
    float4 o = rc.ps.rayO[slot], d = rc.ps.rayD[slot];
    unsigned oct = (d.x < 0 ? 4u : 0u) | (d.y < 0 ? 2u : 0u) | (d.z < 0 ? 1u : 0u);
    int cx = min(3, max(0, (int)((o.x - bmin.x) * binv.x * 4.f)));
    int cy = min(1, max(0, (int)((o.y - bmin.y) * binv.y * 2.f)));
    int cz = min(3, max(0, (int)((o.z - bmin.z) * binv.z * 4.f)));
    keys[i] = (oct << 5) | ((unsigned)cz << 3) | ((unsigned)cy << 2) | (unsigned)cx;
}

// ---------------------------------------------------------------- film
__global__ void __launch_bounds__(256) k_update_film(RenderCtx rc, int x0, int x1, int y0, int y1, int nSamples) {
    int nPix = (x1 - x0) * (y1 - y0);
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < nPix; pix += gridDim.x * blockDim.x) {
        int py = MapRow(rc, y0 + pix / (x1 - x0)), px = x0 + pix % (x1 - x0);
        double4 acc = rc.filmPixels[(size_t)py * rc.film.xres + px];
        for (int s = 0; s < nSamples; ++s) {
            int slot = s * nPix + pix;
            Spec L = ld4(&rc.ps.L[slot]);
            Wavelengths wl = LoadLambda(rc.ps, slot);
            // NaN / infinite radiance scrub (cpu/integrators.cpp:263-273)
            if (L.HasNaNs()) L = Spec(0.f);
            else {
                Spec Ys = DenseSample(rc.sv.cieY, wl);
                float y = SafeDiv(Ys * L, wl.PDF()).Average() / 106.856895f;
                if (IsInf_(y)) L = Spec(0.f);
            }
            // PixelSensor::ToSensorRGB (film.h:95-100)
            Spec Lp = SafeDiv(L, wl.PDF());
            float r = rc.film.imagingRatio * (DenseSample(rc.sv.cieX, wl) * Lp).Average();
            float g = rc.film.imagingRatio * (DenseSample(rc.sv.cieY, wl) * Lp).Average();
            float b = rc.film.imagingRatio * (DenseSample(rc.sv.cieZ, wl) * Lp).Average();
            float m = stdmax(stdmax(r, g), b);
            if (m > rc.film.maxComponentValue) { float sc = rc.film.maxComponentValue / m; r *= sc; g *= sc; b *= sc; }
            float weight = rc.ps.aux[slot].y;
            if (rc.film.isGBuffer) {
                const size_t M = (size_t)rc.maxPaths;
                float4 v0 = rc.ps.vs[slot];
                if (v0.w != 0) {   // visibleSurface && *visibleSurface  (GBufferFilm::AddSample film.cpp:588-641)
                    GBufferPixel &gp = rc.gbuf[(size_t)py * rc.film.xres + px];
                    float4 v1 = rc.ps.vs[M + slot], v2 = rc.ps.vs[2 * M + slot], v3 = rc.ps.vs[3 * M + slot], v4 = rc.ps.vs[4 * M + slot];
                    gp.gBufferWeightSum += (double)weight;
                    const float rgbc[3] = {r, g, b};
                    for (int c = 0; c < 3; ++c) {   // VarianceEstimator<Float>::Add (util/math.h)
                        ++gp.varN[c];
                        float delta = rgbc[c] - gp.varMean[c];
                        gp.varMean[c] += delta / gp.varN[c];
                        float delta2 = rgbc[c] - gp.varMean[c];
                        gp.varS[c] += delta * delta2;
                    }
                    Pt3 P(v0.x, v0.y, v0.z); Nrm3 N(v1.x, v1.y, v1.z), NS(v2.x, v2.y, v2.z); Vec3 DX(v3.x, v3.y, v3.z), DY(v4.x, v4.y, v4.z);
                    Pt3 po; Nrm3 no, nso; float dzx, dzy;
                    if (rc.film.applyInverse) {
                        po = ApplyInversePoint(rc.film.outMInv, P);
                        no = XformNormalByTranspose(rc.film.outM, N); nso = XformNormalByTranspose(rc.film.outM, NS);
                        dzx = XformVector(rc.film.outMInv, DX).z; dzy = XformVector(rc.film.outMInv, DY).z;
                    } else {
                        po = XformPoint(rc.film.outM, P);
                        no = XformNormalByTranspose(rc.film.outMInv, N); nso = XformNormalByTranspose(rc.film.outMInv, NS);
                        dzx = XformVector(rc.film.outM, DX).z; dzy = XformVector(rc.film.outM, DY).z;
                    }
                    gp.pSum[0] += weight * po.x; gp.pSum[1] += weight * po.y; gp.pSum[2] += weight * po.z;
                    gp.nSum[0] += weight * no.x; gp.nSum[1] += weight * no.y; gp.nSum[2] += weight * no.z;
                    gp.nsSum[0] += weight * nso.x; gp.nsSum[1] += weight * nso.y; gp.nsSum[2] += weight * nso.z;
                    gp.dzdxSum += weight * dzx; gp.dzdySum += weight * dzy;
                    gp.uvSum[0] += weight * v1.w; gp.uvSum[1] += weight * v2.w;
                    // albedo * colorSpace->illuminant.Sample(lambda) -> ToRGB(lambda, colorSpace)
                    Spec alb = ld4(&rc.ps.vs[5 * M + slot]) * DenseSample(rc.sv.illumD65, wl);
                    Spec pdfS = wl.PDF();
                    float X = SafeDiv(DenseSample(rc.sv.cieX, wl) * alb, pdfS).Average() / 106.856895f;
                    float Y = SafeDiv(DenseSample(rc.sv.cieY, wl) * alb, pdfS).Average() / 106.856895f;
                    float Z = SafeDiv(DenseSample(rc.sv.cieZ, wl) * alb, pdfS).Average() / 106.856895f;
                    for (int c = 0; c < 3; ++c) {
                        float v = rc.film.rgbFromXYZ[c][0] * X + rc.film.rgbFromXYZ[c][1] * Y + rc.film.rgbFromXYZ[c][2] * Z;
                        gp.rgbAlbedoSum[c] += (double)(weight * v);
                    }
                }
            }
            acc.x += (double)(weight * r);
            acc.y += (double)(weight * g);
            acc.z += (double)(weight * b);
            acc.w += (double)weight;
        }
        rc.filmPixels[(size_t)py * rc.film.xres + px] = acc;
    }
}

// GBufferFilm::GetImage (film.cpp:690-790): 25 channels per pixel in the reference's order
// R G B | Albedo.RGB | P.XYZ | dzdx dzdy | N.XYZ | Ns.XYZ | u v | Variance.RGB | RelativeVariance.RGB
__global__ void k_resolve_gbuffer(const double4 *pix, const GBufferPixel *gb, DeviceFilm film, float *out) {
    int n = film.xres * film.yres;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double4 p = pix[i];
        const GBufferPixel &g = gb[i];
        float rgb[3] = {(float)p.x, (float)p.y, (float)p.z};
        float alb[3] = {(float)g.rgbAlbedoSum[0], (float)g.rgbAlbedoSum[1], (float)g.rgbAlbedoSum[2]};
        float weightSum = (float)p.w, gws = (float)g.gBufferWeightSum;
        float pt[3] = {g.pSum[0], g.pSum[1], g.pSum[2]}, uv[2] = {g.uvSum[0], g.uvSum[1]};
        float dzdx = g.dzdxSum, dzdy = g.dzdySum;
        if (weightSum != 0) { for (int c = 0; c < 3; ++c) { rgb[c] /= weightSum; alb[c] /= weightSum; } }
        if (gws != 0) { for (int c = 0; c < 3; ++c) pt[c] /= gws; uv[0] /= gws; uv[1] /= gws; dzdx /= gws; dzdy /= gws; }
        for (int c = 0; c < 3; ++c) rgb[c] += 0.f;   // splats: none on this path
        float o3[3];
        for (int r = 0; r < 3; ++r) { float v = 0; for (int c = 0; c < 3; ++c) v += film.outRGBFromSensor[r][c] * rgb[c]; o3[r] = v; }
        float *o = out + 25 * (size_t)i;
        o[0] = o3[0]; o[1] = o3[1]; o[2] = o3[2];
        o[3] = alb[0]; o[4] = alb[1]; o[5] = alb[2];
        o[6] = pt[0]; o[7] = pt[1]; o[8] = pt[2];
        o[9] = fabsf(dzdx); o[10] = fabsf(dzdy);
        Nrm3 nn(g.nSum[0], g.nSum[1], g.nSum[2]), ns(g.nsSum[0], g.nsSum[1], g.nsSum[2]);
        nn = LengthSquared(nn) > 0 ? Normalize(nn) : Nrm3(0, 0, 0);
        ns = LengthSquared(ns) > 0 ? Normalize(ns) : Nrm3(0, 0, 0);
        o[11] = nn.x; o[12] = nn.y; o[13] = nn.z; o[14] = ns.x; o[15] = ns.y; o[16] = ns.z;
        o[17] = uv[0]; o[18] = uv[1];
        for (int c = 0; c < 3; ++c) {
            float var = g.varN[c] > 1 ? g.varS[c] / (g.varN[c] - 1) : 0.f;
            o[19 + c] = var;
            o[22 + c] = (g.varN[c] < 1 || g.varMean[c] == 0) ? 0.f : var / g.varMean[c];
        }
    }
}
void LaunchResolveGBuffer(const double4 *pix, const GBufferPixel *gb, const DeviceFilm &film, float *out, int sms, cudaStream_t st) {
    k_resolve_gbuffer<<<sms * 4, 256, 0, st>>>(pix, gb, film, out);
}

__global__ void k_resolve_film(const double4 *pix, DeviceFilm film, float *rgbOut) {
    int n = film.xres * film.yres;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double4 p = pix[i];
        float rgb[3] = {(float)p.x, (float)p.y, (float)p.z};
        float weightSum = (float)p.w;
        if (weightSum != 0) { rgb[0] /= weightSum; rgb[1] /= weightSum; rgb[2] /= weightSum; }
        // rgbSplat is always zero on this path: rgb[c] += splatScale * 0 / filterIntegral
        for (int c = 0; c < 3; ++c) rgb[c] += 0.f;
        float out[3];
        for (int r = 0; r < 3; ++r) {
            float v = 0;
            for (int c = 0; c < 3; ++c) v += film.outRGBFromSensor[r][c] * rgb[c];
            out[r] = v;
        }
        rgbOut[3 * (size_t)i] = out[0]; rgbOut[3 * (size_t)i + 1] = out[1]; rgbOut[3 * (size_t)i + 2] = out[2];
    }
}

// ---------------------------------------------------------------- raw ray entry points
template <bool ANY, bool COUNT>
__global__ void __launch_bounds__(128) k_trace_raw(BVHView bvh, const float4 *rays, int n, float4 *hits, float *tOut, int *cursor,
                                                   unsigned long long *stats) {
    unsigned locNodes = 0, locTris = 0, locRays = 0;
    auto load = [&](int i, Pt3 *o, Vec3 *d, float *tMax) {
        float4 a = rays[2 * (size_t)i], b = rays[2 * (size_t)i + 1];
        *o = Pt3(a.x, a.y, a.z); *d = Vec3(b.x, b.y, b.z); *tMax = a.w;
    };
    auto commit = [&](int i, const HitRec &rec) {
        int orig = -1;
        if (rec.tri >= 0) orig = __float_as_int(__ldg(bvh.tris + (size_t)rec.tri * 3).w);
        hits[i] = make_float4(__int_as_float(orig), rec.b0, rec.b1, rec.b2);
        if (tOut) tOut[i] = rec.tri >= 0 ? rec.t : BPT_INFINITY;
    };
    auto commitInst = [&](int i, const HitRec &rec) {
        int orig = -1;
        if (rec.tri >= 0) orig = __float_as_int(__ldg(bvh.instances[rec.inst].tris + (size_t)rec.tri * 3).w) + bvh.instances[rec.inst].triBase;
        hits[i] = make_float4(__int_as_float(orig), rec.b0, rec.b1, rec.b2);
        if (tOut) tOut[i] = rec.tri >= 0 ? rec.t : BPT_INFINITY;
    };
    if (bvh.instances) TraversePersistentInst<ANY, COUNT, kShapes>(bvh, bvh.instances, n, cursor, load, commitInst, &locNodes, &locTris, &locRays);
    else TraversePersistent<ANY, COUNT, kShapes>(bvh, n, cursor, load, commit, &locNodes, &locTris, &locRays);
    if (COUNT && stats) { atomicAdd(&stats[3], (unsigned long long)locNodes); atomicAdd(&stats[4], (unsigned long long)locTris); }
}

// ---------------------------------------------------------------- probes for unit parity tests
__global__ void k_probe_sampler(SamplerParams sp, const int *pxy, const int *sidx, const int *dim, int n, float *o1, float *o2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sampler s;
    s.StartPixelSample(&sp, pxy[2 * i], pxy[2 * i + 1], sidx[i], dim[i]);
    o1[i] = s.Get1D();
    s.StartPixelSample(&sp, pxy[2 * i], pxy[2 * i + 1], sidx[i], dim[i]);
    Vec2 v = s.Get2D();
    o2[2 * i] = v.x; o2[2 * i + 1] = v.y;
}
__global__ void k_probe_tri(const float *rays, const float *tris, int n, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *r = rays + 7 * (size_t)i, *t = tris + 9 * (size_t)i;
    TriHit th{0, 0, 0, 0};
    bool hit = IntersectTriangle(Pt3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]), r[6], Pt3(t[0], t[1], t[2]), Pt3(t[3], t[4], t[5]),
                                 Pt3(t[6], t[7], t[8]), &th);
    // cross-check the hoisted variant used by traversal
    RayPre rp = MakeRayPre(Pt3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]));
    TriHit th2{0, 0, 0, 0};
    bool degenerate = LengthSquared(Cross(Pt3(t[6], t[7], t[8]) - Pt3(t[0], t[1], t[2]), Pt3(t[3], t[4], t[5]) - Pt3(t[0], t[1], t[2]))) == 0;
    bool hit2 = !degenerate && IntersectTrianglePre(rp, r[6], Pt3(t[0], t[1], t[2]), Pt3(t[3], t[4], t[5]), Pt3(t[6], t[7], t[8]), &th2);
    bool same = hit == hit2 && (!hit || (th.b0 == th2.b0 && th.b1 == th2.b1 && th.b2 == th2.b2 && th.t == th2.t));
    float *o = out + 5 * (size_t)i;
    o[0] = same ? (hit ? 1.f : 0.f) : -1.f; o[1] = th.b0; o[2] = th.b1; o[3] = th.b2; o[4] = th.t;
}

__global__ void k_probe_math(int fn, const float *x, const float *y, int n, float *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = x[i], b = y[i], r = 0;
    switch (fn) {
    case 0: r = sinf(a); break;
    case 1: r = cosf(a); break;
    case 2: r = asinf(a); break;
    case 3: r = acosf(a); break;
    case 4: r = atan2f(a, b); break;
    case 5: r = atanhf(a); break;
    case 6: r = coshf(a); break;
    case 7: r = expf(a); break;
    case 8: r = logf(a); break;
    case 9: r = tanf(a); break;
    case 10: r = (float)sin((double)a); break;
    case 11: r = (float)cos((double)a); break;
    case 12: r = (float)asin((double)a); break;
    case 13: r = (float)acos((double)a); break;
    case 14: r = (float)atan2((double)a, (double)b); break;
    case 15: r = (float)atanh((double)a); break;
    case 16: r = (float)cosh((double)a); break;
    case 17: r = (float)exp((double)a); break;
    case 18: r = (float)log((double)a); break;
    case 19: r = powf(a, b); break;
    case 20: r = (float)pow((double)a, (double)b); break;
    case 21: r = sqrtf(a); break;
    case 22: r = a / b; break;
    case 23: r = (float)tan((double)a); break;
    // glibc's own float algorithms restated in double arithmetic (libm_exact.h; exhaustively equal to the host libm on the CPU)
    case 30: r = bpt_exact_sinf(a); break;
    case 31: r = bpt_exact_cosf(a); break;
    case 32: r = bpt_exact_logf(a); break;
    case 33: r = bpt_exact_asinf(a); break;
    case 34: r = bpt_exact_acosf(a); break;
    case 35: r = bpt_exact_atan2f(a, b); break;
    case 36: r = bpt_exact_atanhf(a); break;
    case 37: r = bpt_exact_coshf(a); break;
    case 38: r = bpt_exact_expf(a); break;
    }
    out[i] = r;
}
void LaunchProbeMath(int fn, const float *x, const float *y, int n, float *out, cudaStream_t st) {
    k_probe_math<<<(n + 127) / 128, 128, 0, st>>>(fn, x, y, n, out);
}

// ================================================================= host-side launchers
static inline int GridFor(int smCount, int perSM) { return smCount * perSM; }

void LaunchSampleMedium(const RenderCtx &rc, int nextQ, int sms, cudaStream_t st) {
    k_sample_medium<<<GridFor(sms, BPT_MEDIUM_CTAS), 128, 0, st>>>(rc);
    k_medium_scatter<<<GridFor(sms, 8), 128, 0, st>>>(rc, nextQ);
}
void LaunchTraceShadowTr(const RenderCtx &rc, int sms, cudaStream_t st) { k_trace_shadow_tr<<<GridFor(sms, 2 * BPT_MEDIUM_CTAS), 128, 0, st>>>(rc); }
// Sorts queue `q` (n entries, n known on the host) by the ray key; returns true if the sorted indices ended up in queueTmp.
bool LaunchRaySort(const RenderCtx &rc, int q, int n, const float bmin[3], const float bmax[3], unsigned *keys, unsigned *keysTmp,
                   int *queueTmp, unsigned *hist, cudaStream_t st) {
    if (n <= 1) return false;
    float3 lo = make_float3(bmin[0], bmin[1], bmin[2]);
    float3 inv = make_float3(1.f / fmaxf(bmax[0] - bmin[0], 1e-20f), 1.f / fmaxf(bmax[1] - bmin[1], 1e-20f), 1.f / fmaxf(bmax[2] - bmin[2], 1e-20f));
    k_ray_sort_keys<<<(n + 255) / 256, 256, 0, st>>>(rc, q, n, lo, inv, keys);
    return RadixSortPairs<unsigned, int>(keys, keysTmp, rc.queues[q], queueTmp, n, 0, 8, hist, st);
}
// Shared-memory carve-out of the persistent traversal kernels: they need BPT_TRACE_CTAS x 9 KB; everything else should be L1
// (node / triangle fetches hit L1 in ~40 % of the sectors, ncu r01h)
void ConfigureKernels() {
    const int pct = 36;   // of 228 KB: >= 8 x 9 KB
    cudaFuncSetAttribute(k_trace_closest<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_closest<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_closest<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_closest<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_shadow<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_shadow<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_shadow<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    cudaFuncSetAttribute(k_trace_shadow<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
void LaunchResetAll(const RenderCtx &rc, cudaStream_t st) { k_reset_counts<<<1, 64, 0, st>>>(rc.counts, 0, Q_NCOUNTERS); }
void LaunchResetDepth(const RenderCtx &rc, int nextQ, cudaStream_t st) { k_reset_depth<<<1, 64, 0, st>>>(rc.counts, nextQ); }
void LaunchGenerateCameraRays(const RenderCtx &rc, int x0, int x1, int y0, int y1, int firstSample, int nSamples, int sms, cudaStream_t st) {
    k_generate_camera_rays<<<GridFor(sms, 8), 256, 0, st>>>(rc, x0, x1, y0, y1, firstSample, nSamples);
}
void LaunchTraceClosest(const RenderCtx &rc, int curQ, bool count, int sms, cudaStream_t st) {
    const bool inst = rc.bvh.instances != nullptr;
    if (count) { if (inst) k_trace_closest<true, true><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc, curQ); else k_trace_closest<true, false><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc, curQ); }
    else { if (inst) k_trace_closest<false, true><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc, curQ); else k_trace_closest<false, false><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc, curQ); }
}
void LaunchTraceShadow(const RenderCtx &rc, bool count, int sms, cudaStream_t st) {
    const bool inst = rc.bvh.instances != nullptr;
    if (count) { if (inst) k_trace_shadow<true, true><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc); else k_trace_shadow<true, false><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc); }
    else { if (inst) k_trace_shadow<false, true><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc); else k_trace_shadow<false, false><<<GridFor(sms, BPT_TRACE_CTAS), 128, 0, st>>>(rc); }
}
void LaunchHandleEmissive(const RenderCtx &rc, int sms, cudaStream_t st) { k_handle_emissive<<<GridFor(sms, 4), 128, 0, st>>>(rc); }
void LaunchHandleEscaped(const RenderCtx &rc, int sms, cudaStream_t st) { k_handle_escaped<<<GridFor(sms, 4), 128, 0, st>>>(rc); }
void LaunchShade(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st) {
    switch (materialTag) {
    case MAT_DIFFUSE: k_shade<DiffuseMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_CONDUCTOR: k_shade<ConductorMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_DIELECTRIC: k_shade<DielectricMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_COATED_CONDUCTOR: k_shade<CoatedConductorMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_COATED_DIFFUSE: k_shade<CoatedDiffuseMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_DIFFUSE_TRANSMISSION: k_shade<DiffuseTransmissionMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_THIN_DIELECTRIC: k_shade<ThinDielectricMaterialEval, false, kShapes><<<GridFor(sms, 2 * BPT_SHADE_CTAS), 128, 0, st>>>(rc, nextQ); break;
    case MAT_INTERFACE: k_shade<InterfaceMaterialEval, false, kShapes><<<GridFor(sms, 4), 128, 0, st>>>(rc, nextQ); break;
    default: break;
    }
}
bool ShadeSupported(int materialTag) {
    return materialTag == MAT_DIFFUSE || materialTag == MAT_INTERFACE || materialTag == MAT_CONDUCTOR || materialTag == MAT_DIELECTRIC || materialTag == MAT_COATED_CONDUCTOR ||
           materialTag == MAT_COATED_DIFFUSE || materialTag == MAT_DIFFUSE_TRANSMISSION || materialTag == MAT_THIN_DIELECTRIC;
}
void LaunchUpdateFilm(const RenderCtx &rc, int x0, int x1, int y0, int y1, int nSamples, int sms, cudaStream_t st) {
    k_update_film<<<GridFor(sms, 8), 256, 0, st>>>(rc, x0, x1, y0, y1, nSamples);
}
void LaunchResolveFilm(const double4 *pix, const DeviceFilm &film, float *rgbOut, int sms, cudaStream_t st) {
    k_resolve_film<<<GridFor(sms, 4), 256, 0, st>>>(pix, film, rgbOut);
}
void LaunchTraceRaw(const BVHView &bvh, const float4 *rays, int n, float4 *hits, float *tOut, int *cursor, unsigned long long *stats,
                    bool any, bool count, int sms, cudaStream_t st) {
    cudaMemsetAsync(cursor, 0, sizeof(int), st);
    int g = GridFor(sms, 8);
    if (any) { if (count) k_trace_raw<true, true><<<g, 128, 0, st>>>(bvh, rays, n, hits, tOut, cursor, stats);
               else k_trace_raw<true, false><<<g, 128, 0, st>>>(bvh, rays, n, hits, tOut, cursor, stats); }
    else { if (count) k_trace_raw<false, true><<<g, 128, 0, st>>>(bvh, rays, n, hits, tOut, cursor, stats);
           else k_trace_raw<false, false><<<g, 128, 0, st>>>(bvh, rays, n, hits, tOut, cursor, stats); }
}
void LaunchProbeSampler(const SamplerParams &sp, const int *pxy, const int *sidx, const int *dim, int n, float *o1, float *o2, cudaStream_t st) {
    k_probe_sampler<<<(n + 127) / 128, 128, 0, st>>>(sp, pxy, sidx, dim, n, o1, o2);
}
void LaunchProbeTri(const float *rays, const float *tris, int n, float *out, cudaStream_t st) {
    k_probe_tri<<<(n + 127) / 128, 128, 0, st>>>(rays, tris, n, out);
}

#if BPT_WITH_SHAPES
}  // namespace tu_shapes
#endif
}  // namespace bpt
