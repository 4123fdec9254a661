// shade.cuh -- the surface-scattering kernel k_shade<Material, TEX> and the small path-state helpers it shares with the other
// wavefront kernels.  Two translation units instantiate it: wavefront.cu with TEX = false (materials whose parameters are constants,
// what the reference's BasicTextureEvaluator handles, textures.h:1158-1213) and shade_tex.cu with TEX = true (image textures, bump
// mapping and the ray / (u,v) differentials they need -- UniversalTextureEvaluator), so that scenes without textures keep the
// register budget and instruction footprint of the constant-parameter kernels.
//   k_shade<Material>         Li body cpu/integrators.cpp:1115-1268 + SampleLd :1273-1390 (wavefront/surfscatter.cpp:57-328)
#pragma once
#include "wavefront.cuh"
#include "bxdfs.h"
#include "materials.cuh"
#include "camera.h"
#include "differentials.h"

namespace bpt {

// ---------------------------------------------------------------- small helpers (transforms and the camera live in camera.h)
__device__ __forceinline__ Spec ld4(const float4 *p) { float4 v = *p; return Spec(v.x, v.y, v.z, v.w); }
__device__ __forceinline__ void st4(float4 *p, Spec s) { *p = make_float4(s[0], s[1], s[2], s[3]); }

__device__ __forceinline__ Wavelengths LoadLambda(const PathState &ps, int slot) {
    Wavelengths wl;
    float4 a = ps.lambda[slot], b = ps.lambdaPdf[slot];
    wl.lambda[0] = a.x; wl.lambda[1] = a.y; wl.lambda[2] = a.z; wl.lambda[3] = a.w;
    wl.pdf[0] = b.x; wl.pdf[1] = b.y; wl.pdf[2] = b.z; wl.pdf[3] = b.w;
    return wl;
}
__device__ __forceinline__ LightSampleContext LoadCtx(const PathState &ps, int slot) {
    float4 a = ps.ctx0[slot], b = ps.ctx1[slot], c = ps.ctx2[slot];
    LightSampleContext ctx;
    ctx.pi.x.low = a.x; ctx.pi.x.high = a.y; ctx.pi.y.low = a.z; ctx.pi.y.high = a.w;
    ctx.pi.z.low = b.x; ctx.pi.z.high = b.y;
    ctx.n = Nrm3(b.z, b.w, c.x);
    ctx.ns = Nrm3(c.y, c.z, c.w);
    return ctx;
}
__device__ __forceinline__ void StoreCtx(const PathState &ps, int slot, const Pt3i &pi, Nrm3 n, Nrm3 ns) {
    ps.ctx0[slot] = make_float4(pi.x.low, pi.x.high, pi.y.low, pi.y.high);
    ps.ctx1[slot] = make_float4(pi.z.low, pi.z.high, n.x, n.y);
    ps.ctx2[slot] = make_float4(n.z, ns.x, ns.y, ns.z);
}

// ---------------------------------------------------------------- queue reset
// Interaction::GetMedium(w) for a surface hit (interaction.h:117-121): the primitive's MediumInterface when it is a
// transition, the ray's medium otherwise (SurfaceInteraction::SetIntersectionProperties, interaction.h:218-229)
__device__ __forceinline__ int SurfaceGetMedium(const DeviceMesh &mesh, Vec3 w, Nrm3 n, int rayMedium) {
    if (mesh.mediumInside != mesh.mediumOutside) return Dot(w, n) > 0 ? mesh.mediumOutside : mesh.mediumInside;
    return rayMedium;
}

__device__ __forceinline__ void PushShadowTr(const RenderCtx &rc, int slot, Pt3 pf, Vec3 d, int medium, Spec bf, Spec Ls, Spec rl,
                                             Spec ru, const Pt3i &pLight, Nrm3 nLight) {
    int p = WarpAggPush(&rc.counts[Q_SHADOW]);
    ShadowTrItem it;
    it.o_slot = make_float4(pf.x, pf.y, pf.z, __int_as_float(slot));
    it.d_med = make_float4(d.x, d.y, d.z, __int_as_float(medium));
    it.bf = make_float4(bf[0], bf[1], bf[2], bf[3]);
    it.Ls = make_float4(Ls[0], Ls[1], Ls[2], Ls[3]);
    it.rl = make_float4(rl[0], rl[1], rl[2], rl[3]);
    it.ru = make_float4(ru[0], ru[1], ru[2], ru[3]);
    it.pl0 = make_float4(pLight.x.low, pLight.x.high, pLight.y.low, pLight.y.high);
    it.pl1 = make_float4(pLight.z.low, pLight.z.high, nLight.x, nLight.y);
    it.pl2 = make_float4(nLight.z, 0.f, 0.f, 0.f);
    rc.shadowTr[p] = it;
}

// ---------------------------------------------------------------- surface scattering, one instantiation per material
#ifndef BPT_SHADE_CTAS
#define BPT_SHADE_CTAS 4      // resident 128-thread CTAs per SM the shading kernels are register-limited for (128 registers: no spills;
                              // r02d sweep with the outlined sampler / libm: 3 -> ?, 4 -> 23.0 ms, 5 -> 25.9, 6 -> 26.2, 8 -> 26.4 per 64 spp)
#endif
// Path flags in misc.w beyond the depth byte: 0x100 specularBounce, 0x200 anyNonSpecularBounces, and for scenes with textures the
// state of the ray's differentials: 0x400 the ray has none (spawned at a non-specular bounce or a medium scattering event),
// 0x800 they are in ps.rdiff (written at a specular bounce or when an interface surface was skipped); neither: the camera ray's own,
// regenerated from the camera sample.
constexpr unsigned kPathNoDiffs = 0x400u, kPathStoredDiffs = 0x800u;

__device__ __forceinline__ RayDiffState LoadRayDiffs(const RenderCtx &rc, int slot, const uint4 &misc) {
    RayDiffState r;
    if (misc.w & kPathNoDiffs) return r;
    r.has = true;
    if (misc.w & kPathStoredDiffs) {
        const size_t M = (size_t)rc.maxPaths;
        float4 a = rc.ps.rdiff[slot], b = rc.ps.rdiff[M + slot], c = rc.ps.rdiff[2 * M + slot], d = rc.ps.rdiff[3 * M + slot];
        r.rxO = Pt3(a.x, a.y, a.z); r.ryO = Pt3(b.x, b.y, b.z); r.rxD = Vec3(c.x, c.y, c.z); r.ryD = Vec3(d.x, d.y, d.z);
        return r;
    }
    Sampler camSampler;
    CamSample cs = GetCameraSample(rc.sampler, rc.filter, camSampler, misc.x & 0xffff, misc.x >> 16, (int)misc.y);
    Pt3 oCam; Vec3 dCam;
    GenerateCameraRay(rc.camera, cs.pFilm, cs.pLens, &oCam, &dCam);
    RayDiffs rd = CameraRayDifferentials(rc.camera, cs, oCam, dCam);
    r.rxO = rd.rxO; r.ryO = rd.ryO; r.rxD = rd.rxD; r.ryD = rd.ryD;
    return r;
}
__device__ __forceinline__ void StoreRayDiffs(const RenderCtx &rc, int slot, const RayDiffState &r) {
    const size_t M = (size_t)rc.maxPaths;
    rc.ps.rdiff[slot] = make_float4(r.rxO.x, r.rxO.y, r.rxO.z, 0.f);
    rc.ps.rdiff[M + slot] = make_float4(r.ryO.x, r.ryO.y, r.ryO.z, 0.f);
    rc.ps.rdiff[2 * M + slot] = make_float4(r.rxD.x, r.rxD.y, r.rxD.z, 0.f);
    rc.ps.rdiff[3 * M + slot] = make_float4(r.ryD.x, r.ryD.y, r.ryD.z, 0.f);
}

template <typename Material, bool TEX, bool SHAPES>
__global__ void __launch_bounds__(128, BPT_SHADE_CTAS) k_shade(RenderCtx rc, int nextQ) {
    const int n = rc.counts[Q_MAT0 + Material::kTag];
    const int *queue = rc.queues[Q_MAT0 + Material::kTag];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int slot = queue[i];
        const PathState &ps = rc.ps;
        float4 h = ps.hit[slot];
        int tri = __float_as_int(h.x);
        TriData td = LoadTriangle(rc.sv, tri);
        SurfaceExtras ex;
        SurfaceIntr isect = HitInteraction<SHAPES>(rc.sv, td, h.y, h.z, h.w, TEX ? &ex : nullptr);
        const DeviceMesh &mesh = rc.sv.meshes[td.mesh];
        float4 d4 = ps.rayD[slot];
        // isect.wo: the Interaction constructor normalises it (interaction.h:32-33, wo(Normalize(wo))) -- not a no-op in float: a third
        // of all unit vectors change in the last bit when normalised again, which the microfacet BxDFs see (r02c bisection).
        // A quadric takes -ray.d to object space first and normalises there and on the way back (HitWo).
        Vec3 wo = HitWo<SHAPES>(rc.sv, td, Vec3(d4.x, d4.y, d4.z));
        if (rc.sv.instances) {
            const DeviceInstance &I = rc.sv.instances[ps.hitInst[slot]];
            if (!I.identity) {
                // ret.wo = Normalize(t(si.wo)), si.wo the interaction's wo in instance space (cpu/primitive.cpp:114-123)
                isect = XformInteraction(I, isect); wo = Normalize(XformVec(I.m, HitWo<SHAPES>(rc.sv, td, XformVec(I.mInv, Vec3(d4.x, d4.y, d4.z)))));
                if (TEX) ex = XformExtras(I, ex);
            }
        }
        Wavelengths wl = LoadLambda(ps, slot);
        uint4 misc = ps.misc[slot];
        int depth = misc.w & 0xff;
        bool anyNonSpecularBounces = (misc.w >> 9) & 1;
        int px = misc.x & 0xffff, py = misc.x >> 16;

        const int rayMedium = rc.nMedia > 0 ? ps.medium[slot] : -1;
        if (Material::kTag == MAT_INTERFACE) {
            // isect.SkipIntersection(&ray, tHit): respawn the ray past the surface, same depth (interaction.h)
            const Vec3 rayDir(d4.x, d4.y, d4.z);   // SpawnRay(ray.d): the ray's own (not re-normalised) direction
            if (rc.film.isGBuffer && depth == 0) {
                // SurfaceInteraction::SkipIntersection (interaction.cpp:91-97): rxOrigin += t * rxDirection, ryOrigin likewise.  The running
                // origins live in the (still unused) dpdx / dpdy slots of the visible-surface record, .w = 1 once initialised.
                const size_t M = (size_t)rc.maxPaths;
                Sampler camSampler;
                CamSample cs = GetCameraSample(rc.sampler, rc.filter, camSampler, px, py, (int)misc.y);
                Pt3 oCam; Vec3 dCam;
                GenerateCameraRay(rc.camera, cs.pFilm, cs.pLens, &oCam, &dCam);
                RayDiffs rdif = CameraRayDifferentials(rc.camera, cs, oCam, dCam);
                float4 sx = ps.vs[3 * M + slot], sy = ps.vs[4 * M + slot];
                Pt3 rxO = sx.w != 0 ? Pt3(sx.x, sx.y, sx.z) : rdif.rxO, ryO = sy.w != 0 ? Pt3(sy.x, sy.y, sy.z) : rdif.ryO;
                const float t = ps.hitT[slot];
                rxO = rxO + rdif.rxD * t; ryO = ryO + rdif.ryD * t;
                ps.vs[3 * M + slot] = make_float4(rxO.x, rxO.y, rxO.z, 1.f);
                ps.vs[4 * M + slot] = make_float4(ryO.x, ryO.y, ryO.z, 1.f);
            }
            if (TEX) {
                // the same advance for the differentials texture filtering will use at the next real surface
                RayDiffState rdiffs = LoadRayDiffs(rc, slot, misc);
                if (rdiffs.has) {
                    const float t = ps.hitT[slot];
                    rdiffs.rxO = rdiffs.rxO + rdiffs.rxD * t; rdiffs.ryO = rdiffs.ryO + rdiffs.ryD * t;
                    StoreRayDiffs(rc, slot, rdiffs);
                    ps.misc[slot] = make_uint4(misc.x, misc.y, misc.z, misc.w | kPathStoredDiffs);
                }
            }
            Pt3 o = OffsetRayOrigin(isect.pi, isect.n, rayDir);
            ps.rayO[slot] = make_float4(o.x, o.y, o.z, 0.f);
            if (rc.nMedia > 0) ps.medium[slot] = SurfaceGetMedium(mesh, rayDir, isect.n, rayMedium);
            int p = WarpAggPush(&rc.counts[nextQ]);
            rc.queues[nextQ][p] = slot;
            continue;
        }
        // Get BSDF (material.GetBxDF may terminate secondary wavelengths)
        const DeviceMaterial &mat = rc.sv.materials[mesh.material];
        bool lambdaChanged = false;
        // SurfaceInteraction::GetBSDF (interaction.cpp:159-199): differentials, bump mapping, then the material's BxDF
        RayDiffState rdiffs;
        UVDiffs uvd;
        TexEvalCtx tc;
        if (TEX) {
            rdiffs = LoadRayDiffs(rc, slot, misc);
            uvd = ComputeDifferentials(rc.camera, isect, ex, rdiffs);
            tc.p = isect.p(); tc.dpdx = uvd.dpdx; tc.dpdy = uvd.dpdy; tc.n = isect.n; tc.uv = isect.uv;
            tc.dudx = uvd.dudx; tc.dudy = uvd.dudy; tc.dvdx = uvd.dvdx; tc.dvdy = uvd.dvdy;
            if (mat.normalMapTex >= 0) {
                // NormalMap (materials.h:86-105): tangent-space normal from the image (bilinear, repeat), then dpdu / dpdv that give it
                const DeviceTexture &nm = rc.sv.textures[mat.normalMapTex];
                const DeviceMIPLevel &lv = rc.sv.mipLevels[nm.firstLevel];
                Vec2 uvn(tc.uv.x, 1 - tc.uv.y);
                Vec3 nsl(2 * MipBilerpChannel(lv, 3, uvn, 0, TEXWRAP_REPEAT) - 1, 2 * MipBilerpChannel(lv, 3, uvn, 1, TEXWRAP_REPEAT) - 1,
                         2 * MipBilerpChannel(lv, 3, uvn, 2, TEXWRAP_REPEAT) - 1);
                nsl = Normalize(nsl);
                Frame frame = Frame::FromXZ(Normalize(isect.dpdus), V(isect.ns));
                nsl = frame.FromLocal(nsl);
                float ulen = Length(isect.dpdus), vlen = Length(ex.dpdvs);
                Vec3 dpdu = Normalize(GramSchmidt(isect.dpdus, nsl)) * ulen;
                Vec3 dpdv = Normalize(Cross(nsl, dpdu)) * vlen;
                Nrm3 ns(Normalize(Cross(dpdu, dpdv)));
                isect.ns = FaceForward(ns, isect.n);
                while (LengthSquared(dpdu) > 1e16f || LengthSquared(dpdv) > 1e16f) { dpdu = dpdu / 1e8f; dpdv = dpdv / 1e8f; }
                isect.dpdus = dpdu; ex.dpdvs = dpdv;
            } else if (mat.displacementTex >= 0) {
                // BumpMap (materials.h:107-138) + SetShadingGeometry(ns, dpdu, dpdv, shading.dndu, shading.dndv, false)
                const int disp = mat.displacementTex;
                TexEvalCtx shifted = tc;
                float du = .5f * (fabsf(tc.dudx) + fabsf(tc.dudy));
                if (du == 0) du = .0005f;
                shifted.p = tc.p + du * isect.dpdus;
                shifted.uv = Vec2(tc.uv.x + du, tc.uv.y + 0.f);
                float uDisplace = EvalFloatTexture<kTexMaxDepth>(rc.sv, disp, shifted);
                float dv = .5f * (fabsf(tc.dvdx) + fabsf(tc.dvdy));
                if (dv == 0) dv = .0005f;
                shifted.p = tc.p + dv * ex.dpdvs;
                shifted.uv = Vec2(tc.uv.x + 0.f, tc.uv.y + dv);
                float vDisplace = EvalFloatTexture<kTexMaxDepth>(rc.sv, disp, shifted);
                float displace = EvalFloatTexture<kTexMaxDepth>(rc.sv, disp, tc);
                Vec3 dpdu = isect.dpdus + (uDisplace - displace) / du * V(isect.ns) + displace * V(ex.dndus);
                Vec3 dpdv = ex.dpdvs + (vDisplace - displace) / dv * V(isect.ns) + displace * V(ex.dndvs);
                Nrm3 ns(Normalize(Cross(dpdu, dpdv)));
                isect.ns = FaceForward(ns, isect.n);
                while (LengthSquared(dpdu) > 1e16f || LengthSquared(dpdv) > 1e16f) { dpdu = dpdu / 1e8f; dpdv = dpdv / 1e8f; }
                isect.dpdus = dpdu; ex.dpdvs = dpdv;
            }
        }
        typename Material::BxDF bxdf = Material::template GetBxDF<TEX>(mat, rc.sv, wl, &lambdaChanged, TEX ? &tc : nullptr);
        BSDF<typename Material::BxDF> bsdf(isect.ns, isect.dpdus, bxdf);
        if (lambdaChanged) ps.lambdaPdf[slot] = make_float4(wl.pdf[0], wl.pdf[1], wl.pdf[2], wl.pdf[3]);

        // Initialize visibleSurf at first intersection (cpu/integrators.cpp:1118-1145): GBufferFilm only
        if (rc.film.isGBuffer && depth == 0) {
            Sampler camSampler;
            CamSample cs = GetCameraSample(rc.sampler, rc.filter, camSampler, px, py, (int)misc.y);
            // the differentials were scaled about the CAMERA ray (ScaleDifferentials right after GenerateRayDifferential) and
            // SkipIntersection leaves them alone when it respawns the ray behind an interface surface (interaction.h): regenerate
            // the camera ray instead of reading the current segment's origin
            Pt3 oCam; Vec3 rd;
            GenerateCameraRay(rc.camera, cs.pFilm, cs.pLens, &oCam, &rd);
            RayDiffs rdif = CameraRayDifferentials(rc.camera, cs, oCam, rd);
            if (rc.storeHitT) {   // differentials advanced by interface crossings (k_shade<Interface>)
                const size_t M = (size_t)rc.maxPaths;
                float4 sx = ps.vs[3 * M + slot], sy = ps.vs[4 * M + slot];
                if (sx.w != 0) rdif.rxO = Pt3(sx.x, sx.y, sx.z);
                if (sy.w != 0) rdif.ryO = Pt3(sy.x, sy.y, sy.z);
            }
            // SurfaceInteraction::ComputeDifferentials (interaction.cpp:42-60), the dpdx / dpdy part
            Vec3 dpdx(0, 0, 0), dpdy(0, 0, 0);
            Pt3 P = isect.p();
            if (Dot(isect.n, rdif.rxD) != 0 && Dot(isect.n, rdif.ryD) != 0) {
                float dd = -Dot(isect.n, Vec3(P.x, P.y, P.z));
                float tx = (-Dot(isect.n, Vec3(rdif.rxO.x, rdif.rxO.y, rdif.rxO.z)) - dd) / Dot(isect.n, rdif.rxD);
                Pt3 pxp = rdif.rxO + rdif.rxD * tx;
                float ty = (-Dot(isect.n, Vec3(rdif.ryO.x, rdif.ryO.y, rdif.ryO.z)) - dd) / Dot(isect.n, rdif.ryD);
                Pt3 pyp = rdif.ryO + rdif.ryD * ty;
                dpdx = pxp - P; dpdy = pyp - P;
            }   // (else: Camera::Approximate_dp_dxy -- a measure-zero case, left at zero)
            // albedo: bsdf.rho(wo, ucRho, uRho) with the 16 fixed samples (BxDF::rho bxdfs.cpp:1131-1144)
            const float ucRho[16] = {0.75741637f, 0.37870818f, 0.7083487f, 0.18935409f, 0.9149363f, 0.35417435f, 0.5990858f, 0.09467703f,
                                     0.8578725f, 0.45746812f, 0.686759f, 0.17708716f, 0.9674518f, 0.2995429f, 0.5083201f, 0.047338516f};
            const float uRho[16][2] = {{0.855985f, 0.570367f}, {0.381823f, 0.851844f}, {0.285328f, 0.764262f}, {0.733380f, 0.114073f},
                                       {0.542663f, 0.344465f}, {0.127274f, 0.414848f}, {0.964700f, 0.947162f}, {0.594089f, 0.643463f},
                                       {0.095109f, 0.170369f}, {0.825444f, 0.263359f}, {0.429467f, 0.454469f}, {0.244460f, 0.816459f},
                                       {0.756135f, 0.731258f}, {0.516165f, 0.152852f}, {0.180888f, 0.214174f}, {0.898579f, 0.503897f}};
            Spec albedo(0.f);
            Vec3 woL = bsdf.RenderToLocal(wo);
            if (woL.z != 0) {
                Spec r(0.f);
                for (int k = 0; k < 16; ++k) {
                    BSDFSample bs = bsdf.bxdf.Sample_f(woL, ucRho[k], Vec2(uRho[k][0], uRho[k][1]), MODE_RADIANCE);
                    if (bs.valid && bs.pdf > 0) r += bs.f * AbsCosTheta(bs.wi) / bs.pdf;
                }
                albedo = r / 16.f;
            }
            Nrm3 vn = FaceForward(isect.n, wo), vns = FaceForward(isect.ns, wo);
            const size_t M = (size_t)rc.maxPaths;
            ps.vs[slot] = make_float4(P.x, P.y, P.z, 1.f);
            ps.vs[M + slot] = make_float4(vn.x, vn.y, vn.z, isect.uv.x);
            ps.vs[2 * M + slot] = make_float4(vns.x, vns.y, vns.z, isect.uv.y);
            ps.vs[3 * M + slot] = make_float4(dpdx.x, dpdx.y, dpdx.z, 0.f);
            ps.vs[4 * M + slot] = make_float4(dpdy.x, dpdy.y, dpdy.z, 0.f);
            ps.vs[5 * M + slot] = make_float4(albedo[0], albedo[1], albedo[2], albedo[3]);
        }
        // Terminate path if maximum depth reached
        if (depth++ >= rc.integ.maxDepth) continue;
        if (rc.integ.regularize && anyNonSpecularBounces) bsdf.Regularize();

        Sampler sampler;
        sampler.StartPixelSample(&rc.sampler, px, py, (int)misc.y, (int)misc.z);
        Spec beta = ld4(&ps.beta[slot]), r_u = ld4(&ps.r_u[slot]);
        int flags = bsdf.Flags();

        // Sample illumination from lights (SampleLd)
        if (IsNonSpecular(flags)) {
            LightSampleContext ctx;
            ctx.pi = isect.pi; ctx.n = isect.n; ctx.ns = isect.ns;
            if (IsReflective(flags) && !IsTransmissive(flags)) ctx.pi = Pt3i(OffsetRayOrigin(isect.pi, isect.n, wo));
            else if (IsTransmissive(flags) && !IsReflective(flags)) ctx.pi = Pt3i(OffsetRayOrigin(isect.pi, isect.n, -wo));
            float u = sampler.Get1D();
            float lightPMF = 0;
            int lightIndex = LightSamplerSample(rc.sv, ctx, u, &lightPMF);
            Vec2 uLight = sampler.Get2D();
            if (lightIndex >= 0) {
                const DeviceLight &light = rc.sv.lights[lightIndex];
                LightLiSample ls;
                if (light.type == LIGHT_DIFFUSE_AREA) ls = AreaLightSampleLi<SHAPES>(rc.sv, light, ctx, uLight, wl);
                else if (IsDeltaLight(light.type)) ls = DeltaLightSampleLi(rc.sv.specPool, light, ctx.p(), wl);
                else ls = InfiniteLightSampleLi(rc.sv, light, ctx, uLight, wl);
                if (ls.valid && ls.L.IsNonZero() && ls.pdf != 0) {
                    float p_l = lightPMF * ls.pdf;
                    Vec3 wi = ls.wi;
                    Spec f_hat = bsdf.f(wo, wi) * AbsDot(wi, isect.ns);
                    // delta lights: the estimate divides by r_l.Average() alone (cpu/integrators.cpp:1385-1388); a zero scattering
                    // pdf makes r_u vanish in (r_l + r_u) and gives the same bits
                    float scatterPDF = IsDeltaLight(light.type) ? 0.f : bsdf.PDF(wo, wi);
                    if (f_hat.IsNonZero()) {
                        // lightRay = intr.SpawnRayTo(ls->pLight)  (ray.h:94-101, both ends offset)
                        Pt3 pf, pt;
                        if (light.type == LIGHT_DIFFUSE_AREA) {
                            pf = OffsetRayOrigin(isect.pi, isect.n, ls.pLight.p() - isect.pi.p());
                            pt = OffsetRayOrigin(ls.pLight, ls.nLight, pf - ls.pLight.p());
                        } else {
                            // infinite light: pLight is a plain point with n = 0: OffsetRayOrigin leaves it unchanged
                            pf = OffsetRayOrigin(isect.pi, isect.n, ls.pLight.p() - isect.pi.p());
                            pt = ls.pLight.p();
                        }
                        Vec3 rdv = pt - pf;
                        if (rc.useTr) {
                            Nrm3 nL = light.type == LIGHT_DIFFUSE_AREA ? ls.nLight : Nrm3(0, 0, 0);
                            PushShadowTr(rc, slot, pf, rdv, SurfaceGetMedium(mesh, rdv, isect.n, rayMedium), beta * f_hat, ls.L,
                                         r_u * p_l, r_u * scatterPDF, ls.pLight, nL);
                        } else if (rdv != Vec3(0, 0, 0)) {
                            Spec r_l = r_u * p_l;
                            Spec r_uu = r_u * scatterPDF;
                            Spec Ld = beta * f_hat * ls.L / (r_l + r_uu).Average();
                            int p = WarpAggPush(&rc.counts[Q_SHADOW]);
                            ShadowItem it;
                            it.o_slot = make_float4(pf.x, pf.y, pf.z, __int_as_float(slot));
                            it.d = make_float4(rdv.x, rdv.y, rdv.z, 1 - kShadowEpsilon);
                            it.Ld = make_float4(Ld[0], Ld[1], Ld[2], Ld[3]);
                            rc.shadowItems[p] = it;
                        } else {
                            // zero-length light ray: the reference's while loop does not run and the contribution is added
                            Spec r_l = r_u * p_l;
                            Spec r_uu = r_u * scatterPDF;
                            Spec Ld = beta * f_hat * ls.L / (r_l + r_uu).Average();
                            float4 L = ps.L[slot];
                            L.x += Ld[0]; L.y += Ld[1]; L.z += Ld[2]; L.w += Ld[3];
                            ps.L[slot] = L;
                        }
                    }
                }
            }
        }
        // prevIntrContext = LightSampleContext(isect)
        StoreCtx(ps, slot, isect.pi, isect.n, isect.ns);

        // Sample BSDF to get new path direction
        float u = sampler.Get1D();
        Vec2 u2 = sampler.Get2D();
        BSDFSample bs = bsdf.Sample_f(wo, u, u2);
        if (!bs.valid) continue;
        beta *= bs.f * AbsDot(bs.wi, isect.ns) / bs.pdf;
        Spec r_l;
        if (bs.pdfIsProportional) r_l = r_u / bsdf.PDF(wo, bs.wi);
        else r_l = r_u / bs.pdf;
        bool specularBounce = IsSpecular(bs.flags);
        anyNonSpecularBounces |= !IsSpecular(bs.flags);
        float2 aux = ps.aux[slot];
        float etaScale = aux.x;
        if (IsTransmissive(bs.flags)) etaScale *= Sqr(bs.eta);
        // ray = isect.SpawnRay(ray, bsdf, bs->wi, bs->flags, bs->eta)
        Pt3 o = OffsetRayOrigin(isect.pi, isect.n, bs.wi);
        unsigned diffBits = 0;
        if (TEX) {
            RayDiffState next = SpawnRayDifferentials(rdiffs, isect.p(), wo, bs.wi, isect.ns, ex.dndus, ex.dndvs, uvd, bs.flags, bs.eta);
            if (next.has) { StoreRayDiffs(rc, slot, next); diffBits = kPathStoredDiffs; }
            else diffBits = kPathNoDiffs;
        }

        // Russian roulette
        if (!beta.IsNonZero()) continue;
        Spec rrBeta = beta * etaScale / r_u.Average();
        float uRR = sampler.Get1D();
        if (rrBeta.MaxComponentValue() < 1 && depth > 1) {
            float q = stdmax(0.f, 1 - rrBeta.MaxComponentValue());
            if (uRR < q) continue;
            beta /= 1 - q;
        }
        ps.rayO[slot] = make_float4(o.x, o.y, o.z, 0.f);
        ps.rayD[slot] = make_float4(bs.wi.x, bs.wi.y, bs.wi.z, d4.w);
        if (rc.nMedia > 0) ps.medium[slot] = SurfaceGetMedium(mesh, bs.wi, isect.n, rayMedium);   // SpawnRay: GetMedium(wi)
        st4(&ps.beta[slot], beta);
        st4(&ps.r_l[slot], r_l);
        ps.aux[slot] = make_float2(etaScale, aux.y);
        ps.misc[slot] = make_uint4(misc.x, misc.y, (unsigned)sampler.dimension,
                                   (unsigned)depth | (specularBounce ? 0x100u : 0u) | (anyNonSpecularBounces ? 0x200u : 0u) | diffBits);
        int p = WarpAggPush(&rc.counts[nextQ]);
        rc.queues[nextQ][p] = slot;
    }
}

}  // namespace bpt
