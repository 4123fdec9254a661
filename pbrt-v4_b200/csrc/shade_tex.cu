// shade_tex.cu -- the TEX = true instantiations of k_shade (shade.cuh): materials whose parameters are image textures, bump mapping,
// and the ray / (u,v) differentials texture filtering needs (what the reference routes through UniversalTextureEvaluator,
// textures.h:1215-1240).  A separate translation unit so that scenes without textures keep the constant-parameter kernels'
// registers and instruction footprint (and so that the two sets compile in parallel); it carries its own copies of the sampler's
// constant-memory tables.
// Compiled four times (ptxas on these kernels is the build's long pole: 20 minutes in one object file): as it is (part 0) and through
// shade_tex_part1 / 2 / 3.cu, two materials per object file; each object file has its own constant tables, part 0 forwards to the others.
#ifndef BPT_TEX_PART
#define BPT_TEX_PART 0
#endif
#define BPT_SOBOL_DEVICE_TABLE
#include "shade.cuh"
#include "kernels.h"

namespace bpt {
#define TEXCAT2(a, b) a##b
#define TEXCAT(a, b) TEXCAT2(a, b)
#if BPT_TEX_PART == 0
#define TEXFN(name) name
#else
#define TEXFN(name) TEXCAT(name##Part, BPT_TEX_PART)
#endif

void TEXFN(UploadSobolTableTex)(const uint32_t *t) {
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
#if BPT_TEX_PART == 0
    UploadSobolTableTexPart1(t); UploadSobolTableTexPart2(t); UploadSobolTableTexPart3(t);
#endif
}

void TEXFN(UploadSamplerConstantsTex)(const SamplerParams &sp, cudaStream_t st) {
    static thread_local uint64_t hashes[kDimHashTable];
    for (int d = 0; d < kDimHashTable; ++d) hashes[d] = Hash_i32x2(d, sp.seed);
    cudaMemcpyToSymbolAsync(c_sampler, &sp, sizeof(SamplerParams), 0, cudaMemcpyHostToDevice, st);
    cudaMemcpyToSymbolAsync(c_dimHash, hashes, sizeof(hashes), 0, cudaMemcpyHostToDevice, st);
#if BPT_TEX_PART == 0
    UploadSamplerConstantsTexPart1(sp, st); UploadSamplerConstantsTexPart2(sp, st); UploadSamplerConstantsTexPart3(sp, st);
#endif
}

void TEXFN(LaunchShadeTex)(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st) {
    const int grid = sms * 2 * BPT_SHADE_CTAS;
    switch (materialTag) {
#if BPT_TEX_PART == 0
    case MAT_DIFFUSE: k_shade<DiffuseMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
    case MAT_INTERFACE: k_shade<InterfaceMaterialEval, true, true><<<sms * 4, 128, 0, st>>>(rc, nextQ); break;
    case MAT_COATED_CONDUCTOR: case MAT_COATED_DIFFUSE: LaunchShadeTexPart1(rc, materialTag, nextQ, sms, st); break;
    case MAT_CONDUCTOR: case MAT_DIELECTRIC: LaunchShadeTexPart2(rc, materialTag, nextQ, sms, st); break;
    case MAT_DIFFUSE_TRANSMISSION: case MAT_THIN_DIELECTRIC: LaunchShadeTexPart3(rc, materialTag, nextQ, sms, st); break;
#elif BPT_TEX_PART == 1
    case MAT_COATED_CONDUCTOR: k_shade<CoatedConductorMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
    case MAT_COATED_DIFFUSE: k_shade<CoatedDiffuseMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
#elif BPT_TEX_PART == 2
    case MAT_CONDUCTOR: k_shade<ConductorMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
    case MAT_DIELECTRIC: k_shade<DielectricMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
#else
    case MAT_DIFFUSE_TRANSMISSION: k_shade<DiffuseTransmissionMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
    case MAT_THIN_DIELECTRIC: k_shade<ThinDielectricMaterialEval, true, true><<<grid, 128, 0, st>>>(rc, nextQ); break;
#endif
    default: break;
    }
}

}  // namespace bpt
