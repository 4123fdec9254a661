// materials.cuh -- Material::GetBxDF for the per-material shading kernels (one k_shade instantiation per
// material tag, mirroring the reference's per-material MaterialEvalQueue kernels, wavefront/surfscatter.cpp:26-55).
//   DiffuseMaterial::GetBxDF          materials.h:465-469
//   ConductorMaterial::GetBxDF        materials.h:491-513
//   DielectricMaterial::GetBxDF       materials.h:182-203
//   CoatedConductorMaterial::GetBxDF  materials.cpp:346-398
// Textures on this path are constant (SpectrumConstantTexture / FloatConstantTexture), i.e. what
// BasicTextureEvaluator handles (textures.h:1158-1213).
#pragma once
#include "scene.h"
#include "lights.h"
#include "bxdfs.h"

namespace bpt {

struct DiffuseMaterialEval {
    static constexpr int kTag = MAT_DIFFUSE;
    typedef DiffuseBxDF BxDF;
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &m, const SceneView &sv, Wavelengths &wl, bool *lambdaChanged) {
        Spec r = ClampSpec(EvalSpectrum(m.reflectance, sv.specPool, wl), 0, 1);
        return DiffuseBxDF{r};
    }
};

struct InterfaceMaterialEval {
    static constexpr int kTag = MAT_INTERFACE;
    typedef DiffuseBxDF BxDF;
    BPT_HD static BxDF GetBxDF(const DeviceMaterial &, const SceneView &, Wavelengths &, bool *) { return DiffuseBxDF{Spec(0.f)}; }
};

}  // namespace bpt
