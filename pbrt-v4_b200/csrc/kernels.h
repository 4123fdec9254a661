// kernels.h -- host-callable launchers of the kernels in wavefront.cu.
#pragma once
#include "wavefront.cuh"

namespace bpt {
#define BPT_WAVEFRONT_LAUNCHERS \
    void UploadSobolTable(const uint32_t *t); \
    void UploadSamplerConstants(const SamplerParams &sp, cudaStream_t st); \
    void ConfigureKernels(); \
    void LaunchResetAll(const RenderCtx &rc, cudaStream_t st); \
    void LaunchResetDepth(const RenderCtx &rc, int nextQ, cudaStream_t st); \
    void LaunchGenerateCameraRays(const RenderCtx &rc, int x0, int x1, int y0, int y1, int firstSample, int nSamples, int sms, cudaStream_t st); \
    void LaunchTraceClosest(const RenderCtx &rc, int curQ, bool count, int sms, cudaStream_t st); \
    void LaunchTraceShadow(const RenderCtx &rc, bool count, int sms, cudaStream_t st); \
    bool LaunchRaySort(const RenderCtx &rc, int q, int n, const float bmin[3], const float bmax[3], unsigned *keys, unsigned *keysTmp, int *queueTmp, unsigned *hist, cudaStream_t st); \
    void LaunchZSobolPrefix(const SamplerParams &sp, uint32_t *table, int sms, cudaStream_t st); \
    void LaunchSampleMedium(const RenderCtx &rc, int nextQ, int sms, cudaStream_t st); \
    void LaunchTraceShadowTr(const RenderCtx &rc, int sms, cudaStream_t st); \
    void LaunchHandleEmissive(const RenderCtx &rc, int sms, cudaStream_t st); \
    void LaunchHandleEscaped(const RenderCtx &rc, int sms, cudaStream_t st); \
    void LaunchShade(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st); \
    bool ShadeSupported(int materialTag); \
    void LaunchUpdateFilm(const RenderCtx &rc, int x0, int x1, int y0, int y1, int nSamples, int sms, cudaStream_t st); \
    void LaunchResolveGBuffer(const double4 *pix, const GBufferPixel *gb, const DeviceFilm &film, float *out, int sms, cudaStream_t st); \
    void LaunchResolveFilm(const double4 *pix, const DeviceFilm &film, float *rgbOut, int sms, cudaStream_t st); \
    void LaunchTraceRaw(const BVHView &bvh, const float4 *rays, int n, float4 *hits, float *tOut, int *cursor, unsigned long long *stats, bool any, bool count, int sms, cudaStream_t st); \
    void LaunchProbeSampler(const SamplerParams &sp, const int *pxy, const int *sidx, const int *dim, int n, float *o1, float *o2, cudaStream_t st); \
    void LaunchProbeMath(int fn, const float *x, const float *y, int n, float *out, cudaStream_t st); \
    void LaunchProbeTri(const float *rays, const float *tris, int n, float *out, cudaStream_t st);
BPT_WAVEFRONT_LAUNCHERS
// shade_tex.cu: the TEX = true instantiations of k_shade (image textures, bump mapping, ray differentials) with their own copies of the
// sampler constants
void UploadSobolTableTex(const uint32_t *t);
void UploadSamplerConstantsTex(const SamplerParams &sp, cudaStream_t st);
void LaunchShadeTex(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st);
// shade_tex_part1 / 2 / 3.cu (reached through the three functions above)
void UploadSobolTableTexPart1(const uint32_t *t);
void UploadSamplerConstantsTexPart1(const SamplerParams &sp, cudaStream_t st);
void LaunchShadeTexPart1(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st);
void UploadSobolTableTexPart2(const uint32_t *t);
void UploadSamplerConstantsTexPart2(const SamplerParams &sp, cudaStream_t st);
void LaunchShadeTexPart2(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st);
void UploadSobolTableTexPart3(const uint32_t *t);
void UploadSamplerConstantsTexPart3(const SamplerParams &sp, cudaStream_t st);
void LaunchShadeTexPart3(const RenderCtx &rc, int materialTag, int nextQ, int sms, cudaStream_t st);
// wavefront_shapes.cu: the same launchers over the kernels compiled with quadric support
namespace tu_shapes {
BPT_WAVEFRONT_LAUNCHERS
}
}  // namespace bpt
