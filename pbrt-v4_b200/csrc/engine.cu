// engine.cu -- implementation of the C ABI in include/b200pt.h: scene intake, host-side scene preparation
// (restating what BasicScene::Create* and the WavefrontPathIntegrator constructor do, wavefront/integrator.cpp:80-287),
// device upload, BVH build and the wavefront render loop (wavefront/integrator.cpp:290-493).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#include "../../include/b200pt.h"
#include "bvh_build.h"
#include "scene_xforms.h"
#include "texture_host.h"
#include "texture.h"
#include "camera.h"
#include "exr_writer.h"
#include "kernels.h"
#include "lightbvh.h"
#include "tables.h"
#include "xform.h"
#include "lights.h"

using namespace bpt;

namespace {

struct OwnedSpectrum {
    bpt_spectrum s;
    std::vector<float> samples;
    void Set(const bpt_spectrum &in) {
        s = in;
        samples.clear();
        if (in.kind == BPT_SPEC_SAMPLED && in.samples && in.n_samples > 0) samples.assign(in.samples, in.samples + 2 * (size_t)in.n_samples);
        s.samples = nullptr;
    }
};
struct OwnedMaterial {
    bpt_material m;
    OwnedSpectrum reflectance, eta, k, iface_eta, conductor_eta, conductor_k, albedo, transmittance;
};
struct OwnedMesh {
    bpt_mesh m;
    std::vector<float> P, N, UV;
    std::vector<int> idx;
    bool isQuadric = false;      // one Sphere / Disk / Cylinder: m carries the common members (n_triangles = 1: its bounds record)
    bpt_quadric q{};
    bool isBilinear = false;     // bilinear patch mesh: idx holds four vertex indices per patch, m.n_triangles = number of patches
};
struct OwnedLight {
    bpt_light l;
    OwnedSpectrum L;
    std::vector<float> image;
};
struct OwnedAreaLight { bpt_area_light a; OwnedSpectrum L; };
struct OwnedTexture { bpt_texture t; HostImage img; OwnedSpectrum value; };
struct OwnedMedium { bpt_medium m; OwnedSpectrum sigma_a, sigma_s, Le; std::vector<float> density, leScale, rgbA, rgbS, rgbLe; };

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t count) {
        if (p && n == count) return cudaSuccess;   // re-commit of a same-sized scene keeps its allocations (cudaMalloc of ~1 GB is not free)
        if (p) { cudaFree(p); p = nullptr; }
        n = count;
        return cudaMalloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
    }
    cudaError_t upload(const std::vector<T> &v) {
        cudaError_t e = alloc(v.size());
        if (e != cudaSuccess) return e;
        if (!v.empty()) e = cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
};

}  // namespace

// Large meshes are assembled on the device: the object-space vertices and the index buffer are uploaded as they come
// (12 + 12 B per triangle for a grid mesh instead of the 48 B leaf-order-independent record) and this kernel applies
// renderFromObject -- Transform::operator()(Point3f), util/transform.h, the same unfused float expression as Xform::operator() --
// and writes the triangle pool records {p0 | mesh id, p1 | triangle index, p2 | class bits}.
__global__ void k_assemble_triangles(const float *__restrict__ P, const int *__restrict__ idx, int nTris, Mat4 M, uint32_t meshId,
                                     uint32_t cls, float4 *__restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTris) return;
    const float (*m)[4] = M.m;
    float4 r[3];
    for (int k = 0; k < 3; ++k) {
        int v = idx[3 * (size_t)t + k];
        float x = P[3 * (size_t)v], y = P[3 * (size_t)v + 1], z = P[3 * (size_t)v + 2];
        float xp = m[0][0] * x + m[0][1] * y + m[0][2] * z + m[0][3];
        float yp = m[1][0] * x + m[1][1] * y + m[1][2] * z + m[1][3];
        float zp = m[2][0] * x + m[2][1] * y + m[2][2] * z + m[2][3];
        float wp = m[3][0] * x + m[3][1] * y + m[3][2] * z + m[3][3];
        if (wp != 1) { xp = xp / wp; yp = yp / wp; zp = zp / wp; }
        r[k] = make_float4(xp, yp, zp, 0.f);
    }
    r[0].w = __uint_as_float(meshId); r[1].w = __uint_as_float((uint32_t)t); r[2].w = __uint_as_float(cls);
    out[3 * (size_t)t] = r[0]; out[3 * (size_t)t + 1] = r[1]; out[3 * (size_t)t + 2] = r[2];
}

// Tight bounds of an instance for the top-level BVH: the object's triangles taken through renderFromInstance vertex by vertex.
// The reference bounds a TransformedPrimitive by the transformed corners of the object's AABB ((*renderFromPrimitive)(primitive.Bounds()),
// cpu/primitive.cpp:106-110), which for a rotated instance of a round object is up to sqrt(3) larger per axis -- and every false
// overlap is an instance entry (ray transform + BLAS root visit) at the bottom of the traversal's lane utilisation (ncu r02g: 46 % of
// the two-level kernel's instructions run with 4-8 lanes).  Bounds only cull: the hit found is the same (tests compare bit for bit).
// The scene bounds the infinite lights are preprocessed with stay the reference's (union of the loose bounds).
__global__ void k_instance_bounds(const float4 *__restrict__ triPool, const DeviceInstance *__restrict__ inst, const int *__restrict__ instFirstTri,
                                  const int *__restrict__ instTriCount, int nInst, float4 *__restrict__ tlasPrims) {
    const int i = blockIdx.x;
    if (i >= nInst) return;
    const float (*m)[4] = inst[i].m.m;
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    const float4 *t = triPool + 3 * (size_t)instFirstTri[i];
    const int nv = 3 * instTriCount[i];
    for (int v = threadIdx.x; v < nv; v += blockDim.x) {
        float4 p = t[v];
        float xp = m[0][0] * p.x + m[0][1] * p.y + m[0][2] * p.z + m[0][3];
        float yp = m[1][0] * p.x + m[1][1] * p.y + m[1][2] * p.z + m[1][3];
        float zp = m[2][0] * p.x + m[2][1] * p.y + m[2][2] * p.z + m[2][3];
        float wp = m[3][0] * p.x + m[3][1] * p.y + m[3][2] * p.z + m[3][3];
        if (wp != 1) { xp /= wp; yp /= wp; zp /= wp; }
        lo[0] = fminf(lo[0], xp); lo[1] = fminf(lo[1], yp); lo[2] = fminf(lo[2], zp);
        hi[0] = fmaxf(hi[0], xp); hi[1] = fmaxf(hi[1], yp); hi[2] = fmaxf(hi[2], zp);
    }
    __shared__ float s[6][32];
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if ((threadIdx.x & 31) == 0) for (int k = 0; k < 3; ++k) { s[k][w] = lo[k]; s[3 + k][w] = hi[k]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int ww = 1; ww < nw; ++ww) for (int k = 0; k < 3; ++k) { lo[k] = fminf(lo[k], s[k][ww]); hi[k] = fmaxf(hi[k], s[3 + k][ww]); }
        // the hit is computed in instance space from the inverse-transformed ray: leave room for the rounding of this forward transform
        for (int k = 0; k < 3; ++k) { float pad = 1e-5f * fmaxf(fmaxf(fabsf(lo[k]), fabsf(hi[k])), hi[k] - lo[k]); lo[k] -= pad; hi[k] += pad; }
        float4 old0 = tlasPrims[3 * (size_t)i];
        // never larger than the reference-style bounds already there (the intersection of two valid bounds is a valid bound)
        float4 old1 = tlasPrims[3 * (size_t)i + 1];
        lo[0] = fmaxf(lo[0], old0.x); lo[1] = fmaxf(lo[1], old0.y); lo[2] = fmaxf(lo[2], old0.z);
        hi[0] = fminf(hi[0], old1.x); hi[1] = fminf(hi[1], old1.y); hi[2] = fminf(hi[2], old1.z);
        tlasPrims[3 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], old0.w);
        tlasPrims[3 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
        tlasPrims[3 * (size_t)i + 2] = make_float4(lo[0], lo[1], lo[2], 0.f);
    }
}

struct bpt_engine {
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    std::string err;
    Tables tables;
    // description
    bpt_camera cam{};
    bpt_film film{};
    bpt_filter filter{};
    bpt_sampler sampler{};
    bpt_integrator integ{};
    bool haveCam = false, haveFilm = false;
    std::vector<OwnedMaterial> materials;
    std::vector<OwnedAreaLight> areaLights;
    std::vector<OwnedMesh> meshes;
    std::vector<bpt_instance> instances;
    std::vector<OwnedLight> lights;
    std::vector<OwnedMedium> media;
    int cameraMedium = -1;
    int space = BPT_SPACE_WORLD;     // how the world_from_* members of the scene calls are to be read (bpt_scene_set_transform_space)
    // device
    DevBuf<float4> triPool;
    DevBuf<float> rawP;              // staging of object-space vertices / indices of device-assembled meshes
    DevBuf<int> rawIdx;
    DevBuf<DeviceMesh> dMeshes;
    DevBuf<float> dMeshAlpha;        // per mesh: constant alpha of the stochastic alpha test (1 = none)
    DevBuf<DeviceShape> dShapes;     // quadrics
    bool haveShapes = false;         // the scene has quadrics: the kernel set of wavefront_shapes.cu runs it
    DevBuf<int4> dMeshShape;         // per mesh: { quadric index | -1, first patch | -1, first pool record relative to its BVH, 0 }
    DevBuf<float4> dPatches;         // bilinear patches, 4 float4 each
    DevBuf<int> indexPool;
    DevBuf<float> nPool, uvPool, specPool;
    DevBuf<DeviceMaterial> dMaterials;
    DevBuf<uint32_t> dPmjSamples;
    DevBuf<uint16_t> dBlueNoise;
    DevBuf<float> dPmjPixel;
    DevBuf<int> dPrimes;
    DevBuf<uint16_t> dHaltonPerms;
    DevBuf<uint32_t> dHaltonPermOffset;
    int haltonPermSeed = 0;
    DevBuf<uint32_t> dSobolFull;
    DevBuf<uint64_t> dVdc, dVdcInv;
    std::vector<OwnedTexture> textures;
    DevBuf<float> dTexels, dEwaLUT;
    DevBuf<DeviceMIPLevel> dMipLevels;
    DevBuf<DeviceTexture> dTextures;
    DevBuf<float4> rayDiffs;
    bool haveTextures = false;
    DevBuf<DeviceLight> dLights;
    DevBuf<DeviceLightBVHNode> dLightNodes;
    DevBuf<int> dInfinite, dLightInBVH;
    DevBuf<uint32_t> dBitTrails;
    DevBuf<float> dTables;     // cieX, cieY, cieZ, illum (4 x 471)
    DevBuf<float> dFilter;     // func, cdf, funcInt, mfunc, mcdf, f
    DevBuf<double4> filmPixels;
    DevBuf<float> filmRGB;
    DevBuf<int> counts;
    DevBuf<unsigned long long> stats;
    DevBuf<int> queueMem;
    DevBuf<ShadowItem> shadowItems;
    DevBuf<float4> stateF4;
    DevBuf<uint4> stateMisc;
    DevBuf<float2> stateAux;
    DevBuf<int> rawCursor;
    DevBuf<DeviceImageLight> dImageLights;
    DevBuf<DeviceInstance> dInstances;
    DevBuf<float4> dTlasPrims;
    DevBuf<int> dInstFirstTri, dInstTriCount;
    DevBuf<int> hitInst;
    DevBuf<float> hitT;
    DevBuf<int> pathMedium;
    DevBuf<ShadowTrItem> shadowTr;
    DevBuf<DeviceMedium> dMedia;
    DevBuf<uint32_t> zsPrefix;
    long long zsKey[5] = {0, 0, 0, 0, 0};
    DevBuf<float4> visSurf;
    DevBuf<GBufferPixel> gbuf;
    DevBuf<float> gbufOut;
    std::vector<std::unique_ptr<DevBuf<float>>> mediumBufs;
    std::vector<std::unique_ptr<DevBuf<float4>>> mediumBufs4;
    bool haveInterface = false;
    bool raySort = false;            // BPT_RAY_SORT=1: sort the extension-ray queue by direction/origin key before each trace
    DevBuf<unsigned> sortKeys, sortKeysTmp, sortHist;
    DevBuf<int> sortQueueTmp;
    std::vector<WideBVH> blas;
    DevBuf<float> dRgb2spec;
    std::vector<std::unique_ptr<DevBuf<float>>> imgBufs;
    WideBVH bvh{};
    RenderCtx rc{};
    bool committed = false;
    bool materialPresent[MAT_NTAGS] = {};
    bool haveAreaLights = false;
    int nTrisTotal = 0;
    int maxPaths = 0;
    bpt_stats st{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool pending = false;            // an asynchronous render (bpt_render_stream + BPT_RENDER_ASYNC) has not been waited for yet
    int pendingLaunches = 0;
    int committedSpp = 0;
    int treeletRounds = -1;          // bpt_set_option("bvh_treelet_rounds"): < 0 = builder default

    int fail(int code, const std::string &m) { err = m; return code; }
    int cudaFail(cudaError_t e, const char *what) { err = std::string(what) + ": " + cudaGetErrorString(e); return BPT_ERR_CUDA; }
};

static std::string g_createError;

#define ECK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return e->cudaFail(e_, #call); } while (0)

// ------------------------------------------------------------------------------------------- spectra (host)
namespace {

// RGBToSpectrumTable::operator() (util/color.cpp:31-68)
void RGBToCoeffs(const Tables &t, const float rgbIn[3], float c[3]) {
    const std::vector<float> &zNodes = t.F("rgb2spec_srgb_scale");
    const std::vector<float> &coeffs = t.F("rgb2spec_srgb_coeffs");
    const int res = 64;
    float rgb[3] = {stdmax(0.f, rgbIn[0]), stdmax(0.f, rgbIn[1]), stdmax(0.f, rgbIn[2])};   // ClampZero
    if (rgb[0] == rgb[1] && rgb[1] == rgb[2]) {
        c[0] = 0; c[1] = 0; c[2] = (rgb[0] - .5f) / std::sqrt(rgb[0] * (1 - rgb[0]));
        return;
    }
    int maxc = (rgb[0] > rgb[1]) ? ((rgb[0] > rgb[2]) ? 0 : 2) : ((rgb[1] > rgb[2]) ? 1 : 2);
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * (res - 1) / z;
    float y = rgb[(maxc + 2) % 3] * (res - 1) / z;
    int xi = std::min((int)x, res - 2), yi = std::min((int)y, res - 2);
    int zi = FindIntervalLT(zNodes.data(), res, z);
    float dx = x - xi, dy = y - yi, dz = (z - zNodes[zi]) / (zNodes[zi + 1] - zNodes[zi]);
    for (int i = 0; i < 3; ++i) {
        auto co = [&](int ddx, int ddy, int ddz) {
            return coeffs[((((size_t)maxc * res + (zi + ddz)) * res + (yi + ddy)) * res + (xi + ddx)) * 3 + i];
        };
        c[i] = Lerp(dz, Lerp(dy, Lerp(dx, co(0, 0, 0), co(1, 0, 0)), Lerp(dx, co(0, 1, 0), co(1, 1, 0))),
                    Lerp(dy, Lerp(dx, co(0, 0, 1), co(1, 0, 1)), Lerp(dx, co(0, 1, 1), co(1, 1, 1))));
    }
}

enum SpecType { ST_ALBEDO, ST_UNBOUNDED, ST_ILLUMINANT };

struct SpecPool {
    std::vector<float> pool;
    int addDense(const std::vector<float> &v) { int o = (int)pool.size(); pool.insert(pool.end(), v.begin(), v.end()); return o; }
};

// host evaluation of a SpectrumRec at one wavelength (Spectrum::operator()(lambda))
float EvalRecAt(const SpectrumRec &r, const std::vector<float> &pool, float lambda) {
    switch (r.kind) {
    case SPEC_CONSTANT: return r.c0;
    case SPEC_RGB_ALBEDO: return SigmoidPoly(r.c0, r.c1, r.c2, lambda);
    case SPEC_RGB_UNBOUNDED: return r.scale * SigmoidPoly(r.c0, r.c1, r.c2, lambda);
    case SPEC_DENSE: return DenseLookup(pool.data() + r.table, lambda);
    case SPEC_PIECEWISE: return PiecewiseLinearEval(pool.data() + r.table, pool.data() + r.table + r.n, r.n, lambda);
    }
    return 0;
}

// Builds a SpectrumRec for a material / medium parameter.  Returns false if unsupported.
bool MakeSpectrumRec(const bpt_engine *e, const OwnedSpectrum &os, SpecType type, SpecPool &sp, SpectrumRec *out, std::string *err) {
    std::memset(out, 0, sizeof(*out));
    const bpt_spectrum &s = os.s;
    switch (s.kind) {
    case BPT_SPEC_CONSTANT: out->kind = SPEC_CONSTANT; out->c0 = s.value; return true;
    case BPT_SPEC_RGB: {
        float c[3];
        if (type == ST_ALBEDO) {
            RGBToCoeffs(e->tables, s.rgb, c);
            out->kind = SPEC_RGB_ALBEDO; out->c0 = c[0]; out->c1 = c[1]; out->c2 = c[2];
        } else {
            float m = stdmax(stdmax(s.rgb[0], s.rgb[1]), s.rgb[2]);
            float scale = 2 * m;
            float rgb[3] = {0, 0, 0};
            if (scale) { rgb[0] = s.rgb[0] / scale; rgb[1] = s.rgb[1] / scale; rgb[2] = s.rgb[2] / scale; }
            RGBToCoeffs(e->tables, rgb, c);
            out->kind = SPEC_RGB_UNBOUNDED; out->scale = scale; out->c0 = c[0]; out->c1 = c[1]; out->c2 = c[2];
            if (type == ST_ILLUMINANT) {
                // RGBIlluminantSpectrum: scale * rsp(lambda) * illuminant(lambda), densely sampled (LightBase::LookupSpectrum)
                const std::vector<float> &illum = e->tables.F("srgb_illuminant");
                std::vector<float> v(kNLambda);
                for (int i = 0; i < kNLambda; ++i) v[i] = scale * SigmoidPoly(c[0], c[1], c[2], (float)(kLambdaMin + i)) * illum[i];
                out->kind = SPEC_DENSE; out->table = sp.addDense(v);
            }
        }
        return true;
    }
    case BPT_SPEC_NAMED: {
        std::string key = std::string("pl:") + s.name;
        auto it = e->tables.f32.find(key);
        if (it == e->tables.f32.end()) {
            auto it2 = e->tables.f32.find(std::string("dense:") + s.name);
            if (it2 == e->tables.f32.end()) { *err = std::string("unknown named spectrum ") + s.name; return false; }
            out->kind = SPEC_DENSE; out->table = sp.addDense(it2->second);
            return true;
        }
        out->kind = SPEC_PIECEWISE; out->n = (int)it->second.size() / 2; out->table = sp.addDense(it->second);
        return true;
    }
    case BPT_SPEC_SAMPLED: {
        // PiecewiseLinearSpectrum::FromInterleaved without normalisation: lambdas then values
        int n = (int)os.samples.size() / 2;
        std::vector<float> v(2 * (size_t)n);
        for (int i = 0; i < n; ++i) { v[i] = os.samples[2 * i]; v[n + i] = os.samples[2 * i + 1]; }
        out->kind = SPEC_PIECEWISE; out->n = n; out->table = sp.addDense(v);
        return true;
    }
    case BPT_SPEC_BLACKBODY: {
        // BlackbodySpectrum (util/spectrum.h:69-82, 497-520), densely sampled the way LightBase::LookupSpectrum stores emission spectra;
        // elsewhere the reference evaluates it at the sampled wavelengths themselves
        if (type != ST_ILLUMINANT) { *err = "blackbody spectra are supported for emission only"; return false; }
        const float T = s.value;
        auto blackbody = [](float lambda, float T) -> float {
            if (T <= 0) return 0;
            const float c = 299792458.f, h = 6.62606957e-34f, kb = 1.3806488e-23f;
            float l = lambda * 1e-9f;
            float l2 = l * l * 1.f;
            float l5 = l2 * l2 * l;
            return (2 * h * c * c) / (l5 * (FastExp((h * c) / (l * kb * T)) - 1));
        };
        float lambdaMax = 2.8977721e-3f / T;
        float normalizationFactor = 1 / blackbody(lambdaMax * 1e9f, T);
        std::vector<float> v(kNLambda);
        for (int i = 0; i < kNLambda; ++i) v[i] = blackbody((float)(kLambdaMin + i), T) * normalizationFactor;
        out->kind = SPEC_DENSE; out->table = sp.addDense(v);
        return true;
    }
    default: *err = "unsupported spectrum kind"; return false;
    }
}

// FilterSampler tables (filters.cpp:FilterSampler ctor) + PiecewiseConstant2D (util/sampling.h)
struct FilterTables {
    int nx = 0, ny = 0;
    std::vector<float> f, func, cdf, funcInt, mfunc, mcdf;
    float mfuncInt = 0;
};
void BuildPC1D(const float *f, int n, float mn, float mx, float *func, float *cdf, float *funcInt) {
    for (int i = 0; i < n; ++i) func[i] = std::abs(f[i]);
    cdf[0] = 0;
    for (int i = 1; i < n + 1; ++i) cdf[i] = cdf[i - 1] + func[i - 1] * (mx - mn) / n;
    *funcInt = cdf[n];
    if (*funcInt == 0) for (int i = 1; i < n + 1; ++i) cdf[i] = float(i) / float(n);
    else for (int i = 1; i < n + 1; ++i) cdf[i] /= *funcInt;
}
void BuildPC2D(const std::vector<float> &f, int nu, int nv, float min0, float max0, float min1, float max1, FilterTables *t) {
    t->func.resize((size_t)nu * nv); t->cdf.resize((size_t)(nu + 1) * nv); t->funcInt.resize(nv);
    for (int v = 0; v < nv; ++v)
        BuildPC1D(f.data() + (size_t)v * nu, nu, min0, max0, t->func.data() + (size_t)v * nu, t->cdf.data() + (size_t)v * (nu + 1), &t->funcInt[v]);
    t->mfunc.resize(nv); t->mcdf.resize(nv + 1);
    std::vector<float> marg(t->funcInt);
    BuildPC1D(marg.data(), nv, min1, max1, t->mfunc.data(), t->mcdf.data(), &t->mfuncInt);
}

std::string DefaultTablesPath() {
    Dl_info info;
    if (dladdr((void *)&DefaultTablesPath, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t s = p.find_last_of('/');
        std::string dir = (s == std::string::npos) ? "." : p.substr(0, s);
        return dir + "/data/pbrt_tables.bin";
    }
    return "pbrt_tables.bin";
}

}  // namespace

// ------------------------------------------------------------------------------------------- lifecycle
extern "C" int bpt_create(int cuda_device, const char *tables_path, bpt_engine **out) {
    if (!out) return BPT_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        g_createError = std::string("no CUDA device: ") + (ce != cudaSuccess ? cudaGetErrorString(ce) : "device count 0") +
                        " (this engine has no CPU fallback)";
        return BPT_ERR_NO_DEVICE;
    }
    if (cuda_device < 0 || cuda_device >= ndev) { g_createError = "invalid device index"; return BPT_ERR_INVALID; }
    std::unique_ptr<bpt_engine> e(new bpt_engine);
    e->device = cuda_device;
    if ((ce = cudaSetDevice(cuda_device)) != cudaSuccess) { g_createError = cudaGetErrorString(ce); return BPT_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cuda_device);
    e->smCount = prop.multiProcessorCount;
    if (prop.major < 10) { g_createError = "device is not sm_100 class"; return BPT_ERR_NO_DEVICE; }
    std::string terr;
    if (!e->tables.Load(tables_path ? tables_path : DefaultTablesPath(), &terr)) { g_createError = terr; return BPT_ERR_IO; }
    if ((ce = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess) { g_createError = cudaGetErrorString(ce); return BPT_ERR_CUDA; }
    cudaEventCreate(&e->ev0); cudaEventCreate(&e->ev1);
    {   // keep freed stream-ordered allocations (BVH builder scratch / outputs) in the device pool across re-commits
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, e->device) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
    }
    UploadSobolTable(e->tables.u32.at("sobol32_2x52").data());
    UploadSobolTableTex(e->tables.u32.at("sobol32_2x52").data());
    tu_shapes::UploadSobolTable(e->tables.u32.at("sobol32_2x52").data());
    ConfigureKernels();
    tu_shapes::ConfigureKernels();
    // defaults (scene.cpp / film.cpp / samplers.cpp defaults)
    e->filter = bpt_filter{BPT_FILTER_GAUSSIAN, 1.5f, 1.5f, 0.5f, 1.f / 3.f, 1.f / 3.f, 3.f};
    e->sampler = bpt_sampler{BPT_SAMPLER_ZSOBOL, 16, 0, BPT_RANDOMIZE_FAST_OWEN};
    e->integ = bpt_integrator{5, 0, BPT_LIGHTSAMPLER_BVH};
    const char *mp = std::getenv("BPT_MAX_PATHS");
    e->maxPaths = mp ? std::atoi(mp) : (32 << 20);   // profiles/r01g: 4M -> 32M paths per wavefront pass is -20 % step time (launch tails)
    *out = e.release();
    return BPT_OK;
}

extern "C" void bpt_destroy(bpt_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    FreeWideBVH(&e->bvh);
    for (WideBVH &b : e->blas) FreeWideBVH(&b);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" const char *bpt_last_error(const bpt_engine *e) { return e ? e->err.c_str() : g_createError.c_str(); }

// ------------------------------------------------------------------------------------------- scene intake
extern "C" int bpt_scene_begin(bpt_engine *e) {
    if (!e) return BPT_ERR_INVALID;
    e->materials.clear(); e->areaLights.clear(); e->meshes.clear(); e->instances.clear(); e->lights.clear(); e->media.clear();
    e->haveCam = e->haveFilm = false; e->committed = false; e->cameraMedium = -1; e->space = BPT_SPACE_WORLD;
    return BPT_OK;
}
extern "C" int bpt_scene_set_camera(bpt_engine *e, const bpt_camera *c) {
    if (!e || !c) return BPT_ERR_INVALID;
    if (c->type != BPT_CAMERA_PERSPECTIVE && c->type != BPT_CAMERA_ORTHOGRAPHIC) return e->fail(BPT_ERR_UNSUPPORTED, "camera not implemented (perspective, orthographic)");
    e->cam = *c; e->haveCam = true; return BPT_OK;
}
extern "C" int bpt_scene_set_film(bpt_engine *e, const bpt_film *f) {
    if (!e || !f) return BPT_ERR_INVALID;
    if (f->type != BPT_FILM_RGB && f->type != BPT_FILM_GBUFFER) return e->fail(BPT_ERR_UNSUPPORTED, "film type not implemented");
    if (f->xres <= 0 || f->yres <= 0 || f->xres > 65535 || f->yres > 65535) return e->fail(BPT_ERR_INVALID, "bad film resolution");
    e->film = *f; e->haveFilm = true; return BPT_OK;
}
extern "C" int bpt_scene_set_filter(bpt_engine *e, const bpt_filter *f) {
    if (!e || !f) return BPT_ERR_INVALID;
    if (f->type < BPT_FILTER_BOX || f->type > BPT_FILTER_TRIANGLE) return e->fail(BPT_ERR_INVALID, "unknown filter type");
    if (!(f->xradius > 0) || !(f->yradius > 0)) return e->fail(BPT_ERR_INVALID, "filter radius must be positive");
    e->filter = *f; return BPT_OK;
}
extern "C" int bpt_scene_set_sampler(bpt_engine *e, const bpt_sampler *s) {
    if (!e || !s) return BPT_ERR_INVALID;
    if (s->type != BPT_SAMPLER_ZSOBOL && s->type != BPT_SAMPLER_INDEPENDENT && s->type != BPT_SAMPLER_STRATIFIED && s->type != BPT_SAMPLER_PADDED_SOBOL &&
        s->type != BPT_SAMPLER_SOBOL && s->type != BPT_SAMPLER_HALTON && s->type != BPT_SAMPLER_PMJ02BN)
        return e->fail(BPT_ERR_UNSUPPORTED, "sampler not implemented (zsobol, sobol, paddedsobol, halton, pmj02bn, stratified, independent)");
    if (s->type == BPT_SAMPLER_HALTON && s->randomize == BPT_RANDOMIZE_FAST_OWEN) return e->fail(BPT_ERR_INVALID, "\"fastowen\" randomization not supported by Halton sampler");
    e->sampler = *s;
    if (s->type == BPT_SAMPLER_STRATIFIED) {
        if (s->x_samples <= 0 || s->y_samples <= 0) return e->fail(BPT_ERR_INVALID, "stratified sampler: xsamples and ysamples must be positive");
        e->sampler.spp = s->x_samples * s->y_samples;   // StratifiedSampler::SamplesPerPixel()
    }
    if (e->sampler.spp <= 0) return e->fail(BPT_ERR_INVALID, "spp must be positive");
    return BPT_OK;
}
extern "C" int bpt_scene_set_integrator(bpt_engine *e, const bpt_integrator *i) {
    if (!e || !i) return BPT_ERR_INVALID;
    if (i->light_sampler != BPT_LIGHTSAMPLER_BVH && i->light_sampler != BPT_LIGHTSAMPLER_UNIFORM)
        return e->fail(BPT_ERR_UNSUPPORTED, "light sampler not implemented (bvh, uniform)");
    if (i->max_depth < 0 || i->max_depth > 200) return e->fail(BPT_ERR_INVALID, "bad maxdepth");
    e->integ = *i; return BPT_OK;
}
extern "C" int bpt_scene_set_materials(bpt_engine *e, const bpt_material *m, int n) {
    if (!e || (n > 0 && !m)) return BPT_ERR_INVALID;
    e->materials.resize(n);
    for (int i = 0; i < n; ++i) {
        OwnedMaterial &o = e->materials[i];
        o.m = m[i];
        o.reflectance.Set(m[i].reflectance); o.eta.Set(m[i].eta); o.k.Set(m[i].k); o.iface_eta.Set(m[i].iface_eta);
        o.conductor_eta.Set(m[i].conductor_eta); o.conductor_k.Set(m[i].conductor_k); o.albedo.Set(m[i].albedo); o.transmittance.Set(m[i].transmittance);
    }
    return BPT_OK;
}
extern "C" int bpt_scene_set_textures(bpt_engine *e, const bpt_texture *t, int n) {
    if (!e || (n > 0 && !t)) return BPT_ERR_INVALID;
    e->textures.clear();
    e->textures.resize(n);
    for (int i = 0; i < n; ++i) {
        const bpt_texture &b = t[i];
        if (b.kind == BPT_TEXTURE_CONSTANT || b.kind == BPT_TEXTURE_SCALE || b.kind == BPT_TEXTURE_MIX || b.kind == BPT_TEXTURE_CHECKERBOARD) {
            // image-less classes: children are earlier entries
            const int nch = b.kind == BPT_TEXTURE_CONSTANT ? 0 : b.kind == BPT_TEXTURE_MIX ? 3 : 2;
            const int ch[3] = {b.tex1, b.tex2, b.tex3};
            for (int k = 0; k < nch; ++k) if (ch[k] < 1 || ch[k] > i) return e->fail(BPT_ERR_INVALID, "texture refers to a texture that is not defined before it");
            OwnedTexture &o = e->textures[i];
            o.t = b; o.t.pixels = nullptr;
            o.value.Set(b.value_s);
            continue;
        }
        if (b.kind != BPT_TEXTURE_IMAGEMAP && b.kind != BPT_TEXTURE_NORMALMAP) return e->fail(BPT_ERR_UNSUPPORTED, "texture class not implemented (imagemap, constant, scale, mix, checkerboard)");
        if (!b.pixels || b.xres <= 0 || b.yres <= 0 || (b.n_channels != 1 && b.n_channels != 3 && b.n_channels != 4))
            return e->fail(BPT_ERR_INVALID, "texture image: 1, 3 or 4 channels of 8-bit or float texels expected");
        if (b.format != BPT_IMAGE_U8 && b.format != BPT_IMAGE_F32) return e->fail(BPT_ERR_UNSUPPORTED, "texture image format not implemented (8-bit, float)");
        if (b.filter < BPT_TEXFILTER_POINT || b.filter > BPT_TEXFILTER_EWA || b.wrap < BPT_TEXWRAP_BLACK || b.wrap > BPT_TEXWRAP_REPEAT)
            return e->fail(BPT_ERR_INVALID, "texture filter / wrap mode out of range");
        OwnedTexture &o = e->textures[i];
        o.t = b; o.t.pixels = nullptr;
        HostImage &im = o.img;
        im.xres = b.xres; im.yres = b.yres; im.nc = b.n_channels;
        const size_t cnt = (size_t)b.xres * b.yres * b.n_channels;
        if (b.format == BPT_IMAGE_U8) { im.is8 = true; im.srgb = b.encoding == BPT_ENCODING_SRGB; im.p8.assign((const uint8_t *)b.pixels, (const uint8_t *)b.pixels + cnt); }
        else { im.is8 = false; im.p32.assign((const float *)b.pixels, (const float *)b.pixels + cnt); }
    }
    return BPT_OK;
}
extern "C" int bpt_scene_set_area_lights(bpt_engine *e, const bpt_area_light *a, int n) {
    if (!e || (n > 0 && !a)) return BPT_ERR_INVALID;
    e->areaLights.resize(n);
    for (int i = 0; i < n; ++i) { e->areaLights[i].a = a[i]; e->areaLights[i].L.Set(a[i].L); }
    return BPT_OK;
}
extern "C" int bpt_scene_add_mesh(bpt_engine *e, const bpt_mesh *m) {
    if (!e || !m || !m->P || !m->indices || m->n_vertices <= 0 || m->n_triangles < 0) return e ? e->fail(BPT_ERR_INVALID, "bad mesh") : BPT_ERR_INVALID;
    for (int i = 0; i < 3 * m->n_triangles; ++i)
        if (m->indices[i] < 0 || m->indices[i] >= m->n_vertices) return e->fail(BPT_ERR_INVALID, "mesh index out of range");
    e->meshes.emplace_back();
    OwnedMesh &o = e->meshes.back();
    o.m = *m;
    o.P.assign(m->P, m->P + 3 * (size_t)m->n_vertices);
    if (m->N) o.N.assign(m->N, m->N + 3 * (size_t)m->n_vertices);
    if (m->UV) o.UV.assign(m->UV, m->UV + 2 * (size_t)m->n_vertices);
    o.idx.assign(m->indices, m->indices + 3 * (size_t)m->n_triangles);
    o.m.P = o.m.N = o.m.UV = nullptr; o.m.indices = nullptr;
    return BPT_OK;
}
extern "C" int bpt_scene_add_bilinear_mesh(bpt_engine *e, const bpt_mesh *m) {
    if (!e || !m || !m->P || !m->indices || m->n_vertices <= 0 || m->n_triangles < 0) return e ? e->fail(BPT_ERR_INVALID, "bad bilinear mesh") : BPT_ERR_INVALID;
    for (int i = 0; i < 4 * m->n_triangles; ++i)
        if (m->indices[i] < 0 || m->indices[i] >= m->n_vertices) return e->fail(BPT_ERR_INVALID, "bilinear mesh index out of range");
    if (m->has_alpha && m->alpha < 1) return e->fail(BPT_ERR_UNSUPPORTED, "alpha on bilinear patches is not implemented");
    e->meshes.emplace_back();
    OwnedMesh &o = e->meshes.back();
    o.m = *m; o.isBilinear = true;
    o.P.assign(m->P, m->P + 3 * (size_t)m->n_vertices);
    if (m->N) o.N.assign(m->N, m->N + 3 * (size_t)m->n_vertices);
    if (m->UV) o.UV.assign(m->UV, m->UV + 2 * (size_t)m->n_vertices);
    o.idx.assign(m->indices, m->indices + 4 * (size_t)m->n_triangles);
    o.m.P = o.m.N = o.m.UV = nullptr; o.m.indices = nullptr;
    return BPT_OK;
}
extern "C" int bpt_scene_add_quadric(bpt_engine *e, const bpt_quadric *q) {
    if (!e || !q) return e ? e->fail(BPT_ERR_INVALID, "bad quadric") : BPT_ERR_INVALID;
    if (q->type != BPT_SHAPE_SPHERE && q->type != BPT_SHAPE_DISK && q->type != BPT_SHAPE_CYLINDER) return e->fail(BPT_ERR_UNSUPPORTED, "shape type not implemented (trianglemesh, sphere, disk, cylinder)");
    if (!(q->radius > 0)) return e->fail(BPT_ERR_INVALID, "quadric radius must be positive");
    e->meshes.emplace_back();
    OwnedMesh &o = e->meshes.back();
    o.isQuadric = true; o.q = *q;
    o.m = bpt_mesh{};
    o.m.n_vertices = 0; o.m.n_triangles = 1;
    o.m.world_from_object = q->world_from_object; o.m.object_from_world = q->object_from_world;
    o.m.reverse_orientation = q->reverse_orientation; o.m.material = q->material; o.m.area_light = q->area_light;
    o.m.medium_inside = q->medium_inside; o.m.medium_outside = q->medium_outside; o.m.object_id = q->object_id;
    o.m.has_alpha = 0; o.m.alpha = 1.f;
    return BPT_OK;
}
// In-library fast path of the .pbrt front end (csrc/pbrt_frontend.cpp lives in the same shared object): the arrays it has just read
// from a PLY file are moved into the engine instead of copied through the public ABI (2 x 19 MB of memcpy per re-commit of the 1 M-
// triangle bench scene).  Same validation as bpt_scene_add_mesh.
int bpt_internal_add_mesh_move(bpt_engine *e, const bpt_mesh *m, std::vector<float> &&P, std::vector<float> &&N, std::vector<float> &&UV, std::vector<int> &&idx) {
    if (!e || !m || P.empty() || idx.empty() || m->n_vertices <= 0 || m->n_triangles < 0) return e ? e->fail(BPT_ERR_INVALID, "bad mesh") : BPT_ERR_INVALID;
    if (P.size() != 3 * (size_t)m->n_vertices || idx.size() != 3 * (size_t)m->n_triangles) return e->fail(BPT_ERR_INVALID, "bad mesh");
    unsigned bad = 0;
    const unsigned nv = (unsigned)m->n_vertices;
    for (int v : idx) bad |= (unsigned)v >= nv;
    if (bad) return e->fail(BPT_ERR_INVALID, "mesh index out of range");
    e->meshes.emplace_back();
    OwnedMesh &o = e->meshes.back();
    o.m = *m;
    o.P = std::move(P); o.N = std::move(N); o.UV = std::move(UV); o.idx = std::move(idx);
    o.m.P = o.m.N = o.m.UV = nullptr; o.m.indices = nullptr;
    return BPT_OK;
}
extern "C" int bpt_scene_add_instance(bpt_engine *e, const bpt_instance *i) {
    if (!e || !i) return BPT_ERR_INVALID;
    e->instances.push_back(*i); return BPT_OK;
}
extern "C" int bpt_scene_add_light(bpt_engine *e, const bpt_light *l) {
    if (!e || !l) return BPT_ERR_INVALID;
    if (l->type != BPT_LIGHT_UNIFORM_INFINITE && l->type != BPT_LIGHT_IMAGE_INFINITE && l->type != BPT_LIGHT_POINT &&
        l->type != BPT_LIGHT_SPOT && l->type != BPT_LIGHT_DISTANT)
        return e->fail(BPT_ERR_UNSUPPORTED, "light type not implemented (infinite, point, spot, distant; area lights through bpt_scene_set_area_lights)");
    if (l->type == BPT_LIGHT_SPOT && l->cone_delta < 0) return e->fail(BPT_ERR_INVALID, "spot light: falloff start beyond the cone angle");
    e->lights.emplace_back();
    OwnedLight &o = e->lights.back();
    o.l = *l; o.L.Set(l->L);
    if (l->image_rgb && l->image_res > 0) o.image.assign(l->image_rgb, l->image_rgb + 3 * (size_t)l->image_res * l->image_res);
    o.l.image_rgb = nullptr;
    return BPT_OK;
}
extern "C" int bpt_scene_add_medium(bpt_engine *e, const bpt_medium *m) {
    if (!e || !m) return BPT_ERR_INVALID;
    if (m->type != BPT_MEDIUM_UNIFORMGRID && m->type != BPT_MEDIUM_HOMOGENEOUS && m->type != BPT_MEDIUM_RGBGRID)
        return e->fail(BPT_ERR_UNSUPPORTED, "medium type not implemented (homogeneous, uniformgrid, rgbgrid)");
    if (m->type == BPT_MEDIUM_RGBGRID) {
        if ((!m->sigma_a_rgb && !m->sigma_s_rgb) || m->nx <= 0 || m->ny <= 0 || m->nz <= 0) return e->fail(BPT_ERR_INVALID, "RGB grid requires \"sigma_a\" and/or \"sigma_s\" parameter values");
        if (m->Le_rgb && !m->sigma_a_rgb) return e->fail(BPT_ERR_INVALID, "RGB grid requires \"sigma_a\" if \"Le\" value provided");
    }
    if (m->type == BPT_MEDIUM_UNIFORMGRID && (!m->density || m->nx <= 0 || m->ny <= 0 || m->nz <= 0)) return e->fail(BPT_ERR_INVALID, "uniformgrid medium needs a density grid");
    e->media.emplace_back();
    OwnedMedium &o = e->media.back();
    o.m = *m; o.sigma_a.Set(m->sigma_a); o.sigma_s.Set(m->sigma_s); o.Le.Set(m->Le);
    if (m->type == BPT_MEDIUM_UNIFORMGRID) o.density.assign(m->density, m->density + (size_t)m->nx * m->ny * m->nz);
    if (m->type == BPT_MEDIUM_UNIFORMGRID && m->Le_scale_grid) o.leScale.assign(m->Le_scale_grid, m->Le_scale_grid + (size_t)m->nx * m->ny * m->nz);
    if (m->type == BPT_MEDIUM_RGBGRID) {
        const size_t n3 = 3 * (size_t)m->nx * m->ny * m->nz;
        if (m->sigma_a_rgb) o.rgbA.assign(m->sigma_a_rgb, m->sigma_a_rgb + n3);
        if (m->sigma_s_rgb) o.rgbS.assign(m->sigma_s_rgb, m->sigma_s_rgb + n3);
        if (m->Le_rgb) o.rgbLe.assign(m->Le_rgb, m->Le_rgb + n3);
    }
    o.m.density = nullptr; o.m.Le_scale_grid = nullptr; o.m.sigma_a_rgb = o.m.sigma_s_rgb = o.m.Le_rgb = nullptr;
    return BPT_OK;
}
// Engine options that are not part of the scene description (they survive bpt_scene_begin).
extern "C" int bpt_set_option(bpt_engine *e, const char *name, long long value) {
    if (!e || !name) return BPT_ERR_INVALID;
    std::string n = name;
    if (n == "bvh_treelet_rounds") { e->treeletRounds = (int)value; return BPT_OK; }
    if (n == "max_paths") { if (value < 1024) return e->fail(BPT_ERR_INVALID, "max_paths too small"); e->maxPaths = (int)std::min<long long>(value, 1ll << 30); return BPT_OK; }
    return e->fail(BPT_ERR_INVALID, "unknown option " + n);
}
extern "C" int bpt_scene_set_transform_space(bpt_engine *e, int space) {
    if (!e || (space != BPT_SPACE_WORLD && space != BPT_SPACE_RENDER)) return BPT_ERR_INVALID;
    e->space = space;
    return BPT_OK;
}
extern "C" int bpt_scene_set_camera_medium(bpt_engine *e, int medium) { if (!e) return BPT_ERR_INVALID; e->cameraMedium = medium; return BPT_OK; }

// ------------------------------------------------------------------------------------------- commit
static Mat4 ToMat4(const M4 &a) { Mat4 r; std::memcpy(r.m, a.m, sizeof(r.m)); return r; }

// TextureMapping2D::Create (textures.cpp:49-73) for the mappings that look at the shading point: textureFromRender = Inverse(renderFromTexture)
static void SetTextureMapping(DeviceTexture *d, const bpt_texture &b, int space, const CameraSpaces &spaces) {
    d->mapping = b.mapping;
    if (b.mapping == BPT_TEXMAP_UV) return;
    Xform renderFromTexture = ResolveRenderFromObject(space, spaces, b.world_from_texture, b.texture_from_world);
    d->textureFromRender = ToMat4(renderFromTexture.mInv);
    d->vs = Vec3(b.v1[0], b.v1[1], b.v1[2]); d->vt = Vec3(b.v2[0], b.v2[1], b.v2[2]);
}
extern "C" int bpt_scene_commit(bpt_engine *e) {
    const bool timing = std::getenv("BPT_TIMING") != nullptr;
    auto tc0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
        if (!timing) return;
        cudaDeviceSynchronize();
        auto t = std::chrono::steady_clock::now();
        fprintf(stderr, "[bpt commit] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t - tc0).count());
        tc0 = t;
    };

    if (!e) return BPT_ERR_INVALID;
    if (!e->haveCam || !e->haveFilm) return e->fail(BPT_ERR_INVALID, "camera and film must be set before commit");
    ECK(cudaSetDevice(e->device));
    auto t0 = std::chrono::steady_clock::now();
    e->committed = false;
    std::string serr;

    // ---- camera transforms (BasicSceneBuilder::Camera scene.cpp:138-156, CameraTransform cameras.cpp, rendering space = CameraWorld)
    CameraSpaces spaces;
    if (!ResolveCameraSpaces(e->cam, &spaces)) return e->fail(BPT_ERR_INVALID, "singular camera transform");
    const int space = e->space;
    const Xform &worldFromRender = spaces.worldFromRender, &renderFromWorld = spaces.renderFromWorld, &renderFromCamera = spaces.renderFromCamera;
    (void)renderFromWorld;

    // ---- film / filter / sensor (film.cpp FilmBaseParameters, PixelSensor::Create, RGBFilm ctor)
    const bpt_film &fl = e->film;
    int pb[4] = {fl.pixel_bounds[0], fl.pixel_bounds[1], fl.pixel_bounds[2], fl.pixel_bounds[3]};
    if (pb[0] == 0 && pb[1] == 0 && pb[2] == 0 && pb[3] == 0) { pb[0] = 0; pb[1] = fl.xres; pb[2] = 0; pb[3] = fl.yres; }
    if (pb[0] < 0 || pb[1] > fl.xres || pb[2] < 0 || pb[3] > fl.yres || pb[0] >= pb[1] || pb[2] >= pb[3]) return e->fail(BPT_ERR_INVALID, "bad pixel bounds");
    DeviceFilm df{};
    df.xres = fl.xres; df.yres = fl.yres; df.pbx0 = pb[0]; df.pbx1 = pb[1]; df.pby0 = pb[2]; df.pby1 = pb[3];
    float exposureTime = fl.exposure_time > 0 ? fl.exposure_time : (e->cam.shutter_close - e->cam.shutter_open);
    if (!(exposureTime > 0)) exposureTime = 1;
    float iso = fl.iso > 0 ? fl.iso : 100.f;
    df.imagingRatio = exposureTime * iso / 100;
    df.maxComponentValue = fl.max_component_value > 0 ? fl.max_component_value : BPT_INFINITY;
    if (fl.white_balance != 0) return e->fail(BPT_ERR_UNSUPPORTED, "sensor white balance is not implemented");
    {
        const std::vector<float> &m = e->tables.F("srgb_rgb_from_xyz");
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) df.outRGBFromSensor[i][j] = m[3 * i + j];   // RGBFromXYZ * identity
    }
    df.isGBuffer = fl.type == BPT_FILM_GBUFFER;
    {
        const std::vector<float> &m = e->tables.F("srgb_rgb_from_xyz");
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) df.rgbFromXYZ[i][j] = m[3 * i + j];
        // outputFromRender: "camera" -> cameraTransform.RenderFromCamera() applied inverted, "world" -> WorldFromRender forward
        Xform out = fl.coordinate_system == 1 ? worldFromRender : renderFromCamera;
        df.applyInverse = fl.coordinate_system == 1 ? 0 : 1;
        df.outM = ToMat4(out.m); df.outMInv = ToMat4(out.mInv);
    }

    DeviceFilter dflt{};
    FilterTables ft;
    dflt.radiusX = e->filter.xradius; dflt.radiusY = e->filter.yradius;
    // FilterSampler (filters.cpp:96-118) for the filters the reference samples from a table: Gaussian, Mitchell, LanczosSinc --
    // the unnormalised filter tabulated at 32 samples per unit radius, PiecewiseConstant2D over |f| (sampling.h), weight f / pdf
    // (signed for the negative lobes of Mitchell and sinc).  Box and Triangle are sampled analytically on the device (filters.h).
    {
        const int type = e->filter.type;
        const float rx = e->filter.xradius, ry = e->filter.yradius;
        // Mitchell1D (filters.h:147-160), WindowedSinc / Sinc / SinXOverX (util/math.h:381-398)
        const float B = e->filter.b, Cc = e->filter.c, tau = e->filter.tau;
        auto mitchell1D = [&](float x) {
            x = std::abs(x);
            if (x <= 1) return ((12 - 9 * B - 6 * Cc) * x * x * x + (-18 + 12 * B + 6 * Cc) * x * x + (6 - 2 * B)) * (1.f / 6.f);
            else if (x <= 2) return ((-B - 6 * Cc) * x * x * x + (6 * B + 30 * Cc) * x * x + (-12 * B - 48 * Cc) * x + (8 * B + 24 * Cc)) * (1.f / 6.f);
            else return 0.f;
        };
        auto sinXOverX = [](float x) { if (1 - x * x == 1) return 1.f; return std::sin(x) / x; };
        auto sinc = [&](float x) { return sinXOverX(kPi * x); };
        auto windowedSinc = [&](float x, float radius) { if (std::abs(x) > radius) return 0.f; return sinc(x) * sinc(x / tau); };
        const float sigma = e->filter.sigma;
        const float expX = type == BPT_FILTER_GAUSSIAN ? Gaussian(rx, 0, sigma) : 0.f, expY = type == BPT_FILTER_GAUSSIAN ? Gaussian(ry, 0, sigma) : 0.f;
        auto evaluate = [&](float px, float py) -> float {
            switch (type) {
            case BPT_FILTER_GAUSSIAN: return stdmax(0.f, Gaussian(px, 0, sigma) - expX) * stdmax(0.f, Gaussian(py, 0, sigma) - expY);
            case BPT_FILTER_MITCHELL: return mitchell1D(2 * px / rx) * mitchell1D(2 * py / ry);
            default: return windowedSinc(px, rx) * windowedSinc(py, ry);
            }
        };
        if (type == BPT_FILTER_GAUSSIAN || type == BPT_FILTER_MITCHELL || type == BPT_FILTER_SINC) {
            ft.nx = int(32 * rx); ft.ny = int(32 * ry);
            if (ft.nx < 1 || ft.ny < 1) return e->fail(BPT_ERR_INVALID, "filter radius too small for the sampling table");
            ft.f.resize((size_t)ft.nx * ft.ny);
            for (int y = 0; y < ft.ny; ++y)
                for (int x = 0; x < ft.nx; ++x) {
                    float px = Lerp((x + 0.5f) / ft.nx, -rx, rx), py = Lerp((y + 0.5f) / ft.ny, -ry, ry);
                    ft.f[(size_t)y * ft.nx + x] = evaluate(px, py);
                }
            BuildPC2D(ft.f, ft.nx, ft.ny, -rx, rx, -ry, ry, &ft);
            dflt.sampled = 1; dflt.nx = ft.nx; dflt.ny = ft.ny;
        } else dflt.sampled = type == BPT_FILTER_TRIANGLE ? 2 : 0;
    }

    // ---- projective camera (cameras.h:254-273, PerspectiveCamera::Create cameras.cpp)
    DeviceCamera dc{};
    {
        float frame = e->cam.frame_aspect_ratio > 0 ? e->cam.frame_aspect_ratio : float(fl.xres) / float(fl.yres);
        float sw[4];
        if (frame > 1.f) { sw[0] = -frame; sw[1] = frame; sw[2] = -1.f; sw[3] = 1.f; }
        else { sw[0] = -1.f; sw[1] = 1.f; sw[2] = -1.f / frame; sw[3] = 1.f / frame; }
        if (e->cam.has_screen_window) for (int i = 0; i < 4; ++i) sw[i] = e->cam.screen_window[i];
        // OrthographicCamera: Orthographic(0, 1) = Scale(1, 1, 1 / (zFar - zNear)) * Translate(0, 0, -zNear) (cameras.h:288, util/transform.cpp:115-117)
        const bool ortho = e->cam.type == BPT_CAMERA_ORTHOGRAPHIC;
        Xform screenFromCamera = ortho ? Scale(1, 1, 1 / (1.f - 0.f)) * Translate(Vec3(0, 0, -0.f)) : Perspective(e->cam.fov > 0 ? e->cam.fov : 90.f, 1e-2f, 1000.f);
        Xform NDCFromScreen = Scale(1 / (sw[1] - sw[0]), 1 / (sw[3] - sw[2]), 1) * Translate(Vec3(-sw[0], -sw[3], 0));
        Xform rasterFromNDC = Scale((float)fl.xres, -(float)fl.yres, 1);
        Xform rasterFromScreen = rasterFromNDC * NDCFromScreen;
        Xform screenFromRaster = Inverse(rasterFromScreen);
        Xform cameraFromRaster = Inverse(screenFromCamera) * screenFromRaster;
        dc.cameraFromRaster = ToMat4(cameraFromRaster.m);
        dc.renderFromCamera = ToMat4(renderFromCamera.m);
        dc.lensRadius = e->cam.lens_radius;
        dc.focalDistance = e->cam.focal_distance > 0 ? e->cam.focal_distance : 1e6f;
        dc.type = e->cam.type;
        // dxCamera = cameraFromRaster(Point3f(1,0,0)) - cameraFromRaster(Point3f(0,0,0))  (cameras.h:347-351)
        Pt3 c0 = cameraFromRaster(Pt3(0, 0, 0)), cx = cameraFromRaster(Pt3(1, 0, 0)), cy = cameraFromRaster(Pt3(0, 1, 0));
        dc.dxCamera = cx - c0; dc.dyCamera = cy - c0;
        if (ortho) { dc.dxCamera = cameraFromRaster(Vec3(1, 0, 0)); dc.dyCamera = cameraFromRaster(Vec3(0, 1, 0)); }   // cameras.h:291-292
        dc.rayDiffScale = std::max(.125f, 1 / std::sqrt((float)std::max(1, e->sampler.spp)));
        if (!e->textures.empty()) {   // Camera::Approximate_dp_dxy needs these; only texture filtering reads them
            dc.cameraFromRender = ToMat4(renderFromCamera.mInv);
            dc.sppScale = std::max(.125f, 1 / std::sqrt((float)std::max(1, e->sampler.spp)));
            FindMinimumDifferentials(&dc, fl.xres, fl.yres);
        }
    }

    // ---- sampler (samplers.h:228-239)
    SamplerParams sp{};
    sp.type = e->sampler.type == BPT_SAMPLER_ZSOBOL ? 0 : e->sampler.type == BPT_SAMPLER_STRATIFIED ? 2 : e->sampler.type == BPT_SAMPLER_PADDED_SOBOL ? 3 :
              e->sampler.type == BPT_SAMPLER_SOBOL ? 4 : e->sampler.type == BPT_SAMPLER_HALTON ? 5 : e->sampler.type == BPT_SAMPLER_PMJ02BN ? 6 : 1;
    sp.xSamples = e->sampler.x_samples; sp.ySamples = e->sampler.y_samples; sp.jitter = e->sampler.jitter;
    sp.randomize = e->sampler.randomize;
    sp.seed = e->sampler.seed;
    sp.samplesPerPixel = e->sampler.spp;
    if (sp.type == 6) {   // PMJ02BNSampler::PMJ02BNSampler (samplers.cpp:173-207)
        if (!e->tables.u32.count("pmj02bn_samples") || !e->tables.u32.count("blue_noise_u16x2"))
            return e->fail(BPT_ERR_IO, "tables blob lacks the pmj02bn sampler tables (regenerate it with oracle/_ref/bin/ref_dump tables)");
        const std::vector<uint32_t> &pm = e->tables.u32.at("pmj02bn_samples"), &bn = e->tables.u32.at("blue_noise_u16x2");
        std::vector<float> pix;
        if (!BuildPmjPixelSamples(pm.data(), sp.samplesPerPixel, &sp.pixelTileSize, &pix)) return e->fail(BPT_ERR_INVALID, "PMJ02BNSampler only supports up to 65536 samples per pixel");
        std::vector<uint16_t> bn16(2 * bn.size());
        for (size_t i = 0; i < bn.size(); ++i) { bn16[2 * i] = (uint16_t)(bn[i] & 0xffffu); bn16[2 * i + 1] = (uint16_t)(bn[i] >> 16); }
        ECK(e->dPmjSamples.upload(pm)); ECK(e->dBlueNoise.upload(bn16)); ECK(e->dPmjPixel.upload(pix));
        sp.pmjSamples = e->dPmjSamples.p; sp.blueNoise = e->dBlueNoise.p; sp.pmjPixelSamples = e->dPmjPixel.p;
    }
    if (sp.type == 5) {   // HaltonSampler::HaltonSampler (samplers.cpp:32-52), ComputeRadicalInversePermutations (lowdiscrepancy.cpp:47-55)
        std::vector<int> primes = FirstPrimes(1000);
        ECK(e->dPrimes.upload(primes));
        sp.primes = e->dPrimes.p;
        if (sp.randomize == RANDOMIZE_PERMUTE_DIGITS && (e->haltonPermSeed != sp.seed || e->dHaltonPerms.n == 0)) {
            std::vector<uint16_t> perms;
            std::vector<uint32_t> offsets;
            BuildHaltonPermutations(primes, sp.seed, &perms, &offsets);
            ECK(e->dHaltonPerms.upload(perms));
            ECK(e->dHaltonPermOffset.upload(offsets));
            e->haltonPermSeed = sp.seed;
        }
        sp.haltonPerms = e->dHaltonPerms.p; sp.haltonPermOffset = e->dHaltonPermOffset.p;
        HaltonPixelConstants(fl.xres, fl.yres, &sp);
    }
    if (sp.type == 4) {   // SobolSampler (samplers.h:482-490): the full generator matrices and SobolIntervalToIndex's matrices
        if (!e->tables.u32.count("sobol32_full") || !e->tables.u32.count("vdc_sobol") || !e->tables.u32.count("vdc_sobol_inv"))
            return e->fail(BPT_ERR_IO, "tables blob lacks the Sobol' sampler tables (regenerate it with oracle/_ref/bin/ref_dump tables)");
        auto pairs = [](const std::vector<uint32_t> &v) { std::vector<uint64_t> o(v.size() / 2); for (size_t i = 0; i < o.size(); ++i) o[i] = (uint64_t)v[2 * i] | ((uint64_t)v[2 * i + 1] << 32); return o; };
        ECK(e->dSobolFull.upload(e->tables.u32.at("sobol32_full")));
        ECK(e->dVdc.upload(pairs(e->tables.u32.at("vdc_sobol"))));
        ECK(e->dVdcInv.upload(pairs(e->tables.u32.at("vdc_sobol_inv"))));
        sp.sobolFull = e->dSobolFull.p; sp.vdc = e->dVdc.p; sp.vdcInv = e->dVdcInv.p;
        int scale = 1; while (scale < std::max(fl.xres, fl.yres)) scale <<= 1;
        sp.scale = scale; sp.log2Scale = 0; while ((1 << sp.log2Scale) < scale) ++sp.log2Scale;
    }
    {
        auto log2i = [](unsigned v) { int l = 0; while (v >>= 1) ++l; return l; };
        auto roundUpPow2 = [](int v) { int r = 1; while (r < v) r <<= 1; return r; };
        sp.log2SamplesPerPixel = log2i((unsigned)e->sampler.spp);
        int res = roundUpPow2(std::max(fl.xres, fl.yres));
        int log4SamplesPerPixel = (sp.log2SamplesPerPixel + 1) / 2;
        sp.nBase4Digits = log2i((unsigned)res) + log4SamplesPerPixel;
    }

    // ---- materials
    SpecPool spool;
    std::vector<DeviceMaterial> dmats(e->materials.size());
    std::fill(e->materialPresent, e->materialPresent + MAT_NTAGS, false);
    for (size_t i = 0; i < e->materials.size(); ++i) {
        const OwnedMaterial &om = e->materials[i];
        DeviceMaterial &dm = dmats[i];
        std::memset(&dm, 0, sizeof(dm));
        dm.type = om.m.type;
        switch (om.m.type) {
        case BPT_MAT_DIFFUSE:
            if (om.reflectance.s.kind == BPT_SPEC_NONE) { dm.reflectance.kind = SPEC_CONSTANT; dm.reflectance.c0 = 0.5f; }
            else if (!MakeSpectrumRec(e, om.reflectance, ST_ALBEDO, spool, &dm.reflectance, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            break;
        case BPT_MAT_INTERFACE: break;
        case BPT_MAT_CONDUCTOR: {   // ConductorMaterial::Create (materials.cpp): eta/k default to Cu unless "reflectance" is given
            bool haveRefl = om.reflectance.s.kind != BPT_SPEC_NONE;
            if (haveRefl && (om.eta.s.kind != BPT_SPEC_NONE || om.k.s.kind != BPT_SPEC_NONE))
                return e->fail(BPT_ERR_INVALID, "conductor: both reflectance and eta/k given");
            if (haveRefl) {
                dm.useReflectance = 1;
                if (!MakeSpectrumRec(e, om.reflectance, ST_ALBEDO, spool, &dm.reflectance, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            } else {
                OwnedSpectrum eta = om.eta, k = om.k;
                if (eta.s.kind == BPT_SPEC_NONE) { eta.s.kind = BPT_SPEC_NAMED; std::strcpy(eta.s.name, "metal-Cu-eta"); }
                if (k.s.kind == BPT_SPEC_NONE) { k.s.kind = BPT_SPEC_NAMED; std::strcpy(k.s.name, "metal-Cu-k"); }
                if (!MakeSpectrumRec(e, eta, ST_UNBOUNDED, spool, &dm.eta, &serr) || !MakeSpectrumRec(e, k, ST_UNBOUNDED, spool, &dm.k, &serr))
                    return e->fail(BPT_ERR_UNSUPPORTED, serr);
            }
            dm.uroughness = om.m.uroughness >= 0 ? om.m.uroughness : om.m.roughness;
            dm.vroughness = om.m.vroughness >= 0 ? om.m.vroughness : om.m.roughness;
            dm.flags = om.m.remap_roughness ? 1 : 0;
            break;
        }
        case BPT_MAT_DIELECTRIC: {  // DielectricMaterial::Create: "eta" float or spectrum, default 1.5
            if (om.eta.s.kind == BPT_SPEC_NONE) { dm.eta.kind = SPEC_CONSTANT; dm.eta.c0 = 1.5f; }
            else if (!MakeSpectrumRec(e, om.eta, ST_UNBOUNDED, spool, &dm.eta, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.uroughness = om.m.uroughness >= 0 ? om.m.uroughness : om.m.roughness;
            dm.vroughness = om.m.vroughness >= 0 ? om.m.vroughness : om.m.roughness;
            dm.flags = om.m.remap_roughness ? 1 : 0;
            break;
        }
        case BPT_MAT_COATED_CONDUCTOR: {   // CoatedConductorMaterial::Create (materials.cpp)
            dm.iface_uroughness = om.m.iface_uroughness >= 0 ? om.m.iface_uroughness : om.m.iface_roughness;
            dm.iface_vroughness = om.m.iface_vroughness >= 0 ? om.m.iface_vroughness : om.m.iface_roughness;
            dm.thickness = om.m.thickness > 0 ? om.m.thickness : .01f;
            if (om.iface_eta.s.kind == BPT_SPEC_NONE) { dm.iface_eta.kind = SPEC_CONSTANT; dm.iface_eta.c0 = 1.5f; }
            else if (!MakeSpectrumRec(e, om.iface_eta, ST_UNBOUNDED, spool, &dm.iface_eta, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.uroughness = om.m.uroughness >= 0 ? om.m.uroughness : om.m.roughness;
            dm.vroughness = om.m.vroughness >= 0 ? om.m.vroughness : om.m.roughness;
            bool haveRefl = om.reflectance.s.kind != BPT_SPEC_NONE;
            if (haveRefl && (om.conductor_eta.s.kind != BPT_SPEC_NONE || om.conductor_k.s.kind != BPT_SPEC_NONE))
                return e->fail(BPT_ERR_INVALID, "coatedconductor: both reflectance and conductor.eta/k given");
            if (haveRefl) {
                dm.useReflectance = 1;
                if (!MakeSpectrumRec(e, om.reflectance, ST_ALBEDO, spool, &dm.reflectance, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            } else {
                OwnedSpectrum eta = om.conductor_eta, k = om.conductor_k;
                if (eta.s.kind == BPT_SPEC_NONE) { eta.s.kind = BPT_SPEC_NAMED; std::strcpy(eta.s.name, "metal-Cu-eta"); }
                if (k.s.kind == BPT_SPEC_NONE) { k.s.kind = BPT_SPEC_NAMED; std::strcpy(k.s.name, "metal-Cu-k"); }
                if (!MakeSpectrumRec(e, eta, ST_UNBOUNDED, spool, &dm.eta, &serr) || !MakeSpectrumRec(e, k, ST_UNBOUNDED, spool, &dm.k, &serr))
                    return e->fail(BPT_ERR_UNSUPPORTED, serr);
            }
            dm.maxDepth = om.m.max_depth > 0 ? om.m.max_depth : 10;
            dm.nSamples = om.m.n_samples > 0 ? om.m.n_samples : 1;
            dm.g = om.m.g;
            if (om.albedo.s.kind == BPT_SPEC_NONE) { dm.albedo.kind = SPEC_CONSTANT; dm.albedo.c0 = 0.f; }
            else if (!MakeSpectrumRec(e, om.albedo, ST_ALBEDO, spool, &dm.albedo, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.flags = om.m.remap_roughness ? 1 : 0;
            dm.optionsSeed = e->sampler.seed;
            break;
        }
        case BPT_MAT_COATED_DIFFUSE: {   // CoatedDiffuseMaterial::Create (materials.cpp:301-345)
            if (om.reflectance.s.kind == BPT_SPEC_NONE) { dm.reflectance.kind = SPEC_CONSTANT; dm.reflectance.c0 = 0.5f; }
            else if (!MakeSpectrumRec(e, om.reflectance, ST_ALBEDO, spool, &dm.reflectance, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.uroughness = om.m.uroughness >= 0 ? om.m.uroughness : om.m.roughness;
            dm.vroughness = om.m.vroughness >= 0 ? om.m.vroughness : om.m.roughness;
            dm.thickness = om.m.thickness > 0 ? om.m.thickness : .01f;
            if (om.eta.s.kind == BPT_SPEC_NONE) { dm.eta.kind = SPEC_CONSTANT; dm.eta.c0 = 1.5f; }
            else if (!MakeSpectrumRec(e, om.eta, ST_UNBOUNDED, spool, &dm.eta, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.maxDepth = om.m.max_depth > 0 ? om.m.max_depth : 10;
            dm.nSamples = om.m.n_samples > 0 ? om.m.n_samples : 1;
            dm.g = om.m.g;
            if (om.albedo.s.kind == BPT_SPEC_NONE) { dm.albedo.kind = SPEC_CONSTANT; dm.albedo.c0 = 0.f; }
            else if (!MakeSpectrumRec(e, om.albedo, ST_ALBEDO, spool, &dm.albedo, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.flags = om.m.remap_roughness ? 1 : 0;
            dm.optionsSeed = e->sampler.seed;
            break;
        }
        case BPT_MAT_DIFFUSE_TRANSMISSION: {   // DiffuseTransmissionMaterial::Create (materials.cpp:575-596)
            if (om.reflectance.s.kind == BPT_SPEC_NONE) { dm.reflectance.kind = SPEC_CONSTANT; dm.reflectance.c0 = 0.25f; }
            else if (!MakeSpectrumRec(e, om.reflectance, ST_ALBEDO, spool, &dm.reflectance, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            if (om.transmittance.s.kind == BPT_SPEC_NONE) { dm.transmittance.kind = SPEC_CONSTANT; dm.transmittance.c0 = 0.25f; }
            else if (!MakeSpectrumRec(e, om.transmittance, ST_ALBEDO, spool, &dm.transmittance, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            dm.scale = om.m.scale > 0 ? om.m.scale : 1.f;
            break;
        }
        case BPT_MAT_THIN_DIELECTRIC: {   // ThinDielectricMaterial::Create: "eta" float or spectrum, default 1.5
            if (om.eta.s.kind == BPT_SPEC_NONE) { dm.eta.kind = SPEC_CONSTANT; dm.eta.c0 = 1.5f; }
            else if (!MakeSpectrumRec(e, om.eta, ST_UNBOUNDED, spool, &dm.eta, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
            break;
        }
        default: return e->fail(BPT_ERR_UNSUPPORTED, "material type not implemented yet");
        }
        // image textures in place of constants; which parameters a material reads as textures follows its Create() in materials.cpp
        auto texId = [&](int id, int *out) -> bool {
            *out = -1;
            if (id == 0) return true;
            if (id < 0 || id > (int)e->textures.size()) return false;
            *out = id - 1;
            return true;
        };
        if (!texId(om.m.reflectance_tex, &dm.reflectanceTex) || !texId(om.m.transmittance_tex, &dm.transmittanceTex) ||
            !texId(om.m.uroughness_tex, &dm.uroughnessTex) || !texId(om.m.vroughness_tex, &dm.vroughnessTex) ||
            !texId(om.m.displacement_tex, &dm.displacementTex) || !texId(om.m.normal_map_tex, &dm.normalMapTex))
            return e->fail(BPT_ERR_INVALID, "material refers to a texture that was not set");
        const int mt = om.m.type;
        const bool reflOK = mt == BPT_MAT_DIFFUSE || mt == BPT_MAT_COATED_DIFFUSE || mt == BPT_MAT_DIFFUSE_TRANSMISSION || mt == BPT_MAT_CONDUCTOR ||
                            mt == BPT_MAT_COATED_CONDUCTOR;
        const bool roughOK = mt == BPT_MAT_CONDUCTOR || mt == BPT_MAT_DIELECTRIC || mt == BPT_MAT_COATED_DIFFUSE || mt == BPT_MAT_COATED_CONDUCTOR;
        if ((dm.reflectanceTex >= 0 && !reflOK) || (dm.transmittanceTex >= 0 && mt != BPT_MAT_DIFFUSE_TRANSMISSION) ||
            ((dm.uroughnessTex >= 0 || dm.vroughnessTex >= 0) && !roughOK) || (dm.displacementTex >= 0 && mt == BPT_MAT_INTERFACE))
            return e->fail(BPT_ERR_UNSUPPORTED, "texture on a material parameter that does not take one");
        auto isSpec = [&](int t) { return t >= 0 && e->textures[t].t.is_spectrum != 0; };
        if ((dm.reflectanceTex >= 0 && !isSpec(dm.reflectanceTex)) || (dm.transmittanceTex >= 0 && !isSpec(dm.transmittanceTex)) || isSpec(dm.uroughnessTex) ||
            isSpec(dm.vroughnessTex) || isSpec(dm.displacementTex))
            return e->fail(BPT_ERR_INVALID, "float texture on a spectrum parameter or the other way round");
        auto kindOf = [&](int t) { return t < 0 ? -1 : e->textures[t].t.kind; };
        if (kindOf(dm.normalMapTex) == BPT_TEXTURE_IMAGEMAP || kindOf(dm.reflectanceTex) == BPT_TEXTURE_NORMALMAP || kindOf(dm.transmittanceTex) == BPT_TEXTURE_NORMALMAP ||
            kindOf(dm.uroughnessTex) == BPT_TEXTURE_NORMALMAP || kindOf(dm.vroughnessTex) == BPT_TEXTURE_NORMALMAP || kindOf(dm.displacementTex) == BPT_TEXTURE_NORMALMAP ||
            (dm.normalMapTex >= 0 && mt == BPT_MAT_INTERFACE))
            return e->fail(BPT_ERR_INVALID, "normal-map images and imagemap textures are not interchangeable");
        if (dm.reflectanceTex >= 0 && (mt == BPT_MAT_CONDUCTOR || mt == BPT_MAT_COATED_CONDUCTOR)) dm.useReflectance = 1;
    }

    // ---- meshes: transform to render space (TriangleMesh ctor util/mesh.cpp:25-70), build pools and lights
    std::vector<float> meshAlpha(e->meshes.size(), 1.f);
    std::vector<int4> meshShape(e->meshes.size(), make_int4(-1, -1, 0, 0));
    std::vector<float4> patchPool;
    std::vector<DeviceShape> dshapes;
    bool instancedQuadrics = false;   // an object definition holds a quadric: its instances keep the reference's (loose) bounds
    bool anyAlpha = false;
    std::unique_ptr<float4[]> tri;   // host-assembled ranges only (untouched pages of the rest are never committed)
    std::vector<DeviceMesh> dmesh(e->meshes.size());
    std::vector<int> indexPool;
    std::vector<float> nPool, uvPool;
    std::vector<DeviceLight> dlights;
    std::vector<LightBoundsH> lbounds;
    size_t totalTris = 0;
    for (const OwnedMesh &om : e->meshes) totalTris += om.m.n_triangles;
    if (totalTris == 0) return e->fail(BPT_ERR_INVALID, "scene has no triangles");
    if (totalTris > (size_t)1 << 30) return e->fail(BPT_ERR_INVALID, "too many triangles");
    tri.reset(new float4[3 * totalTris]);
    struct TriRange { size_t first; int count; bool device; size_t mesh; Mat4 M; uint32_t cls; };
    std::vector<TriRange> triRanges;   // pool ranges in upload order: host-filled, or assembled by k_assemble_triangles
    size_t triCursor = 0;
    size_t devVerts = 0, devIdx = 0;
    const int kDeviceAssembleMin = std::getenv("BPT_HOST_ASSEMBLE") ? (1 << 30) : 16384;
    e->haveAreaLights = false;
    const std::vector<float> &cieY = e->tables.F("cie_y");
    const std::vector<float> &illum = e->tables.F("srgb_illuminant");
    // triangle groups: group 0 = top-level shapes (scene order), then one group per object definition
    int nObjects = 0;
    for (const OwnedMesh &om : e->meshes) nObjects = std::max(nObjects, om.m.object_id + 1);
    for (const bpt_instance &in : e->instances) if (in.object_id < 0 || in.object_id >= nObjects) return e->fail(BPT_ERR_INVALID, "instance of an undefined object");
    std::vector<size_t> meshOrder;
    std::vector<int> groupStart(nObjects + 2, 0);
    for (int g = -1; g < nObjects; ++g) {
        groupStart[g + 1] = 0;
        for (size_t mi = 0; mi < e->meshes.size(); ++mi) if (e->meshes[mi].m.object_id == g) { meshOrder.push_back(mi); }
        // (group extents are fixed up after the loop below, once triangle counts are known)
    }
    {
        int acc = 0; size_t k = 0;
        for (int g = -1; g < nObjects; ++g) {
            groupStart[g + 1] = acc;
            for (; k < meshOrder.size() && e->meshes[meshOrder[k]].m.object_id == g; ++k) acc += e->meshes[meshOrder[k]].m.n_triangles;
        }
        groupStart[nObjects + 1] = acc;
    }
    for (size_t oi = 0; oi < meshOrder.size(); ++oi) {
        size_t mi = meshOrder[oi];
        const OwnedMesh &om = e->meshes[mi];
        if (om.m.object_id >= 0 && om.m.area_light >= 0) return e->fail(BPT_ERR_UNSUPPORTED, "area lights are not supported with object instancing");
        if (om.m.material >= (int)e->materials.size()) return e->fail(BPT_ERR_INVALID, "mesh material index out of range");
        Xform renderFromObject = ResolveRenderFromObject(space, spaces, om.m.world_from_object, om.m.object_from_world);   // scene.h:475-477
        const bool deviceAssemble = !om.isQuadric && !om.isBilinear && om.m.area_light < 0 && om.m.n_triangles >= kDeviceAssembleMin;
        std::vector<Pt3> P(deviceAssemble ? 0 : om.m.n_vertices);
        if (!deviceAssemble)
            for (int v = 0; v < om.m.n_vertices; ++v) P[v] = renderFromObject(Pt3(om.P[3 * v], om.P[3 * v + 1], om.P[3 * v + 2]));
        DeviceMesh &dm = dmesh[mi];
        dm.material = om.m.material;
        int matType = om.m.material < 0 ? MAT_INTERFACE : e->materials[om.m.material].m.type;
        e->materialPresent[matType] = true;
        bool flip = (om.m.reverse_orientation != 0) ^ renderFromObject.SwapsHandedness();
        dm.flags = (flip ? 1 : 0) | (!om.N.empty() ? 2 : 0) | (!om.UV.empty() ? 4 : 0);
        dm.indexOffset = (int)indexPool.size();
        dm.nOffset = (int)(nPool.size() / 3);
        dm.uvOffset = (int)(uvPool.size() / 2);
        dm.mediumInside = om.m.medium_inside; dm.mediumOutside = om.m.medium_outside;
        dm.shape = -1; dm.patchBase = -1;
        DeviceShape dshape{};
        Bounds3 shapeBounds;
        if (om.isQuadric) {
            // Sphere / Disk / Cylinder constructors (shapes.h:118-128, 407-417, 801-812) and ::Bounds (shapes.cpp:33-36, 85-88, 121-124)
            const bpt_quadric &q = om.q;
            Pt3 bmin, bmax;
            dshape = MakeQuadric(q.type == BPT_SHAPE_SPHERE ? SHAPE_SPHERE : (q.type == BPT_SHAPE_DISK ? SHAPE_DISK : SHAPE_CYLINDER), q.radius, q.zmin, q.zmax, q.phimax,
                                 q.height, q.innerradius, q.reverse_orientation != 0, renderFromObject.SwapsHandedness(), ToMat4(renderFromObject.m),
                                 ToMat4(renderFromObject.mInv), &bmin, &bmax);
            for (int c = 0; c < 8; ++c)   // Transform::operator()(const Bounds3f &) (util/transform.cpp:134-139)
                shapeBounds = Union(shapeBounds, renderFromObject(Pt3((c & 1) ? bmax.x : bmin.x, (c & 2) ? bmax.y : bmin.y, (c & 4) ? bmax.z : bmin.z)));
            dm.shape = (int)dshapes.size();
            dm.flags = 0;
            if (om.m.object_id >= 0) instancedQuadrics = true;
        }
        if (om.isBilinear) {
            dm.patchBase = (int)(patchPool.size() / 4);
            if (om.m.object_id >= 0) instancedQuadrics = true;
        }
        meshShape[mi] = make_int4(dm.shape, dm.patchBase, (int)triCursor - groupStart[om.m.object_id + 1], 0);
        if (!om.N.empty() || !om.UV.empty()) indexPool.insert(indexPool.end(), om.idx.begin(), om.idx.end());
        if (!om.N.empty())
            for (int v = 0; v < om.m.n_vertices; ++v) {
                Nrm3 nn = renderFromObject(Nrm3(om.N[3 * v], om.N[3 * v + 1], om.N[3 * v + 2]));
                if (om.m.reverse_orientation) nn = -nn;
                nPool.push_back(nn.x); nPool.push_back(nn.y); nPool.push_back(nn.z);
            }
        if (!om.UV.empty()) uvPool.insert(uvPool.end(), om.UV.begin(), om.UV.end());
        // area light (CreateLights scene.cpp; ignored for "interface" material)
        bool emissive = om.m.area_light >= 0 && matType != MAT_INTERFACE;
        dm.areaLightBase = emissive ? (int)dlights.size() : -1;
        SpectrumRec Lrec{};
        float lightScale = 1, LemitMax = 0;
        const OwnedAreaLight *al = nullptr;
        if (emissive) {
            if (om.m.area_light >= (int)e->areaLights.size()) return e->fail(BPT_ERR_INVALID, "area light index out of range");
            al = &e->areaLights[om.m.area_light];
            std::vector<float> dense(kNLambda);
            if (al->L.s.kind == BPT_SPEC_NONE) dense = illum;
            else {
                if (!MakeSpectrumRec(e, al->L, ST_ILLUMINANT, spool, &Lrec, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
                for (int i = 0; i < kNLambda; ++i) dense[i] = EvalRecAt(Lrec, spool.pool, (float)(kLambdaMin + i));
            }
            Lrec.kind = SPEC_DENSE; Lrec.table = spool.addDense(dense);
            // scale /= SpectrumToPhotometric(L): RGB illuminant spectra use the colour-space illuminant only
            float photometric = 0;
            if (al->L.s.kind == BPT_SPEC_RGB || al->L.s.kind == BPT_SPEC_NONE) { for (int i = 0; i < kNLambda; ++i) photometric += cieY[i] * illum[i]; }
            else { for (int i = 0; i < kNLambda; ++i) photometric += cieY[i] * dense[i]; }
            lightScale = al->a.scale / photometric;
            LemitMax = *std::max_element(dense.begin(), dense.end());
            e->haveAreaLights = true;
        }
        unsigned cls = (unsigned)matType | (emissive ? 0x100u : 0u);
        const bool alphaTested = om.m.has_alpha && om.m.alpha < 1;
        if (alphaTested) { cls |= kClsAlpha; anyAlpha = true; }
        meshAlpha[mi] = alphaTested ? om.m.alpha : 1.f;
        triRanges.push_back(TriRange{triCursor, om.m.n_triangles, deviceAssemble, mi, ToMat4(renderFromObject.m), cls});
        if (deviceAssemble) {
            devVerts = std::max(devVerts, (size_t)om.m.n_vertices); devIdx = std::max(devIdx, (size_t)om.m.n_triangles);
            triCursor += om.m.n_triangles;
            continue;
        }
        if (om.isQuadric) {
            // one pool record: the bounds where a triangle has its vertices (p2 = p0: the builder flags it degenerate, so only the
            // quadric test of the traversal ever looks at it), class bit kClsShape
            int gtri = (int)triCursor;
            float4 *tp = &tri[3 * triCursor++];
            tp[0] = make_float4(shapeBounds.pMin.x, shapeBounds.pMin.y, shapeBounds.pMin.z, b2f((uint32_t)mi));
            tp[1] = make_float4(shapeBounds.pMax.x, shapeBounds.pMax.y, shapeBounds.pMax.z, b2f(0u));
            tp[2] = make_float4(shapeBounds.pMin.x, shapeBounds.pMin.y, shapeBounds.pMin.z, b2f(cls | kClsShape));
            if (emissive) {
                DeviceLight dl{};
                dl.type = LIGHT_DIFFUSE_AREA; dl.tri = gtri; dl.twoSided = al->a.two_sided; dl.LemitTable = Lrec.table;
                float area = QuadricArea(dshape);
                float scale = lightScale;
                if (al->a.power > 0) {
                    float k_e = 1;
                    k_e *= (al->a.two_sided ? 2 : 1) * area * kPi;
                    scale *= al->a.power / k_e;
                }
                dl.scale = scale;
                dlights.push_back(dl);
                // DiffuseAreaLight::Bounds (lights.cpp:788-807) with Shape::NormalBounds: the whole sphere of directions for Sphere /
                // Cylinder (shapes.h:134, 594), the transformed +z normal for a Disk (shapes.cpp:90-95)
                float phi = LemitMax;
                phi *= scale * area * kPi;
                Vec3 w(0, 0, 1); float cosTheta = -1;
                if (dshape.type == SHAPE_DISK) {
                    Nrm3 n = renderFromObject(Nrm3(0, 0, 1));
                    if (om.q.reverse_orientation) n = -n;
                    w = Normalize(V(n)); cosTheta = 1;
                }
                lbounds.push_back(LightBoundsH(shapeBounds, w, phi, cosTheta, std::cos(kPi / 2), al->a.two_sided != 0));
            }
            dshapes.push_back(dshape);
            continue;
        }
        if (om.isBilinear) {
            // one pool record per patch: its bounds (BilinearPatch::Bounds, shapes.cpp:1071-1079), class bit kClsShape; vertices in the patch pool
            for (int t = 0; t < om.m.n_triangles; ++t) {
                PatchData pd;
                pd.p00 = P[om.idx[4 * t]]; pd.p10 = P[om.idx[4 * t + 1]]; pd.p01 = P[om.idx[4 * t + 2]]; pd.p11 = P[om.idx[4 * t + 3]];
                pd.flipNormal = flip; pd.hasN = !om.N.empty(); pd.hasUV = !om.UV.empty();
                if (pd.hasN) {
                    size_t nb = 3 * (size_t)dm.nOffset;
                    auto NN = [&](int v) { return Nrm3(nPool[nb + 3 * v], nPool[nb + 3 * v + 1], nPool[nb + 3 * v + 2]); };
                    pd.n00 = NN(om.idx[4 * t]); pd.n10 = NN(om.idx[4 * t + 1]); pd.n01 = NN(om.idx[4 * t + 2]); pd.n11 = NN(om.idx[4 * t + 3]);
                }
                pd.isRectangle = PatchIsRectangle(pd.p00, pd.p10, pd.p01, pd.p11);
                pd.area = PatchArea(pd.p00, pd.p10, pd.p01, pd.p11);
                patchPool.push_back(make_float4(pd.p00.x, pd.p00.y, pd.p00.z, pd.area));
                patchPool.push_back(make_float4(pd.p10.x, pd.p10.y, pd.p10.z, b2f(pd.isRectangle ? 1u : 0u)));
                patchPool.push_back(make_float4(pd.p01.x, pd.p01.y, pd.p01.z, 0.f));
                patchPool.push_back(make_float4(pd.p11.x, pd.p11.y, pd.p11.z, 0.f));
                Bounds3 pb = Union(Bounds3(pd.p00, pd.p01), Bounds3(pd.p10, pd.p11));
                int gtri = (int)triCursor;
                float4 *tp = &tri[3 * triCursor++];
                tp[0] = make_float4(pb.pMin.x, pb.pMin.y, pb.pMin.z, b2f((uint32_t)mi));
                tp[1] = make_float4(pb.pMax.x, pb.pMax.y, pb.pMax.z, b2f((uint32_t)t));
                tp[2] = make_float4(pb.pMin.x, pb.pMin.y, pb.pMin.z, b2f(cls | kClsShape));
                if (emissive) {
                    DeviceLight dl{};
                    dl.type = LIGHT_DIFFUSE_AREA; dl.tri = gtri; dl.twoSided = al->a.two_sided; dl.LemitTable = Lrec.table;
                    float scale = lightScale;
                    if (al->a.power > 0) {
                        float k_e = 1;
                        k_e *= (al->a.two_sided ? 2 : 1) * pd.area * kPi;
                        scale *= al->a.power / k_e;
                    }
                    dl.scale = scale;
                    dlights.push_back(dl);
                    float phi = LemitMax;
                    phi *= scale * pd.area * kPi;
                    Vec3 w; float cosTheta;
                    PatchNormalBounds(pd, &w, &cosTheta);   // BilinearPatch::NormalBounds (shapes.cpp:1081-1130)
                    lbounds.push_back(LightBoundsH(pb, w, phi, cosTheta, std::cos(kPi / 2), al->a.two_sided != 0));
                }
            }
            continue;
        }
        for (int t = 0; t < om.m.n_triangles; ++t) {
            Pt3 p0 = P[om.idx[3 * t]], p1 = P[om.idx[3 * t + 1]], p2 = P[om.idx[3 * t + 2]];
            int gtri = (int)triCursor;
            float4 *tp = &tri[3 * triCursor++];
            tp[0] = make_float4(p0.x, p0.y, p0.z, b2f((uint32_t)mi));
            tp[1] = make_float4(p1.x, p1.y, p1.z, b2f((uint32_t)t));
            tp[2] = make_float4(p2.x, p2.y, p2.z, b2f(cls));
            if (emissive) {
                DeviceLight dl{};
                dl.type = LIGHT_DIFFUSE_AREA; dl.tri = gtri; dl.twoSided = al->a.two_sided; dl.LemitTable = Lrec.table;
                float area = 0.5f * Length(Cross(p1 - p0, p2 - p0));
                float scale = lightScale;
                if (al->a.power > 0) {
                    float k_e = 1;
                    k_e *= (al->a.two_sided ? 2 : 1) * area * kPi;
                    scale *= al->a.power / k_e;
                }
                dl.scale = scale;
                dlights.push_back(dl);
                // DiffuseAreaLight::Bounds (lights.cpp) + Triangle::NormalBounds (shapes.cpp:283-298)
                float phi = LemitMax;
                phi *= scale * area * kPi;
                Nrm3 n = Normalize(Nrm3(Cross(p1 - p0, p2 - p0)));
                if (!om.N.empty()) {
                    size_t nb = 3 * (size_t)dm.nOffset;
                    auto NN = [&](int v) { return Nrm3(nPool[nb + 3 * v], nPool[nb + 3 * v + 1], nPool[nb + 3 * v + 2]); };
                    Nrm3 ns = NN(om.idx[3 * t]) + NN(om.idx[3 * t + 1]) + NN(om.idx[3 * t + 2]);
                    n = FaceForward(n, ns);
                } else if (flip) n = n * -1.f;
                Bounds3 b = Union(Bounds3(p0, p1), p2);
                lbounds.push_back(LightBoundsH(b, Normalize(V(n)), phi, 1.f /* cone cosTheta */, std::cos(kPi / 2), al->a.two_sided != 0));
            }
        }
    }
    // non-area lights follow the area lights (BasicScene::CreateLights, scene.cpp)
    std::vector<DeviceImageLight> dimg;
    struct ImgTables { std::vector<float> rgb; FilterTables dist, comp; int res; };
    std::vector<ImgTables> imgTables;
    for (const OwnedLight &ol : e->lights) {
        DeviceLight dl{};
        Xform renderFromLight = ResolveRenderFromObject(space, spaces, ol.l.world_from_light, ol.l.light_from_world);
        dl.renderFromLight = ToMat4(renderFromLight.m); dl.lightFromRender = ToMat4(renderFromLight.mInv);
        float photometric = 0;
        for (int i = 0; i < kNLambda; ++i) photometric += cieY[i] * illum[i];
        float scale = ol.l.scale;
        if (ol.l.type == BPT_LIGHT_POINT || ol.l.type == BPT_LIGHT_SPOT || ol.l.type == BPT_LIGHT_DISTANT) {
            // PointLight / SpotLight / DistantLight::Create (lights.cpp:192-213, 1433-1464, 246-276) and ::Bounds (:168-173, 1370-1380)
            std::vector<float> dense(kNLambda);
            float ph = 0;
            if (ol.L.s.kind == BPT_SPEC_NONE) { dense = illum; ph = photometric; }
            else {
                SpectrumRec Lr{};
                if (!MakeSpectrumRec(e, ol.L, ST_ILLUMINANT, spool, &Lr, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
                for (int i = 0; i < kNLambda; ++i) dense[i] = EvalRecAt(Lr, spool.pool, (float)(kLambdaMin + i));
                if (ol.L.s.kind == BPT_SPEC_RGB) ph = photometric; else for (int i = 0; i < kNLambda; ++i) ph += cieY[i] * dense[i];
            }
            scale /= ph;
            dl.LemitTable = spool.addDense(dense);
            const float maxValue = *std::max_element(dense.begin(), dense.end());
            const Pt3 from(ol.l.from[0], ol.l.from[1], ol.l.from[2]), to(ol.l.to[0], ol.l.to[1], ol.l.to[2]);
            Xform fin;
            if (ol.l.type == BPT_LIGHT_POINT) {
                dl.type = LIGHT_POINT;
                if (ol.l.power > 0) { float k_e = 4 * kPi; scale *= ol.l.power / k_e; }
                fin = renderFromLight * Translate(Vec3(from.x, from.y, from.z));
                dl.p = fin(Pt3(0, 0, 0));
                float phi = 4 * kPi * scale * maxValue;
                lbounds.push_back(LightBoundsH(Bounds3(dl.p, dl.p), Vec3(0, 0, 1), phi, std::cos(kPi), std::cos(kPi / 2), false));
            } else if (ol.l.type == BPT_LIGHT_SPOT) {
                dl.type = LIGHT_SPOT;
                const float coneangle = ol.l.cone_angle, conedelta = ol.l.cone_delta;
                Frame fz = Frame::FromZ(Normalize(to - from));
                Xform dirToZ(M4(fz.x.x, fz.x.y, fz.x.z, 0, fz.y.x, fz.y.y, fz.y.z, 0, fz.z.x, fz.z.y, fz.z.z, 0, 0, 0, 0, 1));
                Xform t = Translate(Vec3(from.x, from.y, from.z)) * Inverse(dirToZ);
                fin = renderFromLight * t;
                if (ol.l.power > 0) {
                    float cosFalloffEnd = std::cos(Radians(coneangle));
                    float cosFalloffStart = std::cos(Radians(coneangle - conedelta));
                    float k_e = 2 * kPi * ((1 - cosFalloffStart) + (cosFalloffStart - cosFalloffEnd) / 2);
                    scale *= ol.l.power / k_e;
                }
                dl.cosFalloffEnd = std::cos(Radians(coneangle));
                dl.cosFalloffStart = std::cos(Radians(coneangle - conedelta));
                dl.p = fin(Pt3(0, 0, 0));
                Vec3 w = Normalize(fin(Vec3(0, 0, 1)));
                float phi = scale * maxValue * 4 * kPi;
                float cosTheta_e = std::cos(std::acos(dl.cosFalloffEnd) - std::acos(dl.cosFalloffStart));
                if (cosTheta_e == 1 && dl.cosFalloffEnd != dl.cosFalloffStart) cosTheta_e = 0.999f;
                lbounds.push_back(LightBoundsH(Bounds3(dl.p, dl.p), w, phi, dl.cosFalloffStart, cosTheta_e, false));
            } else {
                dl.type = LIGHT_DISTANT;
                Vec3 w = Normalize(from - to), v1, v2;
                CoordinateSystem(w, &v1, &v2);
                Xform t(M4(v1.x, v2.x, w.x, 0, v1.y, v2.y, w.y, 0, v1.z, v2.z, w.z, 0, 0, 0, 0, 1));
                fin = renderFromLight * t;
                if (ol.l.illuminance > 0) scale *= ol.l.illuminance;
                dl.w = Normalize(fin(Vec3(0, 0, 1)));
                lbounds.push_back(LightBoundsH());   // DistantLight::Bounds() is empty (lights.h:273)
            }
            dl.renderFromLight = ToMat4(fin.m); dl.lightFromRender = ToMat4(fin.mInv);
            dl.scale = scale;
            dlights.push_back(dl);
            continue;
        }
        if (ol.l.type == BPT_LIGHT_UNIFORM_INFINITE) {
            dl.type = LIGHT_UNIFORM_INFINITE;
            std::vector<float> dense(kNLambda);
            if (ol.L.s.kind == BPT_SPEC_NONE) { dense = illum; scale /= photometric; }
            else {
                SpectrumRec Lr{};
                if (!MakeSpectrumRec(e, ol.L, ST_ILLUMINANT, spool, &Lr, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
                for (int i = 0; i < kNLambda; ++i) dense[i] = EvalRecAt(Lr, spool.pool, (float)(kLambdaMin + i));
                float ph = 0;
                if (ol.L.s.kind == BPT_SPEC_RGB) ph = photometric; else for (int i = 0; i < kNLambda; ++i) ph += cieY[i] * dense[i];
                scale /= ph;
            }
            if (ol.l.illuminance > 0) { float k_e = kPi; scale *= ol.l.illuminance / k_e; }
            dl.LemitTable = spool.addDense(dense);
        } else {
            dl.type = LIGHT_IMAGE_INFINITE;
            int res = ol.l.image_res;
            if (res <= 0 || ol.image.size() != 3 * (size_t)res * res) return e->fail(BPT_ERR_INVALID, "infinite light image missing");
            scale /= photometric;
            if (ol.l.illuminance > 0) {   // lights.cpp:1620-1647
                float illuminance = 0;
                const std::vector<float> &Mx = e->tables.F("srgb_xyz_from_rgb");
                float lum[3] = {Mx[3], Mx[4], Mx[5]};
                for (int y = 0; y < res; ++y) {
                    float v = (float(y) + 0.5f) / float(res);
                    for (int x = 0; x < res; ++x) {
                        float u = (x + 0.5f) / res;
                        Vec3 w = EqualAreaSquareToSphere(Vec2(u, v));
                        if (w.z <= 0) continue;
                        for (int c = 0; c < 3; ++c) illuminance += ol.image[3 * ((size_t)y * res + x) + c] * lum[c] * w.z;
                    }
                }
                illuminance *= 2 * kPi / (res * res);
                scale *= ol.l.illuminance / illuminance;
            }
            imgTables.emplace_back();
            ImgTables &it = imgTables.back();
            it.res = res; it.rgb = ol.image;
            // Image::GetSamplingDistribution (util/image.h:449-470) with dxdA = 1 over [0,1]^2, then the compensated copy
            std::vector<float> d((size_t)res * res);
            for (size_t i = 0; i < d.size(); ++i) { float sum = 0; sum += ol.image[3 * i]; sum += ol.image[3 * i + 1]; sum += ol.image[3 * i + 2]; d[i] = (sum / 3) * 1.f; }
            BuildPC2D(d, res, res, 0, 1, 0, 1, &it.dist);
            double acc = 0; for (float v : d) acc += v;
            float average = (float)(acc / d.size());
            bool allZero = true;
            for (float &v : d) { v = stdmax(v - average, 0.f); if (v != 0) allZero = false; }
            if (allZero) std::fill(d.begin(), d.end(), 1.f);
            BuildPC2D(d, res, res, 0, 1, 0, 1, &it.comp);
            dl.image = (int)imgTables.size() - 1;
        }
        dl.scale = scale;
        dlights.push_back(dl);
        lbounds.push_back(LightBoundsH());   // no bounds: infinite light
    }
    // a scene without lights is legal (it renders black; the raw ray entry points still work)

    // ---- light sampler: LightSampler::Create(strategy, lights) as the CPU VolPathIntegrator does (cpu/integrators.h:270) -- the parity
    // oracle.  (The reference's wavefront integrator swaps in a uniform sampler when there is exactly one light,
    // wavefront/integrator.cpp:182-186; with one bounded light that differs from the BVH sampler wherever the light's importance is
    // zero, so the CPU behaviour is the one followed.)
    LightBVHBuilder lb;
    DeviceLightSampler dls{};
    dls.nLights = (int)dlights.size();
    int lsType = e->integ.light_sampler;
    dls.type = lsType;
    if (lsType == BPT_LIGHTSAMPLER_BVH) lb.Build(lbounds);
    else {
        lb.bitTrails.assign(dlights.size(), 0); lb.lightInBVH.assign(dlights.size(), 0);
        // the integrator's own list of infinite lights (cpu/integrators.cpp: infiniteLights) does not depend on the sampler
        for (size_t i = 0; i < lbounds.size(); ++i) if (!lbounds[i].valid) lb.infiniteLights.push_back((int)i);
    }
    dls.nInfinite = (int)lb.infiniteLights.size();
    dls.nNodes = (int)lb.nodes.size();

    lap("host scene preparation");
    // ---- media (HomogeneousMedium / GridMedium constructors, media.h:222-232, media.cpp:212-247)
    std::vector<DeviceMedium> dmedia(e->media.size());
    e->mediumBufs.clear();
    e->mediumBufs4.clear();
    for (size_t i = 0; i < e->media.size(); ++i) {
        const OwnedMedium &om = e->media[i];
        DeviceMedium &dm = dmedia[i];
        std::memset(&dm, 0, sizeof(dm));
        dm.type = om.m.type == BPT_MEDIUM_HOMOGENEOUS ? MEDIUM_HOMOGENEOUS : om.m.type == BPT_MEDIUM_RGBGRID ? MEDIUM_RGBGRID : MEDIUM_GRID;
        dm.g = om.m.g;
        if (dm.type == MEDIUM_RGBGRID) {
            // RGBGridMedium::Create / constructor (media.cpp:339-453): voxels become RGBUnboundedSpectrum / RGBIlluminantSpectrum, the
            // 16^3 majorant grid holds sigmaScale * (max sigma_a + max sigma_s) over the voxels a cell overlaps
            Xform renderFromMedium = ResolveRenderFromObject(space, spaces, om.m.world_from_medium, om.m.medium_from_world);
            dm.mInv = ToMat4(renderFromMedium.mInv);
            for (int k = 0; k < 3; ++k) { dm.bmin[k] = om.m.bounds[k]; dm.bmax[k] = om.m.bounds[3 + k]; }
            const int nx = om.m.nx, ny = om.m.ny, nz = om.m.nz;
            const size_t nv = (size_t)nx * ny * nz;
            dm.nx = nx; dm.ny = ny; dm.nz = nz;
            dm.sigmaScale = om.m.scale; dm.rgbLeScale = om.m.Le_scale; dm.Le = -1;
            dm.illum = spool.addDense(illum);
            auto voxels = [&](const std::vector<float> &rgb, std::vector<float4> *out, std::vector<float> *maxv) {
                out->resize(nv); if (maxv) maxv->resize(nv);
                for (size_t v = 0; v < nv; ++v) {
                    const float *c = &rgb[3 * v];
                    float m = stdmax(stdmax(c[0], c[1]), c[2]), scale = 2 * m, q[3] = {0, 0, 0}, co[3];
                    if (scale) { q[0] = c[0] / scale; q[1] = c[1] / scale; q[2] = c[2] / scale; }
                    RGBToCoeffs(e->tables, q, co);
                    (*out)[v] = make_float4(scale, co[0], co[1], co[2]);
                    if (maxv) {   // RGBUnboundedSpectrum::MaxValue = scale * RGBSigmoidPolynomial::MaxValue (util/color.h:346-352)
                        float result = stdmax(SigmoidPoly(co[0], co[1], co[2], 360.f), SigmoidPoly(co[0], co[1], co[2], 830.f));
                        float lambda = -co[1] / (2 * co[0]);
                        if (lambda >= 360 && lambda <= 830) result = stdmax(result, SigmoidPoly(co[0], co[1], co[2], lambda));
                        (*maxv)[v] = scale * result;
                    }
                }
            };
            std::vector<float4> va, vs, vl;
            std::vector<float> ma, ms;
            auto up4 = [&](const std::vector<float4> &v, const float4 **dst) -> cudaError_t {
                e->mediumBufs4.emplace_back(new DevBuf<float4>());
                cudaError_t rcu = e->mediumBufs4.back()->upload(v);
                *dst = e->mediumBufs4.back()->p;
                return rcu;
            };
            if (!om.rgbA.empty()) { voxels(om.rgbA, &va, &ma); ECK(up4(va, &dm.rgbA)); }
            if (!om.rgbS.empty()) { voxels(om.rgbS, &vs, &ms); ECK(up4(vs, &dm.rgbS)); }
            if (!om.rgbLe.empty()) { voxels(om.rgbLe, &vl, nullptr); ECK(up4(vl, &dm.rgbLe)); }
            const int R = 16;
            dm.mres[0] = dm.mres[1] = dm.mres[2] = R;
            std::vector<float> maj((size_t)R * R * R), majA((size_t)R * R * R, 1.f), majS((size_t)R * R * R, 1.f);
            if (!ma.empty()) ComputeMajorantGrid(ma.data(), nx, ny, nz, R, majA.data());
            if (!ms.empty()) ComputeMajorantGrid(ms.data(), nx, ny, nz, R, majS.data());
            for (size_t c = 0; c < maj.size(); ++c) { float maxSigma_t = majA[c] + majS[c]; maj[c] = om.m.scale * maxSigma_t; }
            e->mediumBufs.emplace_back(new DevBuf<float>());
            ECK(e->mediumBufs.back()->upload(maj));
            dm.majorant = e->mediumBufs.back()->p;
            continue;
        }
        for (int k = 0; k < 2; ++k) {
            // DenselySampledSpectrum(spec) at integer wavelengths, then Scale(sigmaScale)
            const OwnedSpectrum &os = k == 0 ? om.sigma_a : om.sigma_s;
            std::vector<float> v(kNLambda, 1.f);   // no parameter: ConstantSpectrum(1)
            if (os.s.kind != BPT_SPEC_NONE) {
                SpectrumRec r; std::string serr;
                if (!MakeSpectrumRec(e, os, ST_UNBOUNDED, spool, &r, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, "medium spectrum: " + serr);
                for (int l = 0; l < kNLambda; ++l) v[l] = EvalRecAt(r, spool.pool, (float)(kLambdaMin + l));
            }
            for (float &x : v) x *= om.m.scale;
            (k == 0 ? dm.sigmaA : dm.sigmaS) = spool.addDense(v);
        }
        // emission (HomogeneousMedium::Create media.cpp:165-201, GridMedium::Create :249-330): Le is an illuminant spectrum normalised
        // by its photometric integral; homogeneous media fold Lescale into the table, grid media keep it in the LeScale grid
        dm.Le = -1;
        std::vector<float> leGrid;
        if (om.Le.s.kind != BPT_SPEC_NONE) {
            SpectrumRec r; std::string serr;
            if (!MakeSpectrumRec(e, om.Le, ST_ILLUMINANT, spool, &r, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, "medium Le: " + serr);
            std::vector<float> v(kNLambda);
            for (int l = 0; l < kNLambda; ++l) v[l] = EvalRecAt(r, spool.pool, (float)(kLambdaMin + l));
            float maxV = *std::max_element(v.begin(), v.end());
            if (maxV > 0) {
                // SpectrumToPhotometric (util/spectrum.cpp): an RGBIlluminantSpectrum contributes its illuminant only
                float photometric = 0;
                if (om.Le.s.kind == BPT_SPEC_RGB) { for (int l = 0; l < kNLambda; ++l) photometric += cieY[l] * illum[l]; }
                else { for (int l = 0; l < kNLambda; ++l) photometric += cieY[l] * v[l]; }
                if (dm.type == MEDIUM_HOMOGENEOUS) {
                    float LeScale = om.m.Le_scale / photometric;
                    for (float &x : v) x *= LeScale;
                } else {
                    float LeNorm = 1 / photometric;
                    if (om.leScale.empty()) leGrid = {LeNorm};
                    else { leGrid = om.leScale; for (float &x : leGrid) x *= LeNorm; }
                }
                dm.Le = spool.addDense(v);
            }
        }
        if (dm.type == MEDIUM_GRID) {
            Xform renderFromMedium = ResolveRenderFromObject(space, spaces, om.m.world_from_medium, om.m.medium_from_world);   // RenderFromObject() (scene.h:475-477)
            dm.mInv = ToMat4(renderFromMedium.mInv);
            for (int k = 0; k < 3; ++k) { dm.bmin[k] = om.m.bounds[k]; dm.bmax[k] = om.m.bounds[3 + k]; }
            const int nx = om.m.nx, ny = om.m.ny, nz = om.m.nz;
            dm.nx = nx; dm.ny = ny; dm.nz = nz;
            const int R = 16;
            dm.mres[0] = dm.mres[1] = dm.mres[2] = R;
            std::vector<float> maj((size_t)R * R * R);
            ComputeMajorantGrid(om.density.data(), nx, ny, nz, R, maj.data());
            e->mediumBufs.emplace_back(new DevBuf<float>());
            ECK(e->mediumBufs.back()->upload(om.density));
            dm.density = e->mediumBufs.back()->p;
            e->mediumBufs.emplace_back(new DevBuf<float>());
            ECK(e->mediumBufs.back()->upload(maj));
            dm.majorant = e->mediumBufs.back()->p;
            if (dm.Le >= 0) {
                e->mediumBufs.emplace_back(new DevBuf<float>());
                ECK(e->mediumBufs.back()->upload(leGrid));
                dm.leScale = e->mediumBufs.back()->p;
                if (leGrid.size() == 1) dm.lnx = dm.lny = dm.lnz = 1; else { dm.lnx = nx; dm.lny = ny; dm.lnz = nz; }
            }
        }
    }
    ECK(e->dMedia.upload(dmedia));
    e->haveInterface = e->materialPresent[MAT_INTERFACE];
    for (const OwnedMesh &om : e->meshes)
        if (om.m.medium_inside >= (int)e->media.size() || om.m.medium_outside >= (int)e->media.size()) return e->fail(BPT_ERR_INVALID, "mesh medium index out of range");
    if (e->cameraMedium >= (int)e->media.size()) return e->fail(BPT_ERR_INVALID, "camera medium index out of range");
    if ((e->haveInterface || !e->media.empty()) && !e->instances.empty())
        return e->fail(BPT_ERR_UNSUPPORTED, "media / interface surfaces are not supported together with object instancing");

    // ---- upload
    ECK(e->triPool.alloc(3 * totalTris));
    if (devVerts) { ECK(e->rawP.alloc(3 * devVerts)); ECK(e->rawIdx.alloc(3 * devIdx)); }
    for (const TriRange &r : triRanges) {
        if (r.count == 0) continue;
        if (!r.device) { ECK(cudaMemcpyAsync(e->triPool.p + 3 * r.first, tri.get() + 3 * r.first, sizeof(float4) * 3 * (size_t)r.count, cudaMemcpyHostToDevice, e->stream)); continue; }
        const OwnedMesh &om = e->meshes[r.mesh];
        // (pageable host memory: the copies return once staged, so the single staging pair can be reused mesh after mesh in stream order)
        ECK(cudaMemcpyAsync(e->rawP.p, om.P.data(), sizeof(float) * 3 * (size_t)om.m.n_vertices, cudaMemcpyHostToDevice, e->stream));
        ECK(cudaMemcpyAsync(e->rawIdx.p, om.idx.data(), sizeof(int) * 3 * (size_t)r.count, cudaMemcpyHostToDevice, e->stream));
        k_assemble_triangles<<<(r.count + 255) / 256, 256, 0, e->stream>>>(e->rawP.p, e->rawIdx.p, r.count, r.M, (uint32_t)r.mesh, r.cls, e->triPool.p + 3 * r.first);
    }
    ECK(cudaStreamSynchronize(e->stream));
    ECK(e->dMeshes.upload(dmesh));
    ECK(e->dMeshAlpha.upload(meshAlpha));
    ECK(e->dMeshShape.upload(meshShape));
    if (!dshapes.empty()) ECK(e->dShapes.upload(dshapes));
    if (!patchPool.empty()) ECK(e->dPatches.upload(patchPool));
    e->haveShapes = !dshapes.empty() || !patchPool.empty();
    (void)anyAlpha;
    ECK(e->indexPool.upload(indexPool));
    ECK(e->nPool.upload(nPool));
    ECK(e->uvPool.upload(uvPool));
    ECK(e->specPool.upload(spool.pool));
    ECK(e->dMaterials.upload(dmats));
    ECK(e->dLights.upload(dlights));
    ECK(e->dLightNodes.upload(lb.nodes));
    ECK(e->dInfinite.upload(lb.infiniteLights));
    ECK(e->dLightInBVH.upload(lb.lightInBVH));
    ECK(e->dBitTrails.upload(lb.bitTrails));
    {
        std::vector<float> tb;
        tb.insert(tb.end(), e->tables.F("cie_x").begin(), e->tables.F("cie_x").end());
        tb.insert(tb.end(), cieY.begin(), cieY.end());
        tb.insert(tb.end(), e->tables.F("cie_z").begin(), e->tables.F("cie_z").end());
        tb.insert(tb.end(), illum.begin(), illum.end());
        ECK(e->dTables.upload(tb));
    }
    if (dflt.sampled) {
        std::vector<float> fb;
        size_t oFunc = 0; fb.insert(fb.end(), ft.func.begin(), ft.func.end());
        size_t oCdf = fb.size(); fb.insert(fb.end(), ft.cdf.begin(), ft.cdf.end());
        size_t oFI = fb.size(); fb.insert(fb.end(), ft.funcInt.begin(), ft.funcInt.end());
        size_t oMF = fb.size(); fb.insert(fb.end(), ft.mfunc.begin(), ft.mfunc.end());
        size_t oMC = fb.size(); fb.insert(fb.end(), ft.mcdf.begin(), ft.mcdf.end());
        size_t oF = fb.size(); fb.insert(fb.end(), ft.f.begin(), ft.f.end());
        ECK(e->dFilter.upload(fb));
        dflt.distrib.func = e->dFilter.p + oFunc; dflt.distrib.cdf = e->dFilter.p + oCdf; dflt.distrib.funcInt = e->dFilter.p + oFI;
        dflt.distrib.mfunc = e->dFilter.p + oMF; dflt.distrib.mcdf = e->dFilter.p + oMC; dflt.distrib.mfuncInt = ft.mfuncInt;
        dflt.distrib.nu = ft.nx; dflt.distrib.nv = ft.ny;
        dflt.distrib.min0 = -dflt.radiusX; dflt.distrib.max0 = dflt.radiusX; dflt.distrib.min1 = -dflt.radiusY; dflt.distrib.max1 = dflt.radiusY;
        dflt.f = e->dFilter.p + oF;
    }

    // ---- BVH (single level, or per-definition BVHs + a top-level BVH over instances)
    FreeWideBVH(&e->bvh);
    for (WideBVH &b : e->blas) FreeWideBVH(&b);
    e->blas.clear();
    char berr[256] = {0};
    lap("uploads");
    auto tb0 = std::chrono::steady_clock::now();
    std::vector<DeviceInstance> dinst;
    const bool instanced = !e->instances.empty();
    if (!instanced) {
        if (nObjects > 0) return e->fail(BPT_ERR_INVALID, "object definitions without instances");
        if (!BuildWideBVH(e->triPool.p, (int)totalTris, &e->bvh, e->stream, berr, 0, e->treeletRounds)) return e->fail(BPT_ERR_CUDA, berr);
    } else {
        e->blas.resize(nObjects + 1);
        std::vector<float4> tlasPrims;
        std::vector<int> instGroup;
        auto addInst = [&](int group, const Xform *T) {
            const WideBVH &b = e->blas[group];
            DeviceInstance di{};
            di.nodes = b.nodes; di.tris = b.tris; di.triBase = groupStart[group]; di.identity = T ? 0 : 1;
            Bounds3 bb(Pt3(b.bmin[0], b.bmin[1], b.bmin[2]), Pt3(b.bmax[0], b.bmax[1], b.bmax[2]));
            if (T) {
                di.m = ToMat4(T->m); di.mInv = ToMat4(T->mInv);
                Bounds3 bt;   // Transform::operator()(const Bounds3f &) (util/transform.cpp:134-139)
                for (int c = 0; c < 8; ++c)
                    bt = Union(bt, (*T)(Pt3((c & 1) ? bb.pMax.x : bb.pMin.x, (c & 2) ? bb.pMax.y : bb.pMin.y, (c & 4) ? bb.pMax.z : bb.pMin.z)));
                bb = bt;
            } else { M4 I; di.m = ToMat4(I); di.mInv = ToMat4(I); }
            int id = (int)dinst.size();
            dinst.push_back(di);
            instGroup.push_back(group);
            // top-level "primitive": a degenerate triangle whose bounds are the instance bounds; p0.w carries the instance id
            tlasPrims.push_back(make_float4(bb.pMin.x, bb.pMin.y, bb.pMin.z, b2f((uint32_t)id)));
            tlasPrims.push_back(make_float4(bb.pMax.x, bb.pMax.y, bb.pMax.z, 0.f));
            tlasPrims.push_back(make_float4(bb.pMin.x, bb.pMin.y, bb.pMin.z, 0.f));
        };
        for (int g = 0; g <= nObjects; ++g) {
            int n = groupStart[g + 1] - groupStart[g];
            if (n == 0) continue;
            if (!BuildWideBVH(e->triPool.p + 3 * (size_t)groupStart[g], n, &e->blas[g], e->stream, berr, 0, e->treeletRounds)) return e->fail(BPT_ERR_CUDA, berr);
        }
        if (groupStart[1] > groupStart[0]) addInst(0, nullptr);
        for (const bpt_instance &in : e->instances) {
            if (!e->blas[in.object_id + 1].nodes) continue;   // empty definition
            // scene.cpp:375-393: renderFromInstance = RenderFromObject(0) * worldFromRender
            Xform renderFromInstance = ResolveRenderFromInstance(space, spaces, in.world_from_instance, in.instance_from_world);
            addInst(in.object_id + 1, &renderFromInstance);
        }
        if (dinst.empty()) return e->fail(BPT_ERR_INVALID, "scene has no geometry");
        ECK(e->dTlasPrims.upload(tlasPrims));
        ECK(e->dInstances.upload(dinst));
        // scene bounds as the reference computes them (union of the transformed object AABBs): what the infinite lights see
        Bounds3 looseScene;
        for (size_t k = 0; k < dinst.size(); ++k) {
            looseScene = Union(looseScene, Pt3(tlasPrims[3 * k].x, tlasPrims[3 * k].y, tlasPrims[3 * k].z));
            looseScene = Union(looseScene, Pt3(tlasPrims[3 * k + 1].x, tlasPrims[3 * k + 1].y, tlasPrims[3 * k + 1].z));
        }
        if (!std::getenv("BPT_LOOSE_INSTANCE_BOUNDS") && !instancedQuadrics) {
            std::vector<int> first(dinst.size()), count(dinst.size());
            for (size_t k = 0; k < dinst.size(); ++k) { first[k] = groupStart[instGroup[k]]; count[k] = groupStart[instGroup[k] + 1] - groupStart[instGroup[k]]; }
            ECK(e->dInstFirstTri.upload(first)); ECK(e->dInstTriCount.upload(count));
            k_instance_bounds<<<(unsigned)dinst.size(), 128, 0, e->stream>>>(e->triPool.p, e->dInstances.p, e->dInstFirstTri.p, e->dInstTriCount.p, (int)dinst.size(), e->dTlasPrims.p);
        }
        {
            const char *tl = std::getenv("BPT_TLAS_LEAF");
            if (!BuildWideBVH(e->dTlasPrims.p, (int)dinst.size(), &e->bvh, e->stream, berr, tl ? std::atoi(tl) : 1, e->treeletRounds)) return e->fail(BPT_ERR_CUDA, berr);
        }
        // (the top-level root now carries the tight bounds; the scene bounds stay the reference's)
        e->bvh.bmin[0] = looseScene.pMin.x; e->bvh.bmin[1] = looseScene.pMin.y; e->bvh.bmin[2] = looseScene.pMin.z;
        e->bvh.bmax[0] = looseScene.pMax.x; e->bvh.bmax[1] = looseScene.pMax.y; e->bvh.bmax[2] = looseScene.pMax.z;
    }
    ECK(cudaStreamSynchronize(e->stream));
    auto tb1 = std::chrono::steady_clock::now();
    e->nTrisTotal = (int)totalTris;
    // infinite lights: Preprocess(sceneBounds) -> bounding sphere of the aggregate bounds (lights.h: UniformInfiniteLight::Preprocess)
    {
        Pt3 bmn(e->bvh.bmin[0], e->bvh.bmin[1], e->bvh.bmin[2]), bmx(e->bvh.bmax[0], e->bvh.bmax[1], e->bvh.bmax[2]);
        Pt3 center = (bmn + bmx) / 2;
        bool inside = center.x >= bmn.x && center.x <= bmx.x && center.y >= bmn.y && center.y <= bmx.y && center.z >= bmn.z && center.z <= bmx.z;
        float radius = inside ? Distance(center, bmx) : 0;
        bool any = false;
        for (DeviceLight &dl : dlights) if (dl.type != LIGHT_DIFFUSE_AREA) { dl.sceneCenter = center; dl.sceneRadius = radius; any = true; }
        if (any) ECK(e->dLights.upload(dlights));
    }
    e->imgBufs.clear();
    for (auto &it : imgTables) {
        std::vector<float> fb;
        size_t oRGB = 0; fb.insert(fb.end(), it.rgb.begin(), it.rgb.end());
        auto put = [&](const std::vector<float> &v) { size_t o = fb.size(); fb.insert(fb.end(), v.begin(), v.end()); return o; };
        size_t o1 = put(it.dist.func), o2 = put(it.dist.cdf), o3 = put(it.dist.funcInt), o4 = put(it.dist.mfunc), o5 = put(it.dist.mcdf);
        size_t c1 = put(it.comp.func), c2 = put(it.comp.cdf), c3 = put(it.comp.funcInt), c4 = put(it.comp.mfunc), c5 = put(it.comp.mcdf);
        e->imgBufs.emplace_back(new DevBuf<float>());
        ECK(e->imgBufs.back()->upload(fb));
        const float *b = e->imgBufs.back()->p;
        DeviceImageLight di{};
        di.rgb = b + oRGB; di.res = it.res;
        auto view = [&](const FilterTables &t, size_t a1, size_t a2, size_t a3, size_t a4, size_t a5) {
            PC2DView v{};
            v.func = b + a1; v.cdf = b + a2; v.funcInt = b + a3; v.mfunc = b + a4; v.mcdf = b + a5; v.mfuncInt = t.mfuncInt;
            v.nu = it.res; v.nv = it.res; v.min0 = 0; v.max0 = 1; v.min1 = 0; v.max1 = 1;
            return v;
        };
        di.distrib = view(it.dist, o1, o2, o3, o4, o5);
        di.compensatedDistrib = view(it.comp, c1, c2, c3, c4, c5);
        dimg.push_back(di);
    }
    ECK(e->dImageLights.upload(dimg));
    // ---- image textures: MIP pyramids (Image::GeneratePyramid restated in texture_host.h), decoded to float
    e->haveTextures = !e->textures.empty();
    if (e->haveTextures) {
        if (!e->tables.f32.count("srgb8_to_linear") || !e->tables.f32.count("mip_filter_lut"))
            return e->fail(BPT_ERR_IO, "tables blob lacks the texture tables (regenerate it with oracle/_ref/bin/ref_dump tables)");
        texhost::Pyramid pyr{e->tables.F("srgb8_to_linear").data()};
        std::vector<DeviceTexture> dtex(e->textures.size());
        std::vector<std::vector<HostMIPLevel>> pyramids(e->textures.size());
        size_t nTexels = 0, nLevels = 0;
        std::vector<int> depth(e->textures.size(), 0);
        for (size_t i = 0; i < e->textures.size(); ++i) {
            const bpt_texture &tb = e->textures[i].t;
            if (tb.kind == BPT_TEXTURE_CONSTANT) {
                DeviceTexture &d = dtex[i];
                std::memset(&d, 0, sizeof(d));
                d.kind = TEXKIND_CONSTANT; d.cval = tb.value_f;
                if (tb.is_spectrum) {
                    // SpectrumConstantTexture::Create: GetOneSpectrum("value", ConstantSpectrum(1), spectrumType); every spectrum-valued
                    // parameter textures feed here is a reflectance or transmittance (SpectrumType::Albedo)
                    if (e->textures[i].value.s.kind == BPT_SPEC_NONE) { d.sval.kind = SPEC_CONSTANT; d.sval.c0 = 1.f; }
                    else if (!MakeSpectrumRec(e, e->textures[i].value, ST_ALBEDO, spool, &d.sval, &serr)) return e->fail(BPT_ERR_UNSUPPORTED, serr);
                }
                continue;
            }
            if (tb.kind == BPT_TEXTURE_SCALE || tb.kind == BPT_TEXTURE_MIX || tb.kind == BPT_TEXTURE_CHECKERBOARD) {
                auto isFloat = [&](int id) { return e->textures[id].t.is_spectrum == 0 && e->textures[id].t.kind != BPT_TEXTURE_NORMALMAP; };
                auto sameType = [&](int id) { return (e->textures[id].t.is_spectrum != 0) == (tb.is_spectrum != 0) && e->textures[id].t.kind != BPT_TEXTURE_NORMALMAP; };
                DeviceTexture &d = dtex[i];
                if (tb.kind == BPT_TEXTURE_CHECKERBOARD) {   // tex[0] where the weight is 0, tex[1] where it is 1 (textures.h:372-415)
                    int t1 = tb.tex1 - 1, t2 = tb.tex2 - 1;
                    if (!sameType(t1) || !sameType(t2)) return e->fail(BPT_ERR_INVALID, "checkerboard texture: value types of tex1 / tex2 do not fit");
                    std::memset(&d, 0, sizeof(d));
                    d.kind = TEXKIND_CHECKERBOARD; d.child0 = t1; d.child1 = t2; d.child2 = -1;
                    d.su = tb.su; d.sv = tb.sv; d.du = tb.du; d.dv = tb.dv;
                    SetTextureMapping(&d, tb, space, spaces);
                    depth[i] = 1 + std::max(depth[t1], depth[t2]);
                } else if (tb.kind == BPT_TEXTURE_MIX) {
                    int t1 = tb.tex1 - 1, t2 = tb.tex2 - 1, am = tb.tex3 - 1;
                    if (!sameType(t1) || !sameType(t2) || !isFloat(am)) return e->fail(BPT_ERR_INVALID, "mix texture: value types of tex1 / tex2 / amount do not fit");
                    std::memset(&d, 0, sizeof(d));
                    d.kind = TEXKIND_MIX; d.child0 = t1; d.child1 = t2; d.child2 = am;
                    depth[i] = 1 + std::max(depth[t1], std::max(depth[t2], depth[am]));
                } else {
                    // FloatScaledTexture::Create / SpectrumScaledTexture::Create (textures.cpp:892-957): a constant scale of 1 returns "tex"
                    // itself, a constant scale on an image texture is folded into the image texture's own scale
                    int tex = tb.tex1 - 1, scale = tb.tex2 - 1;
                    if (!sameType(tex) || !isFloat(scale)) return e->fail(BPT_ERR_INVALID, "scale texture: value types of tex / scale do not fit");
                    bool done = false;
                    for (int pass = 0; pass < (tb.is_spectrum ? 1 : 2) && !done; ++pass) {
                        if (dtex[scale].kind == TEXKIND_CONSTANT) {
                            float cs = dtex[scale].cval;
                            if (cs == 1) { d = dtex[tex]; depth[i] = depth[tex]; done = true; }
                            else if (dtex[tex].kind == TEXKIND_IMAGEMAP) { d = dtex[tex]; d.scale *= cs; depth[i] = 0; done = true; }
                        }
                        if (!done && !tb.is_spectrum) std::swap(tex, scale);
                    }
                    if (!done) {
                        if (!tb.is_spectrum) std::swap(tex, scale);   // the reference's final swap: FloatScaledTexture(scale, tex)
                        std::memset(&d, 0, sizeof(d));
                        d.kind = TEXKIND_SCALE; d.child0 = tex; d.child1 = scale;
                        depth[i] = 1 + std::max(depth[tex], depth[scale]);
                    }
                }
                if (depth[i] > kTexMaxDepth) return e->fail(BPT_ERR_UNSUPPORTED, "texture graph deeper than the engine evaluates (3 levels of scale / mix)");
                continue;
            }
            HostImage im = e->textures[i].img;
            if (e->textures[i].t.kind == BPT_TEXTURE_NORMALMAP) {
                // the image itself, RGB channels (BasicScene::startLoadingNormalMaps, scene.cpp:886-910), decoded as Image::GetChannel does
                if (im.nc < 3) return e->fail(BPT_ERR_INVALID, "normal map image must contain R, G, and B channels");
                HostMIPLevel lv; lv.xres = im.xres; lv.yres = im.yres; lv.texels.resize((size_t)im.xres * im.yres * 3);
                for (size_t px = 0; px < (size_t)im.xres * im.yres; ++px)
                    for (int c = 0; c < 3; ++c) lv.texels[3 * px + c] = im.is8 ? pyr.Decode(im, im.p8[px * im.nc + c]) : im.p32[px * im.nc + c];
                im.nc = 3;
                pyramids[i].push_back(std::move(lv));
            } else {
                if (e->textures[i].t.wrap == BPT_TEXWRAP_BLACK && ((im.xres & (im.xres - 1)) || (im.yres & (im.yres - 1))))
                    return e->fail(BPT_ERR_UNSUPPORTED, "wrap mode \"black\" on an image whose sides are not powers of two (the reference aborts on it, util/image.cpp:114)");
                DropOpaqueAlpha(&im);
                pyramids[i] = pyr.Generate(im, e->textures[i].t.wrap);
            }
            const bpt_texture &b = e->textures[i].t;
            DeviceTexture &d = dtex[i];
            d.firstLevel = (int)nLevels; d.nLevels = (int)pyramids[i].size(); d.nc = im.nc;
            d.wrap = b.wrap; d.filter = b.filter; d.maxAnisotropy = b.max_anisotropy;
            d.su = b.su; d.sv = b.sv; d.du = b.du; d.dv = b.dv; d.scale = b.scale; d.invert = b.invert;
            SetTextureMapping(&d, b, space, spaces);
            d.kind = TEXKIND_IMAGEMAP; d.child0 = d.child1 = d.child2 = -1; d.cval = 0; std::memset(&d.sval, 0, sizeof(d.sval));
            for (const HostMIPLevel &lv : pyramids[i]) nTexels += lv.texels.size();
            nLevels += pyramids[i].size();
        }
        ECK(e->dTexels.alloc(nTexels));
        std::vector<DeviceMIPLevel> dlev;
        std::vector<float> all; all.reserve(nTexels);
        for (auto &py : pyramids)
            for (const HostMIPLevel &lv : py) {
                dlev.push_back(DeviceMIPLevel{e->dTexels.p + all.size(), lv.xres, lv.yres});
                all.insert(all.end(), lv.texels.begin(), lv.texels.end());
            }
        ECK(cudaMemcpy(e->dTexels.p, all.data(), all.size() * sizeof(float), cudaMemcpyHostToDevice));
        ECK(e->dMipLevels.upload(dlev));
        ECK(e->dTextures.upload(dtex));
        ECK(e->dEwaLUT.upload(e->tables.F("mip_filter_lut")));
        ECK(e->specPool.upload(spool.pool));   // constant spectrum textures added their tables
    }
    if ((!dimg.empty() || e->haveTextures) && e->dRgb2spec.n == 0) {
        std::vector<float> t(e->tables.F("rgb2spec_srgb_scale"));
        t.insert(t.end(), e->tables.F("rgb2spec_srgb_coeffs").begin(), e->tables.F("rgb2spec_srgb_coeffs").end());
        ECK(e->dRgb2spec.upload(t));
    }

    lap("BVH build");
    // ---- path state and queues
    int nPixPass = (pb[1] - pb[0]) * (pb[3] - pb[2]);
    // pool size: enough for every sample of the frame in one pass, capped by BPT_MAX_PATHS (default 32 Mi paths, ~9 GB)
    long long wantPaths = (long long)nPixPass * std::max(1, e->sampler.spp);
    int maxPaths = (int)std::max<long long>(nPixPass, std::min<long long>(e->maxPaths, wantPaths));
    // one pass covers the whole pixel bounds for S samples; S is chosen at render time with S * nPix <= maxPaths
    e->rc = RenderCtx{};
    RenderCtx &rc = e->rc;
    rc.maxPaths = maxPaths;
    ECK(e->stateF4.alloc((size_t)maxPaths * 12));
    ECK(e->stateMisc.alloc(maxPaths));
    ECK(e->stateAux.alloc(maxPaths));
    float4 *f4 = e->stateF4.p;
    rc.ps.rayO = f4; rc.ps.rayD = f4 + (size_t)maxPaths; rc.ps.hit = f4 + 2 * (size_t)maxPaths; rc.ps.beta = f4 + 3 * (size_t)maxPaths;
    rc.ps.r_u = f4 + 4 * (size_t)maxPaths; rc.ps.r_l = f4 + 5 * (size_t)maxPaths; rc.ps.L = f4 + 6 * (size_t)maxPaths;
    rc.ps.lambda = f4 + 7 * (size_t)maxPaths; rc.ps.lambdaPdf = f4 + 8 * (size_t)maxPaths; rc.ps.ctx0 = f4 + 9 * (size_t)maxPaths;
    rc.ps.ctx1 = f4 + 10 * (size_t)maxPaths; rc.ps.ctx2 = f4 + 11 * (size_t)maxPaths;
    rc.ps.misc = e->stateMisc.p; rc.ps.aux = e->stateAux.p;
    if (df.isGBuffer) {
        ECK(e->visSurf.alloc((size_t)maxPaths * 6));
        ECK(e->gbuf.alloc((size_t)fl.xres * fl.yres));
        ECK(cudaMemset(e->gbuf.p, 0, sizeof(GBufferPixel) * (size_t)fl.xres * fl.yres));
        ECK(e->gbufOut.alloc(25 * (size_t)fl.xres * fl.yres));
    }
    rc.ps.vs = df.isGBuffer ? e->visSurf.p : nullptr;
    rc.gbuf = df.isGBuffer ? e->gbuf.p : nullptr;
    int nQueues = 0;
    int qIndex[Q_CURSOR_CLOSEST];
    for (int q = 0; q < Q_CURSOR_CLOSEST; ++q) {
        bool need = q == Q_RAY0 || q == Q_RAY1 || q == Q_ESCAPED || q == Q_HITLIGHT || ((q == Q_MEDIUM || q == Q_MSCATTER) && !e->media.empty()) ||
                    (q >= Q_MAT0 && e->materialPresent[q - Q_MAT0]);
        qIndex[q] = need ? nQueues++ : -1;
    }
    ECK(e->queueMem.alloc((size_t)nQueues * maxPaths));
    for (int q = 0; q < Q_CURSOR_CLOSEST; ++q) rc.queues[q] = qIndex[q] >= 0 ? e->queueMem.p + (size_t)qIndex[q] * maxPaths : nullptr;
    rc.nMedia = (int)e->media.size();
    rc.media = e->dMedia.p;
    rc.cameraMedium = e->cameraMedium;
    rc.useTr = (rc.nMedia > 0 || e->haveInterface) ? 1 : 0;
    ECK(e->shadowItems.alloc(rc.useTr ? 1 : maxPaths));
    ECK(e->shadowTr.alloc(rc.useTr ? maxPaths : 1));
    rc.shadowItems = e->shadowItems.p;
    rc.shadowTr = e->shadowTr.p;
    // SkipIntersection advances the ray differentials by tHit (read by GBufferFilm's dzdx / dzdy and by texture filtering)
    rc.storeHitT = ((df.isGBuffer || e->haveTextures) && e->haveInterface) ? 1 : 0;
    ECK(e->rayDiffs.alloc(e->haveTextures ? (size_t)maxPaths * 4 : 1));
    rc.ps.rdiff = e->haveTextures ? e->rayDiffs.p : nullptr;
    ECK(e->hitT.alloc((rc.nMedia > 0 || rc.storeHitT) ? maxPaths : 1));
    ECK(e->pathMedium.alloc(rc.nMedia > 0 ? maxPaths : 1));
    rc.ps.hitT = e->hitT.p; rc.ps.medium = e->pathMedium.p;
    ECK(e->counts.alloc(Q_NCOUNTERS));
    ECK(cudaMemset(e->counts.p, 0, sizeof(int) * Q_NCOUNTERS));
    rc.counts = e->counts.p;
    ECK(e->stats.alloc(16));
    ECK(cudaMemset(e->stats.p, 0, sizeof(unsigned long long) * 16));
    rc.stats = e->stats.p;
    ECK(e->rawCursor.alloc(1));
    {
        const char *rs = std::getenv("BPT_RAY_SORT");
        e->raySort = rs && std::atoi(rs) != 0;
        if (e->raySort) {
            ECK(e->sortKeys.alloc(maxPaths)); ECK(e->sortKeysTmp.alloc(maxPaths)); ECK(e->sortQueueTmp.alloc(maxPaths));
            ECK(e->sortHist.alloc(256 * (size_t)((maxPaths + 4095) / 4096)));
        }
    }
    rc.bvh.nodes = e->bvh.nodes; rc.bvh.tris = e->bvh.tris;
    rc.bvh.instances = instanced ? e->dInstances.p : nullptr;
    rc.bvh.meshAlpha = e->dMeshAlpha.p;
    rc.bvh.shapes = e->dShapes.p; rc.bvh.patches = e->dPatches.p; rc.bvh.meshShape = e->dMeshShape.p;
    ECK(e->hitInst.alloc(instanced ? maxPaths : 1));
    rc.ps.hitInst = e->hitInst.p;
    SceneView &sv = rc.sv;
    sv.triPool = e->triPool.p; sv.meshes = e->dMeshes.p; sv.indexPool = e->indexPool.p; sv.nPool = e->nPool.p; sv.uvPool = e->uvPool.p;
    sv.materials = e->dMaterials.p; sv.lights = e->dLights.p; sv.specPool = e->specPool.p;
    dls.nodes = e->dLightNodes.p; dls.infiniteLights = e->dInfinite.p; dls.bitTrails = e->dBitTrails.p; dls.lightInBVH = e->dLightInBVH.p;
    sv.ls = dls;
    sv.allbMin = lb.allLightBounds.pMin; sv.allbMax = lb.allLightBounds.pMax;
    sv.imageLights = e->dImageLights.p;
    sv.instances = instanced ? e->dInstances.p : nullptr;
    sv.cieX = e->dTables.p; sv.cieY = e->dTables.p + kNLambda; sv.cieZ = e->dTables.p + 2 * kNLambda; sv.illumD65 = e->dTables.p + 3 * kNLambda;
    sv.textures = e->haveTextures ? e->dTextures.p : nullptr; sv.mipLevels = e->haveTextures ? e->dMipLevels.p : nullptr;
    sv.ewaLUT = e->haveTextures ? e->dEwaLUT.p : nullptr;
    sv.shapes = e->dShapes.p; sv.patches = e->dPatches.p;
    sv.rgb2specScale = e->dRgb2spec.n ? e->dRgb2spec.p : nullptr; sv.rgb2specCoeffs = e->dRgb2spec.n ? e->dRgb2spec.p + 64 : nullptr;
    sv.nLights = (int)dlights.size();
    // ZSobol digit-prefix table: [dimension][pixel] uint32 (only when the permuted index fits in 32 bits)
    sp.prefix = nullptr; sp.prefixDims = 0; sp.prefixW = fl.xres; sp.prefixH = fl.yres;
    sp.nLowDigits = (sp.log2SamplesPerPixel + 1) / 2;
    {
        const char *pd = std::getenv("BPT_ZSOBOL_PREFIX");
        bool enable = !(pd && std::atoi(pd) == 0);
        int dims = std::min(6 + 7 * (e->integ.max_depth + 1), 64);
        size_t entries = (size_t)dims * fl.xres * fl.yres;
        if (enable && sp.type == 0 && 2 * sp.nBase4Digits <= 32 && sp.nBase4Digits > sp.nLowDigits && entries * 4 <= (size_t)2 << 30) {
            // the table depends on (film size, digit counts, dimensions) only -- not on the seed, the scene or the sample index: a
            // re-commit with the same film and sampler keeps it (it is 200 MB at 1024^2 and took 1/3 of a re-commit)
            const long long key[5] = {fl.xres, fl.yres, dims, sp.nBase4Digits, sp.log2SamplesPerPixel};
            const bool same = e->zsPrefix.p && e->zsPrefix.n == entries && std::memcmp(key, e->zsKey, sizeof(key)) == 0;
            ECK(e->zsPrefix.alloc(entries));
            sp.prefixDims = dims;
            if (!same) LaunchZSobolPrefix(sp, e->zsPrefix.p, e->smCount, e->stream);
            std::memcpy(e->zsKey, key, sizeof(key));
            sp.prefix = e->zsPrefix.p;
        }
    }
    rc.camera = dc; rc.film = df; rc.filter = dflt; rc.sampler = sp;
    rc.integ.maxDepth = e->integ.max_depth; rc.integ.regularize = e->integ.regularize;
    rc.nInfinite = dls.nInfinite;
    ECK(e->filmPixels.alloc((size_t)fl.xres * fl.yres));
    ECK(cudaMemset(e->filmPixels.p, 0, sizeof(double4) * (size_t)fl.xres * fl.yres));
    rc.filmPixels = e->filmPixels.p;
    ECK(e->filmRGB.alloc(3 * (size_t)fl.xres * fl.yres));
    ECK(cudaDeviceSynchronize());

    lap("path state / film allocation");
    auto t1 = std::chrono::steady_clock::now();
    e->st = bpt_stats{};
    e->st.seconds_build = std::chrono::duration<double>(tb1 - tb0).count();
    e->st.bvh_nodes = e->bvh.nNodes; e->st.bvh_triangles = e->bvh.nTris; e->st.bvh_depth = e->bvh.depth;
    e->st.bytes_bvh_nodes = (uint64_t)e->bvh.nNodes * 80; e->st.bytes_bvh_tris = (uint64_t)e->bvh.nTris * 48;
    (void)t0; (void)t1;
    e->committed = true;
    e->committedSpp = e->sampler.spp;
    return BPT_OK;
}

// ------------------------------------------------------------------------------------------- rendering
// Waits for an asynchronous render and books its device time.
static int Settle(bpt_engine *e) {
    if (!e || !e->pending) return BPT_OK;
    ECK(cudaSetDevice(e->device));
    ECK(cudaEventSynchronize(e->ev1));
    ECK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->st.seconds_render += ms * 1e-3;
    e->st.kernel_launches += e->pendingLaunches;
    e->pending = false; e->pendingLaunches = 0;
    return BPT_OK;
}
extern "C" int bpt_sync(bpt_engine *e) { return e ? Settle(e) : BPT_ERR_INVALID; }
extern "C" int bpt_sampler_spp(const bpt_engine *e) { return e && e->committed ? e->committedSpp : BPT_ERR_INVALID; }

static int RenderImpl(bpt_engine *e, int first_sample, int n_samples, const int *pixel_bounds, int flags, int band_rows, int n_parts, int part,
                      cudaStream_t userStream = nullptr) {
    if (!e) return BPT_ERR_INVALID;
    if (!e->committed) return e->fail(BPT_ERR_INVALID, "scene not committed");
    if (n_samples <= 0 || first_sample < 0) return e->fail(BPT_ERR_INVALID, "bad sample range");
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    RenderCtx &rc = e->rc;
    int x0 = rc.film.pbx0, x1 = rc.film.pbx1, y0 = rc.film.pby0, y1 = rc.film.pby1;
    if (pixel_bounds) {
        x0 = std::max(x0, pixel_bounds[0]); x1 = std::min(x1, pixel_bounds[1]);
        y0 = std::max(y0, pixel_bounds[2]); y1 = std::min(y1, pixel_bounds[3]);
    }
    rc.bandRows = 0; rc.bandParts = 1; rc.bandPart = 0; rc.bandY0 = 0;
    if (band_rows > 0) {
        // interleaved bands of the film's pixel bounds: this call covers the local rows [0, nMine) of part `part`
        if (n_parts <= 0 || part < 0 || part >= n_parts) return e->fail(BPT_ERR_INVALID, "bad band partition");
        int nMine = 0;
        for (int k = 0, y = y0; y < y1; ++k, y += band_rows) if (k % n_parts == part) nMine += std::min(band_rows, y1 - y);
        rc.bandRows = band_rows; rc.bandParts = n_parts; rc.bandPart = part; rc.bandY0 = y0;
        y0 = 0; y1 = nMine;   // local rows from here on
    }
    if (x0 >= x1 || y0 >= y1) return BPT_OK;   // empty tile
    const bool count = flags & 1;
    const bool profile = flags & 2;
    cudaStream_t st = userStream ? userStream : e->stream;
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> prof;   // (class, (start, stop))
    auto profBegin = [&](int cls) {
        if (!profile) return;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, st); prof.push_back({cls, {a, b}});
    };
    auto profEnd = [&]() { if (profile) cudaEventRecord(prof.back().second.second, st); };
    const int sms = e->smCount;
    int launches = 0;
    // kernel set: wavefront.cu (triangles only) or wavefront_shapes.cu (quadric leaves, interactions, area lights compiled in)
    const bool useShapes = e->haveShapes;
#define KSEL(f) (useShapes ? &tu_shapes::f : &f)
    KSEL(UploadSamplerConstants)(rc.sampler, st);
    if (e->haveTextures) UploadSamplerConstantsTex(rc.sampler, st);
    ECK(cudaEventRecord(e->ev0, st));
    // scanline bands so that a band of S samples fits in the path-state arrays
    int width = x1 - x0;
    int rowsPerBand = std::max(1, std::min(y1 - y0, rc.maxPaths / width));
    if (width > rc.maxPaths) return e->fail(BPT_ERR_INVALID, "image row exceeds path-state capacity");
    for (int by0 = y0; by0 < y1; by0 += rowsPerBand) {
        int by1 = std::min(y1, by0 + rowsPerBand);
        int nPix = width * (by1 - by0);
        int S = std::max(1, std::min(n_samples, rc.maxPaths / nPix));
        for (int s0 = 0; s0 < n_samples; s0 += S) {
            int ns = std::min(S, n_samples - s0);
            KSEL(LaunchResetAll)(rc, st); ++launches;
            KSEL(LaunchGenerateCameraRays)(rc, x0, x1, by0, by1, first_sample + s0, ns, sms, st); ++launches;
            int cur = Q_RAY0, next = Q_RAY1;
            int maxDepth = rc.integ.maxDepth;
            // Scenes with media or interface surfaces need one wavefront iteration per path *segment* (null collisions and
            // interface crossings do not advance depth): iterate until the ray queue is empty.  Otherwise depth == iteration.
            const bool dynamic = rc.useTr != 0;
            for (int depth = 0; dynamic ? depth < 100000 : depth <= maxDepth; ++depth) {
                KSEL(LaunchResetDepth)(rc, next, st); ++launches;
                profBegin(0); KSEL(LaunchTraceClosest)(rc, cur, count, sms, st); ++launches; profEnd();
                if (rc.nMedia > 0) { profBegin(4); KSEL(LaunchSampleMedium)(rc, next, sms, st); launches += 2; profEnd(); }
                if (rc.nInfinite > 0) { KSEL(LaunchHandleEscaped)(rc, sms, st); ++launches; }
                if (e->haveAreaLights) { KSEL(LaunchHandleEmissive)(rc, sms, st); ++launches; }
                // (GBufferFilm with maxdepth 0: the visible surface is still initialised at the first hit, cpu/integrators.cpp:1118-1145,
                // before Li() returns at the depth test -- the shading kernels do both, so they run once more)
                if (!dynamic && depth == maxDepth && !(rc.film.isGBuffer && depth == 0)) break;
                profBegin(2);
                for (int m = 0; m < MAT_NTAGS; ++m)
                    if (e->materialPresent[m]) { if (e->haveTextures) LaunchShadeTex(rc, m, next, sms, st); else KSEL(LaunchShade)(rc, m, next, sms, st); ++launches; }
                profEnd();
                profBegin(1);
                if (rc.useTr) KSEL(LaunchTraceShadowTr)(rc, sms, st); else KSEL(LaunchTraceShadow)(rc, count, sms, st);
                ++launches; profEnd();
                std::swap(cur, next);
                if (dynamic || e->raySort) {
                    int live = 0;
                    ECK(cudaMemcpyAsync(&live, rc.counts + cur, sizeof(int), cudaMemcpyDeviceToHost, st));
                    ECK(cudaStreamSynchronize(st));
                    if (live == 0) break;
                    if (e->raySort && !(!dynamic && depth + 1 > maxDepth)) {
                        profBegin(3);
                        bool flipped = KSEL(LaunchRaySort)(rc, cur, live, e->bvh.bmin, e->bvh.bmax, e->sortKeys.p, e->sortKeysTmp.p, e->sortQueueTmp.p,
                                                     e->sortHist.p, st);
                        launches += 4;
                        if (flipped) ECK(cudaMemcpyAsync(rc.queues[cur], e->sortQueueTmp.p, sizeof(int) * (size_t)live, cudaMemcpyDeviceToDevice, st));
                        profEnd();
                    }
                }
            }
            KSEL(LaunchUpdateFilm)(rc, x0, x1, by0, by1, ns, sms, st); ++launches;
#undef KSEL
        }
    }
    ECK(cudaEventRecord(e->ev1, st));
    if ((flags & 4) && !profile && rc.useTr == 0 && !e->raySort) {   // BPT_RENDER_ASYNC: no host round trip inside the loop above
        e->pending = true; e->pendingLaunches = launches;
        return BPT_OK;
    }
    ECK(cudaStreamSynchronize(st));
    ECK(cudaGetLastError());
    float ms = 0;
    cudaEventElapsedTime(&ms, e->ev0, e->ev1);
    e->st.seconds_render += ms * 1e-3;
    e->st.kernel_launches += launches;
    for (auto &pe : prof) {
        float pm = 0;
        cudaEventElapsedTime(&pm, pe.second.first, pe.second.second);
        if (pe.first == 0) { e->st.ms_trace_closest += pm; e->st.launches_trace_closest++; }
        else if (pe.first == 1) { e->st.ms_trace_shadow += pm; e->st.launches_trace_shadow++; }
        else if (pe.first == 3) e->st.ms_other += pm;
        else if (pe.first == 4) e->st.ms_medium += pm;
        else e->st.ms_shade += pm;
        cudaEventDestroy(pe.second.first); cudaEventDestroy(pe.second.second);
    }
    return BPT_OK;
}

extern "C" int bpt_render(bpt_engine *e, int first_sample, int n_samples, const int *pixel_bounds, int flags) {
    return RenderImpl(e, first_sample, n_samples, pixel_bounds, flags, 0, 1, 0);
}
extern "C" int bpt_render_bands(bpt_engine *e, int first_sample, int n_samples, int band_rows, int n_parts, int part, int flags) {
    if (band_rows <= 0) return e ? e->fail(BPT_ERR_INVALID, "band_rows must be positive") : BPT_ERR_INVALID;
    return RenderImpl(e, first_sample, n_samples, nullptr, flags, band_rows, n_parts, part);
}
extern "C" int bpt_render_stream(bpt_engine *e, int first_sample, int n_samples, const int *pixel_bounds, int flags, void *cuda_stream) {
    return RenderImpl(e, first_sample, n_samples, pixel_bounds, flags, 0, 1, 0, (cudaStream_t)cuda_stream);
}
extern "C" int bpt_render_bands_stream(bpt_engine *e, int first_sample, int n_samples, int band_rows, int n_parts, int part, int flags, void *cuda_stream) {
    if (band_rows <= 0) return e ? e->fail(BPT_ERR_INVALID, "band_rows must be positive") : BPT_ERR_INVALID;
    return RenderImpl(e, first_sample, n_samples, nullptr, flags, band_rows, n_parts, part, (cudaStream_t)cuda_stream);
}

extern "C" int bpt_film_clear(bpt_engine *e) {
    if (!e || !e->committed) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    ECK(cudaMemset(e->filmPixels.p, 0, sizeof(double4) * e->filmPixels.n));
    if (e->rc.gbuf) ECK(cudaMemset(e->gbuf.p, 0, sizeof(GBufferPixel) * e->gbuf.n));
    return BPT_OK;
}
extern "C" int bpt_film_resolve_gbuffer(bpt_engine *e, float *out25) {
    if (!e || !e->committed || !out25) return BPT_ERR_INVALID;
    if (!e->rc.gbuf) return e->fail(BPT_ERR_INVALID, "the film is not a GBufferFilm");
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    LaunchResolveGBuffer(e->filmPixels.p, e->gbuf.p, e->rc.film, e->gbufOut.p, e->smCount, e->stream);
    ECK(cudaMemcpyAsync(out25, e->gbufOut.p, sizeof(float) * e->gbufOut.n, cudaMemcpyDeviceToHost, e->stream));
    ECK(cudaStreamSynchronize(e->stream));
    return BPT_OK;
}
extern "C" int bpt_stats_reset(bpt_engine *e) {
    if (!e || !e->committed) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    ECK(cudaMemset(e->stats.p, 0, sizeof(unsigned long long) * 16));
    e->st.seconds_render = 0; e->st.kernel_launches = 0;
    e->st.ms_trace_closest = e->st.ms_trace_shadow = e->st.ms_shade = e->st.ms_other = e->st.ms_medium = 0;
    e->st.launches_trace_closest = e->st.launches_trace_shadow = 0;
    return BPT_OK;
}
extern "C" int bpt_film_read(bpt_engine *e, double *out) {
    if (!e || !e->committed || !out) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    ECK(cudaMemcpy(out, e->filmPixels.p, sizeof(double4) * e->filmPixels.n, cudaMemcpyDeviceToHost));
    return BPT_OK;
}
extern "C" int bpt_film_write(bpt_engine *e, const double *in) {
    if (!e || !e->committed || !in) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    ECK(cudaMemcpy(e->filmPixels.p, in, sizeof(double4) * e->filmPixels.n, cudaMemcpyHostToDevice));
    return BPT_OK;
}
extern "C" int bpt_film_copy_to_device(bpt_engine *e, void *dst) {
    if (!e || !e->committed || !dst) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    ECK(cudaMemcpy(dst, e->filmPixels.p, sizeof(double4) * e->filmPixels.n, cudaMemcpyDeviceToDevice));
    return BPT_OK;
}
extern "C" int bpt_film_copy_from_device(bpt_engine *e, const void *src) {
    if (!e || !e->committed || !src) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    ECK(cudaMemcpy(e->filmPixels.p, src, sizeof(double4) * e->filmPixels.n, cudaMemcpyDeviceToDevice));
    return BPT_OK;
}
extern "C" int bpt_film_resolve(bpt_engine *e, float *rgb) {
    if (!e || !e->committed || !rgb) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    LaunchResolveFilm(e->filmPixels.p, e->rc.film, e->filmRGB.p, e->smCount, e->stream);
    ECK(cudaStreamSynchronize(e->stream));
    ECK(cudaMemcpy(rgb, e->filmRGB.p, sizeof(float) * e->filmRGB.n, cudaMemcpyDeviceToHost));
    return BPT_OK;
}
// Host-only entry point of the writer (CPU-testable, no engine needed)
extern "C" int bpt_write_exr_file(const char *path, int width, int height, int n_channels, const char *const *names, const float *interleaved) {
    if (!path || !names) return BPT_ERR_INVALID;
    std::vector<std::string> nm(names, names + std::max(0, n_channels));
    std::string err;
    return WriteEXR(path, width, height, nm, interleaved, {}, &err) ? BPT_OK : BPT_ERR_IO;
}
// RGBFilm::WriteImage / GBufferFilm::WriteImage (film.cpp:533-565, 688-803): the resolved film as an OpenEXR file -- channels R G B, or
// the 25 GBuffer channels under the reference's names -- with the reference's pixelBounds / fullResolution / samplesPerPixel metadata
extern "C" int bpt_film_write_exr(bpt_engine *e, const char *path) {
    if (!e || !e->committed || !path) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
    const int w = e->film.xres, h = e->film.yres;
    static const char *gb[25] = {"R", "G", "B", "Albedo.R", "Albedo.G", "Albedo.B", "P.X", "P.Y", "P.Z", "dzdx", "dzdy", "N.X", "N.Y", "N.Z", "Ns.X", "Ns.Y",
                                 "Ns.Z", "u", "v", "Variance.R", "Variance.G", "Variance.B", "RelativeVariance.R", "RelativeVariance.G", "RelativeVariance.B"};
    std::vector<std::string> names;
    std::vector<float> px;
    int rc;
    if (e->rc.film.isGBuffer) { names.assign(gb, gb + 25); px.resize(25 * (size_t)w * h); rc = bpt_film_resolve_gbuffer(e, px.data()); }
    else { names = {"R", "G", "B"}; px.resize(3 * (size_t)w * h); rc = bpt_film_resolve(e, px.data()); }
    if (rc != BPT_OK) return rc;
    std::vector<ExrAttr> extra;
    auto bytes = [](const void *p, size_t n) { return std::vector<unsigned char>((const unsigned char *)p, (const unsigned char *)p + n); };
    { int32_t b[4] = {e->rc.film.pbx0, e->rc.film.pby0, e->rc.film.pbx1 - 1, e->rc.film.pby1 - 1}; extra.push_back({"pixelBounds", "box2i", bytes(b, 16)}); }
    { int32_t r[2] = {w, h}; extra.push_back({"fullResolution", "v2i", bytes(r, 8)}); }
    { int32_t s = e->committedSpp; extra.push_back({"samplesPerPixel", "int", bytes(&s, 4)}); }
    std::string err;
    if (!WriteEXR(path, w, h, names, px.data(), extra, &err)) return e->fail(BPT_ERR_IO, err);
    return BPT_OK;
}
extern "C" int bpt_film_dims(const bpt_engine *e, int *xres, int *yres) {
    if (!e || !e->haveFilm) return BPT_ERR_INVALID;
    if (xres) *xres = e->film.xres;
    if (yres) *yres = e->film.yres;
    return BPT_OK;
}
extern "C" int bpt_get_stats(bpt_engine *e, bpt_stats *out) {
    if (!e || !out) return BPT_ERR_INVALID;
    if (e->committed) {
        unsigned long long h[16];
        ECK(cudaSetDevice(e->device));
        { int rcS = Settle(e); if (rcS != BPT_OK) return rcS; }
        ECK(cudaMemcpy(h, e->stats.p, sizeof(h), cudaMemcpyDeviceToHost));
        e->st.camera_rays = h[0];
        e->st.indirect_rays = h[1] >= h[0] ? h[1] - h[0] : 0;
        e->st.shadow_rays = h[2];
        e->st.nodes_visited = h[3]; e->st.tris_tested = h[4]; e->st.counted_rays = h[5];
        e->st.nodes_visited_shadow = h[6]; e->st.tris_tested_shadow = h[7]; e->st.counted_rays_shadow = h[8];
    }
    *out = e->st;
    return BPT_OK;
}

// ------------------------------------------------------------------------------------------- raw rays / probes
extern "C" int bpt_trace_closest_device(bpt_engine *e, const void *rays_dev, int n, void *hits_dev, int flags) {
    if (!e || !e->committed || !rays_dev || !hits_dev || n < 0) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    (e->haveShapes ? &tu_shapes::LaunchTraceRaw : &LaunchTraceRaw)(e->rc.bvh, (const float4 *)rays_dev, n, (float4 *)hits_dev, nullptr, e->rawCursor.p, e->stats.p, (flags & 2) != 0, (flags & 1) != 0,
                   e->smCount, e->stream);
    ECK(cudaStreamSynchronize(e->stream));
    return BPT_OK;
}
static int TraceHost(bpt_engine *e, const float *rays, int n, float *hits, float *tOut, int *occ, bool any) {
    if (!e || !e->committed || !rays || n < 0) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    DevBuf<float4> dr, dh; DevBuf<float> dt;
    ECK(dr.alloc(2 * (size_t)n)); ECK(dh.alloc(n)); ECK(dt.alloc(n));
    ECK(cudaMemcpy(dr.p, rays, sizeof(float) * 8 * (size_t)n, cudaMemcpyHostToDevice));
    (e->haveShapes ? &tu_shapes::LaunchTraceRaw : &LaunchTraceRaw)(e->rc.bvh, dr.p, n, dh.p, dt.p, e->rawCursor.p, e->stats.p, any, false, e->smCount, e->stream);
    ECK(cudaStreamSynchronize(e->stream));
    ECK(cudaGetLastError());
    std::vector<float> hh(4 * (size_t)n);
    ECK(cudaMemcpy(hh.data(), dh.p, sizeof(float) * 4 * (size_t)n, cudaMemcpyDeviceToHost));
    if (hits) std::memcpy(hits, hh.data(), sizeof(float) * 4 * (size_t)n);
    if (tOut) ECK(cudaMemcpy(tOut, dt.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    if (occ) for (int i = 0; i < n; ++i) { int prim; std::memcpy(&prim, &hh[4 * (size_t)i], 4); occ[i] = prim >= 0; }
    return BPT_OK;
}
extern "C" int bpt_trace_closest(bpt_engine *e, const float *rays, int n, float *hits, float *t_out) { return TraceHost(e, rays, n, hits, t_out, nullptr, false); }
extern "C" int bpt_trace_any(bpt_engine *e, const float *rays, int n, int *occluded) { return TraceHost(e, rays, n, nullptr, nullptr, occluded, true); }

extern "C" int bpt_probe_sampler(bpt_engine *e, const int *pxy, const int *sidx, const int *dim, int n, float *o1, float *o2) {
    if (!e || !e->committed || n <= 0) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    DevBuf<int> a, b, c; DevBuf<float> r1, r2;
    ECK(a.alloc(2 * (size_t)n)); ECK(b.alloc(n)); ECK(c.alloc(n)); ECK(r1.alloc(n)); ECK(r2.alloc(2 * (size_t)n));
    ECK(cudaMemcpy(a.p, pxy, sizeof(int) * 2 * (size_t)n, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(b.p, sidx, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(c.p, dim, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice));
    UploadSamplerConstants(e->rc.sampler, e->stream);
    if (e->haveTextures) UploadSamplerConstantsTex(e->rc.sampler, e->stream);
    LaunchProbeSampler(e->rc.sampler, a.p, b.p, c.p, n, r1.p, r2.p, e->stream);
    ECK(cudaStreamSynchronize(e->stream));
    ECK(cudaMemcpy(o1, r1.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    ECK(cudaMemcpy(o2, r2.p, sizeof(float) * 2 * (size_t)n, cudaMemcpyDeviceToHost));
    return BPT_OK;
}
extern "C" int bpt_probe_intersect_triangle(bpt_engine *e, const float *rays, const float *tris, int n, float *out) {
    if (!e || n <= 0) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    DevBuf<float> a, b, c;
    ECK(a.alloc(7 * (size_t)n)); ECK(b.alloc(9 * (size_t)n)); ECK(c.alloc(5 * (size_t)n));
    ECK(cudaMemcpy(a.p, rays, sizeof(float) * 7 * (size_t)n, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(b.p, tris, sizeof(float) * 9 * (size_t)n, cudaMemcpyHostToDevice));
    LaunchProbeTri(a.p, b.p, n, c.p, e->stream);
    ECK(cudaStreamSynchronize(e->stream));
    ECK(cudaMemcpy(out, c.p, sizeof(float) * 5 * (size_t)n, cudaMemcpyDeviceToHost));
    return BPT_OK;
}

extern "C" int bpt_probe_math(bpt_engine *e, int fn, const float *x, const float *y, int n, float *out) {
    if (!e || n <= 0) return BPT_ERR_INVALID;
    ECK(cudaSetDevice(e->device));
    DevBuf<float> a, b, c;
    ECK(a.alloc(n)); ECK(b.alloc(n)); ECK(c.alloc(n));
    ECK(cudaMemcpy(a.p, x, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
    ECK(cudaMemcpy(b.p, y, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice));
    LaunchProbeMath(fn, a.p, b.p, n, c.p, e->stream);
    ECK(cudaStreamSynchronize(e->stream));
    ECK(cudaMemcpy(out, c.p, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost));
    return BPT_OK;
}
