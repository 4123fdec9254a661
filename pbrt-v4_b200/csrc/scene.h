// scene.h -- POD scene records as they live in HBM (shared by host setup and kernels).
// Tag values follow the reference's TaggedPointer type lists (tag = position in list):
//   Material  base/material.h:36-39   Light  base/light.h:40-43   BxDF base/bxdf.h:167-169
//   Camera    base/camera.h:34        Film   base/film.h:25       Sampler base/sampler.h:40-42
#pragma once
#include <cuda_runtime.h>
#include "pmath.h"
#include "spectrum.h"

namespace bpt {

enum MaterialTag { MAT_COATED_DIFFUSE = 0, MAT_COATED_CONDUCTOR = 1, MAT_CONDUCTOR = 2, MAT_DIELECTRIC = 3, MAT_DIFFUSE = 4,
                   MAT_DIFFUSE_TRANSMISSION = 5, MAT_HAIR = 6, MAT_MEASURED = 7, MAT_SUBSURFACE = 8, MAT_THIN_DIELECTRIC = 9,
                   MAT_MIX = 10, MAT_INTERFACE = 11 /* null material ("interface") */, MAT_NTAGS = 12 };
enum LightTag { LIGHT_POINT = 0, LIGHT_DISTANT = 1, LIGHT_PROJECTION = 2, LIGHT_GONIOMETRIC = 3, LIGHT_SPOT = 4,
                LIGHT_DIFFUSE_AREA = 5, LIGHT_UNIFORM_INFINITE = 6, LIGHT_IMAGE_INFINITE = 7, LIGHT_PORTAL_IMAGE_INFINITE = 8 };
enum LightSamplerTag { LS_UNIFORM = 0, LS_POWER = 1, LS_EXHAUSTIVE = 2, LS_BVH = 3 };

struct Mat4 { float m[4][4]; };

struct DeviceCamera {
    Mat4 cameraFromRaster;      // ProjectiveCamera::cameraFromRaster (cameras.h:272)
    Mat4 renderFromCamera;      // CameraTransform::renderFromCamera (start transform; static camera)
    float lensRadius, focalDistance;
    int type, pad;
    Vec3 dxCamera, dyCamera;    // PerspectiveCamera::dxCamera / dyCamera (cameras.h:347-351), used for GBufferFilm ray differentials
    float rayDiffScale;         // max(.125, 1 / sqrt(spp))  (cpu/integrators.cpp:251-254)
    // Camera::Approximate_dp_dxy (cameras.h:155-183), scenes with textures only: the minimum differentials found by
    // CameraBase::FindMinimumDifferentials (cameras.cpp:154-200), the inverse camera transform and the spp scale
    Mat4 cameraFromRender;
    Vec3 minPosDifferentialX, minPosDifferentialY, minDirDifferentialX, minDirDifferentialY;
    float sppScale;             // max(.125, 1 / sqrt(samplesPerPixel))
};

struct DeviceFilm {
    int xres, yres;
    int pbx0, pby0, pbx1, pby1;          // pixel bounds
    float imagingRatio, maxComponentValue;
    float outRGBFromSensor[3][3];
    int isGBuffer, pad;
    // GBufferFilm (film.cpp:588-641): outputFromRender and whether it is applied inverted ("camera": RenderFromCamera inverted,
    // "world": WorldFromRender forward), plus the colour space's RGBFromXYZ for the albedo channel
    Mat4 outM, outMInv;
    int applyInverse, pad2;
    float rgbFromXYZ[3][3];
};

// GBufferFilm::Pixel members beyond {rgbSum, weightSum} (film.h:432-444); float / double exactly as the reference declares them
struct GBufferPixel {
    float pSum[3], dzdxSum, dzdySum, nSum[3], nsSum[3], uvSum[2];
    float varMean[3], varS[3];
    long long varN[3];
    double gBufferWeightSum;
    double rgbAlbedoSum[3];
};

struct DeviceFilter {          // FilterSampler tables (filters.cpp FilterSampler ctor)
    PC2DView distrib;
    const float *f;            // ny*nx tabulated filter values
    int nx, ny;
    float radiusX, radiusY;
    int sampled;               // 0: BoxFilter::Sample  1: FilterSampler table (gaussian, mitchell, sinc)  2: TriangleFilter::Sample
};

// One level of an image texture's MIP pyramid (Image::GeneratePyramid), texels decoded to float, nc interleaved channels
struct DeviceMIPLevel { const float *texels; int xres, yres; };
// FloatImageTexture / SpectrumImageTexture with a UVMapping (textures.h:86-106, 525-626)
struct DeviceTexture {
    int firstLevel, nLevels;   // into SceneView::mipLevels
    int nc;                    // channels of the image (1, 3 or 4)
    int wrap, filter;          // WrapMode, FilterFunction
    float maxAnisotropy;
    float su, sv, du, dv;
    float scale;
    int invert;
    // Texture classes without an image (textures.h:137-196 constant, 843-890 scale, 199-260 mix): children are texture indices
    int kind;                  // TEXKIND_*
    int child0, child1, child2;
    float cval;                // constant float
    SpectrumRec sval;          // constant spectrum
    int mapping;               // TextureMapping2D: 0 uv, 1 spherical, 2 cylindrical, 3 planar (textures.h:86-215)
    Mat4 textureFromRender;    // Inverse(renderFromTexture) for the three point mappings
    Vec3 vs, vt;               // planar (its ds / dt are du / dv)
};
enum { TEXKIND_IMAGEMAP = 0, TEXKIND_CONSTANT = 2, TEXKIND_SCALE = 3, TEXKIND_MIX = 4, TEXKIND_CHECKERBOARD = 5 };

struct DeviceMaterial {
    int type;                  // MaterialTag
    int flags;                 // bit0: remaproughness
    SpectrumRec reflectance;   // diffuse
    SpectrumRec eta, k;        // conductor (spectral), dielectric eta (eta.kind constant or spectrum)
    float uroughness, vroughness;
    // coated conductor
    float iface_uroughness, iface_vroughness, thickness, g;
    SpectrumRec iface_eta, albedo;
    int maxDepth, nSamples;
    int useReflectance;        // conductor specified by "reflectance" instead of eta/k
    int optionsSeed;           // Options->seed (LayeredBxDF seeds its RNG with it, bxdfs.h:510)
    SpectrumRec transmittance; // diffusetransmission
    float scale;               // diffusetransmission "scale"
    // image textures in place of the constants above (-1: the constant) and the "displacement" float texture of bump mapping
    int reflectanceTex, transmittanceTex, uroughnessTex, vroughnessTex, displacementTex;
    int normalMapTex;          // "normalmap" image (one level, RGB, linear)
};

struct DeviceMesh {
    int material;              // index into materials, -1 = none ("interface")
    int areaLightBase;         // index of the DiffuseAreaLight of triangle 0, or -1
    int flags;                 // bit0 flipNormal (reverseOrientation ^ swapsHandedness), bit1 hasN, bit2 hasUV
    int indexOffset;           // offset (in ints) of this mesh's vertexIndices in the global index pool
    int nOffset;               // offset (in float3) into the normal pool
    int uvOffset;              // offset (in float2) into the uv pool
    int mediumInside, mediumOutside;   // -1 = none
    int shape;                 // >= 0: this "mesh" is one quadric (index into SceneView::shapes, quadrics.h); its single pool record holds the bounds
    int patchBase;             // >= 0: a bilinear patch mesh; patch i of the mesh is SceneView::patches[4 * (patchBase + i)] (bilinear.h), its pool
                               // record holds the patch's bounds, indexPool has four vertex indices per patch (for N / uv)
};

struct DeviceLight {
    int type;                  // LightTag
    int tri;                   // DIFFUSE_AREA: global triangle id (pre-reorder id; tri data pool index)
    float scale;
    int twoSided;
    int LemitTable;            // offset into the dense table pool
    int instance;              // instance index for lights on instanced meshes (unused: pbrt forbids)
    // infinite lights
    float sceneRadius;
    Pt3 sceneCenter;
    int image;                 // ImageInfinite: index of image record
    Mat4 renderFromLight, lightFromRender;
    // point / spot: p = renderFromLight(Point3f(0,0,0)); distant: w = Normalize(renderFromLight(Vector3f(0,0,1))) (host, once)
    Pt3 p;
    Vec3 w;
    float cosFalloffStart, cosFalloffEnd;   // SpotLight (lights.cpp:1345-1354)
};

// Light BVH node with CompactLightBounds already de-quantised on the host (lightsamplers.h:104-200):
// the de-quantisation is deterministic float arithmetic, so doing it once is value-identical.
struct DeviceLightBVHNode {
    Vec3 w;                    // Normalize(Vector3f(OctahedralVector))
    float phi;
    float cosTheta_o, cosTheta_e;
    Pt3 bmin, bmax;            // CompactLightBounds::Bounds(allb)
    int twoSided;
    int childOrLightIndex;
    int isLeaf;
    int pad;
};

struct DeviceLightSampler {
    int type;                  // LightSamplerTag
    int nLights, nInfinite, nNodes;
    const DeviceLightBVHNode *nodes;
    const int *infiniteLights; // light indices
    const uint32_t *bitTrails; // per light; 0xffffffff for lights not in the BVH
    const int *lightInBVH;     // per light 0/1
};

struct DeviceImageLight {      // ImageInfiniteLight (lights.h:544-628)
    const float *rgb;          // res*res*3
    int res;
    PC2DView distrib, compensatedDistrib;
    // image colour space = sRGB (tables shared)
};

// One object instance (TransformedPrimitive, cpu/primitive.cpp:112-130): renderFromInstance with its tracked inverse,
// the 8-wide BVH of its object definition, and the global id of the definition's first triangle.
// Instance 0 of an instanced scene is the identity pseudo-instance that holds all top-level (non-instanced) triangles.
struct DeviceInstance {
    Mat4 m, mInv;
    const float4 *nodes;
    const float4 *tris;
    int triBase;
    int identity;
    int pad0, pad1;
};

struct IntegratorParams {
    int maxDepth;
    int regularize;
};

// Transform::ApplyInverse(const Ray &, Float *tMax) (util/transform.h:416-429): dt is what the caller subtracts from tMax
struct InstRay { Pt3 o; Vec3 d; float dt; };
BPT_HD InstRay ApplyInverseRay(const Mat4 &Minv, Pt3 o, Vec3 d) {
    const float (*m)[4] = Minv.m;
    float x = o.x, y = o.y, z = o.z;
    float xp = (m[0][0] * x + m[0][1] * y) + (m[0][2] * z + m[0][3]);
    float yp = (m[1][0] * x + m[1][1] * y) + (m[1][2] * z + m[1][3]);
    float zp = (m[2][0] * x + m[2][1] * y) + (m[2][2] * z + m[2][3]);
    float wp = (m[3][0] * x + m[3][1] * y) + (m[3][2] * z + m[3][3]);
    Vec3 e;   // exact input point: util/transform.cpp:263-276 (no |m[i][3]| term)
    e.x = gamma_n(3) * (fabsf(m[0][0] * x) + fabsf(m[0][1] * y) + fabsf(m[0][2] * z));
    e.y = gamma_n(3) * (fabsf(m[1][0] * x) + fabsf(m[1][1] * y) + fabsf(m[1][2] * z));
    e.z = gamma_n(3) * (fabsf(m[2][0] * x) + fabsf(m[2][1] * y) + fabsf(m[2][2] * z));
    Pt3i oi(Pt3(xp, yp, zp), e);
    if (wp != 1) {
        auto divI = [](Interval i, float f) { return f > 0 ? Interval(DivRoundDown(i.low, f), DivRoundUp(i.high, f))
                                                          : Interval(DivRoundDown(i.high, f), DivRoundUp(i.low, f)); };
        oi.x = divI(oi.x, wp); oi.y = divI(oi.y, wp); oi.z = divI(oi.z, wp);
    }
    Vec3 dn(m[0][0] * d.x + m[0][1] * d.y + m[0][2] * d.z, m[1][0] * d.x + m[1][1] * d.y + m[1][2] * d.z,
            m[2][0] * d.x + m[2][1] * d.y + m[2][2] * d.z);
    InstRay r; r.dt = 0;
    float lengthSquared = LengthSquared(dn);
    if (lengthSquared > 0) {
        Vec3 oError(oi.x.Width() / 2, oi.y.Width() / 2, oi.z.Width() / 2);
        float dt = Dot(Abs(dn), oError) / lengthSquared;
        Vec3 off = dn * dt;
        oi.x = oi.x + off.x; oi.y = oi.y + off.y; oi.z = oi.z + off.z;
        r.dt = dt;
    }
    r.o = oi.p(); r.d = dn;
    return r;
}


}  // namespace bpt
