/* b200pt.h -- C ABI of the B200-native wavefront path-tracing engine.
 *
 * This is the boundary a pbrt-v4 maintainer binds to replace the WavefrontPathIntegrator hot path
 * (reference: src/pbrt/wavefront/integrator.h:32-54 WavefrontAggregate, integrator.cpp:80-287 ctor,
 * :290-493 Render(); src/pbrt/gpu/optix/aggregate.h:32-37 OptiXAggregate ctor).  Plain C structs and
 * pointers only: caller-owned host arrays are copied at bpt_scene_* time; the engine owns all device memory.
 * One engine per GPU; an engine is not thread-safe.  Every function returns BPT_OK (0) or a negative
 * error code; bpt_last_error() gives the message (the reference aborts via LOG_FATAL instead,
 * gpu/util.h:35-49).  There is no CPU fallback: every entry point that computes fails with
 * BPT_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Tag values (camera/film/filter/sampler/material/light/light-sampler types) are the positions in the
 * reference's TaggedPointer type lists: base/camera.h:34, base/film.h:25, base/filter.h:25,
 * base/sampler.h:40-42, base/material.h:36-39, base/light.h:40-43, base/lightsampler.h:29.
 */
#ifndef B200PT_H
#define B200PT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPT_OK 0
#define BPT_ERR_INVALID (-1)
#define BPT_ERR_NO_DEVICE (-2)
#define BPT_ERR_CUDA (-3)
#define BPT_ERR_UNSUPPORTED (-4)
#define BPT_ERR_IO (-5)

typedef struct bpt_engine bpt_engine;

/* ---- tags ---- */
enum { BPT_CAMERA_PERSPECTIVE = 0, BPT_CAMERA_ORTHOGRAPHIC = 1, BPT_CAMERA_SPHERICAL = 2, BPT_CAMERA_REALISTIC = 3 };
enum { BPT_FILM_RGB = 0, BPT_FILM_GBUFFER = 1, BPT_FILM_SPECTRAL = 2 };
enum { BPT_FILTER_BOX = 0, BPT_FILTER_GAUSSIAN = 1, BPT_FILTER_MITCHELL = 2, BPT_FILTER_SINC = 3, BPT_FILTER_TRIANGLE = 4 };
enum { BPT_SAMPLER_PMJ02BN = 0, BPT_SAMPLER_INDEPENDENT = 1, BPT_SAMPLER_STRATIFIED = 2, BPT_SAMPLER_HALTON = 3,
       BPT_SAMPLER_PADDED_SOBOL = 4, BPT_SAMPLER_SOBOL = 5, BPT_SAMPLER_ZSOBOL = 6 };
enum { BPT_RANDOMIZE_NONE = 0, BPT_RANDOMIZE_PERMUTE_DIGITS = 1, BPT_RANDOMIZE_FAST_OWEN = 2, BPT_RANDOMIZE_OWEN = 3 };
enum { BPT_MAT_COATED_DIFFUSE = 0, BPT_MAT_COATED_CONDUCTOR = 1, BPT_MAT_CONDUCTOR = 2, BPT_MAT_DIELECTRIC = 3,
       BPT_MAT_DIFFUSE = 4, BPT_MAT_DIFFUSE_TRANSMISSION = 5, BPT_MAT_HAIR = 6, BPT_MAT_MEASURED = 7,
       BPT_MAT_SUBSURFACE = 8, BPT_MAT_THIN_DIELECTRIC = 9, BPT_MAT_MIX = 10, BPT_MAT_INTERFACE = 11 };
enum { BPT_LIGHT_POINT = 0, BPT_LIGHT_DISTANT = 1, BPT_LIGHT_PROJECTION = 2, BPT_LIGHT_GONIOMETRIC = 3, BPT_LIGHT_SPOT = 4,
       BPT_LIGHT_DIFFUSE_AREA = 5, BPT_LIGHT_UNIFORM_INFINITE = 6, BPT_LIGHT_IMAGE_INFINITE = 7,
       BPT_LIGHT_PORTAL_IMAGE_INFINITE = 8 };
enum { BPT_LIGHTSAMPLER_UNIFORM = 0, BPT_LIGHTSAMPLER_POWER = 1, BPT_LIGHTSAMPLER_EXHAUSTIVE = 2, BPT_LIGHTSAMPLER_BVH = 3 };

/* How a spectrum-valued parameter was given in the scene description (paramdict.cpp GetOneSpectrum). */
enum { BPT_SPEC_NONE = 0, BPT_SPEC_RGB = 1, BPT_SPEC_CONSTANT = 2, BPT_SPEC_NAMED = 3, BPT_SPEC_SAMPLED = 4,
       BPT_SPEC_BLACKBODY = 5 };
typedef struct {
    int kind;
    float rgb[3];         /* BPT_SPEC_RGB */
    float value;          /* BPT_SPEC_CONSTANT value / BPT_SPEC_BLACKBODY temperature */
    char name[48];        /* BPT_SPEC_NAMED, e.g. "metal-Cu-eta" */
    const float *samples; /* BPT_SPEC_SAMPLED: n (lambda, value) pairs */
    int n_samples;
} bpt_spectrum;

typedef struct { float m[4][4]; } bpt_mat4;

typedef struct {
    int type;                     /* BPT_CAMERA_PERSPECTIVE | BPT_CAMERA_ORTHOGRAPHIC */
    bpt_mat4 camera_from_world;   /* CTM at the Camera directive */
    bpt_mat4 world_from_camera;   /* its tracked inverse (pbrt keeps both); all zeros => computed numerically */
    float fov, lens_radius, focal_distance;
    float frame_aspect_ratio;     /* <= 0: xres / yres */
    float screen_window[4];       /* used when has_screen_window */
    int has_screen_window;
    float shutter_open, shutter_close;
} bpt_camera;

typedef struct {
    int type;                     /* BPT_FILM_RGB | BPT_FILM_GBUFFER */
    int xres, yres;
    int pixel_bounds[4];          /* x0, x1, y0, y1 (exclusive upper); all 0 => full image */
    float iso, exposure_time, white_balance;
    float max_component_value;    /* <= 0: infinity */
    float diagonal;               /* mm, default 35 */
    int coordinate_system;        /* BPT_FILM_GBUFFER: 0 = "camera" (default), 1 = "world"  (film.cpp GBufferFilm::Create) */
} bpt_film;

typedef struct { int type; float xradius, yradius; float sigma, b, c, tau; } bpt_filter;
typedef struct {
    int type;                     /* BPT_SAMPLER_*: all but the MLT samplers are implemented */
    int spp; int seed; int randomize;
    int x_samples, y_samples, jitter;   /* STRATIFIED (samplers.cpp:297-319); spp is x_samples * y_samples then */
} bpt_sampler;
typedef struct { int max_depth; int regularize; int light_sampler; } bpt_integrator;

typedef struct {
    int type;                     /* BPT_MAT_* */
    bpt_spectrum reflectance;     /* diffuse; conductor "reflectance"; coatedconductor "reflectance" */
    bpt_spectrum eta, k;          /* conductor eta/k; dielectric eta (constant or spectrum) */
    float roughness, uroughness, vroughness;   /* < 0 means "not given" for u/v */
    int remap_roughness;
    /* coated conductor */
    float iface_roughness, iface_uroughness, iface_vroughness, thickness, g;
    bpt_spectrum iface_eta, conductor_eta, conductor_k, albedo;
    int max_depth, n_samples;
    /* coateddiffuse uses reflectance, roughness / uroughness / vroughness (the interface), thickness, eta, g, albedo, max_depth, n_samples;
     * thindielectric uses eta; diffusetransmission uses reflectance, transmittance and scale (<= 0 means 1) */
    bpt_spectrum transmittance;
    float scale;
    /* image textures in place of constants: 0 = none, k > 0 = texture k-1 of bpt_scene_set_textures.  "reflectance" / "transmittance"
     * spectrum textures (albedo), "roughness" / "uroughness" / "vroughness" float textures of conductor, dielectric, coateddiffuse,
     * and the "displacement" float texture of bump mapping (materials.h:107-138) */
    int reflectance_tex, transmittance_tex, uroughness_tex, vroughness_tex, displacement_tex;
    int normal_map_tex;           /* "normalmap": a BPT_TEXTURE_NORMALMAP entry of bpt_scene_set_textures (materials.h:86-105) */
} bpt_material;

/* Texture "imagemap" (FloatImageTexture / SpectrumImageTexture, textures.h:525-626) with a "uv" mapping (textures.h:86-106).  The image
 * is handed over as the reference's Image holds it after reading the file (util/image.cpp:877-921): 8-bit texels with their colour
 * encoding, or 32-bit float; 1 ("Y"), 3 (RGB) or 4 (RGBA) interleaved channels, top row first.  The engine builds the MIP pyramid. */
enum { BPT_TEXFILTER_POINT = 0, BPT_TEXFILTER_BILINEAR = 1, BPT_TEXFILTER_TRILINEAR = 2, BPT_TEXFILTER_EWA = 3 };
enum { BPT_TEXWRAP_BLACK = 0, BPT_TEXWRAP_CLAMP = 1, BPT_TEXWRAP_REPEAT = 2 };
enum { BPT_IMAGE_U8 = 0, BPT_IMAGE_F32 = 1 };
enum { BPT_ENCODING_LINEAR = 0, BPT_ENCODING_SRGB = 1 };
typedef struct {
    int format, encoding;         /* BPT_IMAGE_*, BPT_ENCODING_* (8-bit texels only) */
    int xres, yres, n_channels;
    const void *pixels;           /* copied */
    int filter;                   /* "filter": BPT_TEXFILTER_* (the reference's default is bilinear) */
    float max_anisotropy;         /* "maxanisotropy" (8) */
    int wrap;                     /* "wrap": BPT_TEXWRAP_* (repeat) */
    float scale; int invert;      /* "scale" (1), "invert" */
    float su, sv, du, dv;         /* "uscale" "vscale" "udelta" "vdelta" */
    int kind;                     /* BPT_TEXTURE_IMAGEMAP, or BPT_TEXTURE_NORMALMAP: the RGB image of a material's "normalmap" (read with the
                                   * linear encoding, looked up bilinearly without a pyramid; only the image fields are used), or one of the
                                   * image-less classes below */
    /* Texture "constant" | "scale" | "mix" (textures.h:137-196, 843-890, 199-260): children are texture ids (index + 1) of entries
     * EARLIER in the array; a constant parameter is its own BPT_TEXTURE_CONSTANT entry, as the reference wraps it in a
     * Float / SpectrumConstantTexture.  scale: tex1 = "tex", tex2 = "scale" (float).  mix: tex1, tex2, tex3 = "amount" (float). */
    int is_spectrum;              /* value type of this texture: 0 float, 1 spectrum */
    int tex1, tex2, tex3;
    float value_f;                /* BPT_TEXTURE_CONSTANT, float */
    bpt_spectrum value_s;         /* BPT_TEXTURE_CONSTANT, spectrum ("value") */
    /* "mapping" of imagemap / checkerboard textures (TextureMapping2D::Create, textures.cpp:49-73): uv (su sv du dv above), or spherical /
     * cylindrical / planar (textures.h:108-215), which look the shading point up through the inverse of the CTM at the Texture directive */
    int mapping;                  /* BPT_TEXMAP_* */
    bpt_mat4 world_from_texture, texture_from_world;   /* read like a mesh's matrices (bpt_scene_set_transform_space) */
    float v1[3], v2[3];           /* planar: "v1" (1 0 0), "v2" (0 1 0); its "udelta" / "vdelta" are du / dv */
} bpt_texture;
enum { BPT_TEXMAP_UV = 0, BPT_TEXMAP_SPHERICAL = 1, BPT_TEXMAP_CYLINDRICAL = 2, BPT_TEXMAP_PLANAR = 3 };
enum { BPT_TEXTURE_IMAGEMAP = 0, BPT_TEXTURE_NORMALMAP = 1, BPT_TEXTURE_CONSTANT = 2, BPT_TEXTURE_SCALE = 3, BPT_TEXTURE_MIX = 4,
       BPT_TEXTURE_CHECKERBOARD = 5 /* two-dimensional (textures.cpp:183-207): tex1, tex2 and the uv mapping su sv du dv */ };

typedef struct {
    bpt_spectrum L;               /* BPT_SPEC_NONE => colour-space illuminant */
    float scale;
    int two_sided;
    float power;                  /* <= 0: unused */
} bpt_area_light;

typedef struct {
    int n_vertices, n_triangles;
    const float *P;               /* 3*n_vertices, object space */
    const float *N;               /* optional, 3*n_vertices */
    const float *UV;              /* optional, 2*n_vertices */
    const int *indices;           /* 3*n_triangles */
    bpt_mat4 world_from_object;   /* CTM at the Shape directive */
    bpt_mat4 object_from_world;   /* its tracked inverse (pbrt composes inverses analytically, util/transform.h:148-151); all zeros => computed numerically */
    int reverse_orientation;
    int material;                 /* index into the material array, -1 = "interface" */
    int area_light;               /* index into the area-light array, -1 = none */
    int medium_inside, medium_outside;   /* -1 = none */
    int object_id;                /* -1: top-level shape; >= 0: belongs to instance definition object_id */
    int has_alpha;                /* Shape "float alpha" < 1: the stochastic alpha test of GeometricPrimitive::Intersect (cpu/primitive.cpp:56-75) */
    float alpha;                  /* constant alpha (texture-valued alpha: not implemented) */
} bpt_mesh;

/* Shape "sphere" / "disk" / "cylinder" (shapes.cpp:72-83, 107-117, 135-146: Sphere / Disk / Cylinder::Create).  Parameters as the
 * .pbrt file gives them (phimax in degrees; the constructors' clamping is the engine's business).  Intersected as custom primitives of
 * the BVH like the reference's GPU path does (gpu/optix/aggregate.cpp:848-1025, optix.cu:301-377).  Stochastic alpha on a quadric
 * is not implemented (the reference re-intersects the same shape behind a rejected hit, cpu/primitive.cpp:66-73). */
#define BPT_SHAPE_SPHERE 0
#define BPT_SHAPE_DISK 1
#define BPT_SHAPE_CYLINDER 2
typedef struct {
    int type;
    float radius;                 /* all three */
    float zmin, zmax;             /* sphere, cylinder */
    float phimax;                 /* degrees */
    float height, innerradius;    /* disk */
    bpt_mat4 world_from_object;   /* as bpt_mesh */
    bpt_mat4 object_from_world;
    int reverse_orientation;
    int material, area_light;
    int medium_inside, medium_outside;
    int object_id;
} bpt_quadric;

typedef struct {
    int object_id;
    bpt_mat4 world_from_instance;  /* CTM at the ObjectInstance directive */
    bpt_mat4 instance_from_world;  /* its tracked inverse; all zeros => computed numerically */
} bpt_instance;

typedef struct {
    int type;                     /* BPT_LIGHT_UNIFORM_INFINITE | _IMAGE_INFINITE | _POINT | _SPOT | _DISTANT */
    bpt_mat4 world_from_light;    /* CTM at the LightSource directive */
    bpt_mat4 light_from_world;    /* its tracked inverse; all zeros => computed numerically */
    bpt_spectrum L;               /* "L" (infinite, distant) or "I" (point, spot); BPT_SPEC_NONE: the colour space's illuminant */
    float scale;
    float illuminance;            /* infinite, distant: <= 0: unused */
    const float *image_rgb;       /* IMAGE_INFINITE: res*res*3 floats, equal-area octahedral parameterisation */
    int image_res;
    /* point / spot / distant lights (lights.cpp:192-213, 246-276, 1433-1464): the parameters as the scene file gives them; the
     * engine composes renderFromLight with Translate(from) / the spot's frame / the distant light's frame as the reference does */
    float from[3], to[3];         /* defaults (0,0,0), (0,0,1) */
    float power;                  /* point, spot: <= 0: unused */
    float cone_angle, cone_delta; /* spot: "coneangle" (30), "conedeltaangle" (5), degrees */
} bpt_light;

enum { BPT_MEDIUM_UNIFORMGRID = 0, BPT_MEDIUM_HOMOGENEOUS = 1, BPT_MEDIUM_RGBGRID = 2 };
/* Medium (media.h:217-352: HomogeneousMedium, GridMedium; created in media.cpp:165-201,249-330), emission through "Le" / "Lescale"
 * (no "temperature" grids). */
typedef struct {
    int type;                     /* BPT_MEDIUM_* */
    bpt_mat4 world_from_medium;   /* CTM at MakeNamedMedium */
    bpt_mat4 medium_from_world;   /* its tracked inverse; all zeros => computed numerically */
    bpt_spectrum sigma_a, sigma_s;
    float scale, g;
    float bounds[6];              /* p0, p1 */
    const float *density; int nx, ny, nz;
    bpt_spectrum Le; float Le_scale;                /* emission: "Le" (BPT_SPEC_NONE: none) and, homogeneous, the float "Lescale" */
    const float *Le_scale_grid;                      /* uniformgrid: the "Lescale" float array (nx*ny*nz values) or NULL */
    /* "rgbgrid" (RGBGridMedium, media.h:355-427, media.cpp:380-453): nx*ny*nz RGB triples each or NULL; scale = "scale", Le_scale = "Lescale" */
    const float *sigma_a_rgb, *sigma_s_rgb, *Le_rgb;
} bpt_medium;

typedef struct {
    double seconds_build, seconds_render;
    uint64_t camera_rays, indirect_rays, shadow_rays;
    /* traversal counters, filled only by a counting render (bpt_render flags bit0): closest-hit and shadow kernels */
    uint64_t nodes_visited, tris_tested, counted_rays;            /* closest-hit rays */
    uint64_t nodes_visited_shadow, tris_tested_shadow, counted_rays_shadow;
    int bvh_nodes, bvh_triangles, bvh_depth;
    uint64_t bytes_bvh_nodes, bytes_bvh_tris;
    int kernel_launches;
    /* per-kernel-class CUDA-event sums, filled only by a profiling render (bpt_render flags bit1) */
    double ms_trace_closest, ms_trace_shadow, ms_shade, ms_other;
    int launches_trace_closest, launches_trace_shadow;
    double ms_medium;             /* k_sample_medium + k_medium_scatter (profiling render) */
} bpt_stats;

/* ---- lifecycle ---- */
int bpt_create(int cuda_device, const char *tables_path /* NULL: <lib dir>/../data/pbrt_tables.bin */, bpt_engine **out);
void bpt_destroy(bpt_engine *e);
const char *bpt_last_error(const bpt_engine *e /* may be NULL for create errors */);

/* Engine options outside the scene description (kept across bpt_scene_begin).  "bvh_treelet_rounds": SAH treelet-restructuring rounds
 * of the BVH builder, 0 = plain LBVH (fastest build), default 2, < 0 restores the default -- the build-time / trace-time trade a
 * caller makes by how many samples it is going to render; "max_paths": path-state pool size (paths per wavefront pass). */
int bpt_set_option(bpt_engine *e, const char *name, long long value);

/* ---- scene description (copies everything) ---- */
int bpt_scene_begin(bpt_engine *e);
/* How the world_from_* matrices of meshes, instances, lights and media are to be read (the camera's are always world space).
 * BPT_SPACE_WORLD (default): the CTM at the directive; the engine composes RenderFromObject() = renderFromWorld * CTM and
 * RenderFromObject() * worldFromRender for instances exactly as BasicSceneBuilder does (scene.h:475-477, scene.cpp:375-393) --
 * inverses recomputed from the product matrix as the reference does, so the *_from_world members are ignored there.
 * BPT_SPACE_RENDER: the caller walks an already built BasicScene and passes ShapeSceneEntity::renderFromObject,
 * InstanceSceneEntity::renderFromInstance, the light's / medium's renderFromObject with their inverses as they are. */
#define BPT_SPACE_WORLD 0
#define BPT_SPACE_RENDER 1
int bpt_scene_set_transform_space(bpt_engine *e, int space);
int bpt_scene_set_camera(bpt_engine *e, const bpt_camera *c);
int bpt_scene_set_film(bpt_engine *e, const bpt_film *f);
int bpt_scene_set_filter(bpt_engine *e, const bpt_filter *f);
int bpt_scene_set_sampler(bpt_engine *e, const bpt_sampler *s);
int bpt_scene_set_integrator(bpt_engine *e, const bpt_integrator *i);
int bpt_scene_set_materials(bpt_engine *e, const bpt_material *m, int n);
int bpt_scene_set_textures(bpt_engine *e, const bpt_texture *t, int n);
int bpt_scene_set_area_lights(bpt_engine *e, const bpt_area_light *a, int n);
int bpt_scene_add_mesh(bpt_engine *e, const bpt_mesh *m);
/* Shape "bilinearmesh" (BilinearPatch::CreateMesh, shapes.cpp:910-999) and the quad faces of a PLY file (shapes.cpp:1435-1461): a bpt_mesh
 * whose `indices` hold FOUR vertex indices per patch (p00, p10, p01, p11) and whose n_triangles is the number of patches.  Intersected
 * as custom primitives like the quadrics (gpu/optix/optix.cu:379-469); area lights sample them as the reference does (spherical
 * rectangles for rectangular patches).  "emissionfilename" and alpha are not implemented. */
int bpt_scene_add_bilinear_mesh(bpt_engine *e, const bpt_mesh *m);
int bpt_scene_add_quadric(bpt_engine *e, const bpt_quadric *q);   /* takes its place among the meshes in scene order (area-light order, scene.cpp CreateLights) */
int bpt_scene_add_instance(bpt_engine *e, const bpt_instance *i);
int bpt_scene_add_light(bpt_engine *e, const bpt_light *l);
int bpt_scene_add_medium(bpt_engine *e, const bpt_medium *m);
int bpt_scene_set_camera_medium(bpt_engine *e, int medium);
/* Upload, build the BVH(s), light sampler and all tables.  After this the engine can render. */
int bpt_scene_commit(bpt_engine *e);

/* Convenience front end: parse a .pbrt file (subset: see INTEGRATION.md) and commit it. */
int bpt_load_pbrt(bpt_engine *e, const char *path);
/* Same with the command-line overrides of the reference (cmd/pbrt.cpp --spp, --seed); spp_override <= 0: none. */
int bpt_load_pbrt_ex(bpt_engine *e, const char *path, int spp_override, int seed);
const char *bpt_frontend_error(void);

/* ---- rendering ---- */
/* Render sample indices [first_sample, first_sample + n_samples) of the pixels in pixel_bounds
 * (x0, x1, y0, y1; NULL = the film's pixel bounds) and accumulate them into the film.
 * flags: bit0 = count traversal statistics (slower counting kernels); bit1 = time every kernel class with CUDA events. */
int bpt_render(bpt_engine *e, int first_sample, int n_samples, const int *pixel_bounds, int flags);
/* Multi-GPU tiling inside one call: renders the rows y of the film's pixel bounds with ((y - y_min) / band_rows) % n_parts == part
 * (the scanline-band partition of SURVEY 8(e)); all of a rank's bands share the wavefront passes, which keeps the ray queues as
 * long as in a single-GPU render. */
int bpt_render_bands(bpt_engine *e, int first_sample, int n_samples, int band_rows, int n_parts, int part, int flags);
/* The same two calls on a caller-supplied CUDA stream (a cudaStream_t passed as void *; NULL = the engine's own stream) -- the
 * `cudaStream_t` argument SURVEY 8(b) proposes; the reference issues everything on one stream with a single GPUWait() at the end
 * (wavefront/integrator.cpp:485).  With BPT_RENDER_ASYNC in flags the call returns as soon as the kernels are enqueued (scenes
 * without media / interface surfaces; those need a host-side queue-empty test per segment and synchronise as bpt_render does);
 * bpt_sync() -- or any bpt_film_* / bpt_get_stats call -- waits for it. */
#define BPT_RENDER_COUNT 1
#define BPT_RENDER_PROFILE 2
#define BPT_RENDER_ASYNC 4
int bpt_render_stream(bpt_engine *e, int first_sample, int n_samples, const int *pixel_bounds, int flags, void *cuda_stream);
int bpt_render_bands_stream(bpt_engine *e, int first_sample, int n_samples, int band_rows, int n_parts, int part, int flags, void *cuda_stream);
int bpt_sync(bpt_engine *e);
/* pixelsamples of the committed sampler (what `--spp` / the scene file asked for); negative before commit */
int bpt_sampler_spp(const bpt_engine *e);
/* GBufferFilm::GetImage (film.cpp:690-790): 25 floats per pixel of the full film, in the reference's channel order
 * R G B Albedo.{R,G,B} P.{X,Y,Z} dzdx dzdy N.{X,Y,Z} Ns.{X,Y,Z} u v Variance.{R,G,B} RelativeVariance.{R,G,B}.
 * Only valid when the film type is BPT_FILM_GBUFFER. */
int bpt_film_resolve_gbuffer(bpt_engine *e, float *out25);
int bpt_film_clear(bpt_engine *e);
int bpt_stats_reset(bpt_engine *e);
/* Film accumulators: xres*yres*4 doubles {rgbSum[3], weightSum} over the FULL resolution (zero outside
 * rendered tiles), row-major, y down -- RGBFilm::Pixel (film.h:302-307). */
int bpt_film_read(bpt_engine *e, double *rgbw_host);
int bpt_film_write(bpt_engine *e, const double *rgbw_host);                /* replace accumulators (after a reduce) */
int bpt_film_copy_to_device(bpt_engine *e, void *dst_device_ptr);         /* d2d copy of the same layout */
int bpt_film_copy_from_device(bpt_engine *e, const void *src_device_ptr);
/* Final image, RGBFilm::GetPixelRGB (film.h:258-274): xres*yres*3 floats, output colour space. */
int bpt_film_resolve(bpt_engine *e, float *rgb_host);
int bpt_film_dims(const bpt_engine *e, int *xres, int *yres);
/* The resolved film as an OpenEXR file (single-part scan-line, FLOAT, uncompressed): R G B, or GBufferFilm's 25 channels under
 * the reference's channel names -- what RGBFilm / GBufferFilm::WriteImage hand to a denoiser (film.cpp:533-565, 688-803). */
int bpt_film_write_exr(bpt_engine *e, const char *path);
/* the writer alone (host only): n_channels names, pixels interleaved [y][x][channel] */
int bpt_write_exr_file(const char *path, int width, int height, int n_channels, const char *const *names, const float *interleaved);
int bpt_get_stats(bpt_engine *e, bpt_stats *out);

/* ---- ray-level entry points (WavefrontAggregate::IntersectClosest / IntersectShadow on raw rays) ---- */
/* rays: n x 8 floats {o.xyz, tMax, d.xyz, unused}; hits: n x 4 {prim (int bits, -1 = miss), b0, b1, b2}, t in t_out */
int bpt_trace_closest(bpt_engine *e, const float *rays_host, int n, float *hits_host, float *t_out_host);
int bpt_trace_any(bpt_engine *e, const float *rays_host, int n, int *occluded_host);
/* device-resident variants for benchmarking (pointers are device pointers) */
int bpt_trace_closest_device(bpt_engine *e, const void *rays_dev, int n, void *hits_dev, int flags);

/* ---- unit-level probes used by the parity tests (device evaluation of restated reference functions) ---- */
int bpt_probe_sampler(bpt_engine *e, const int *pixels_xy, const int *sample_index, const int *dimension, int n,
                      float *out_1d, float *out_2d /* 2n */);
int bpt_probe_intersect_triangle(bpt_engine *e, const float *rays /* n x 7: o, d, tMax */, const float *tris /* n x 9 */,
                                 int n, float *out /* n x 5: hit, b0, b1, b2, t */);

/* device evaluation of libdevice / double-rounded transcendental candidates (fn: see wavefront.cu k_probe_math) */
int bpt_probe_math(bpt_engine *e, int fn, const float *x, const float *y, int n, float *out);

#ifdef __cplusplus
}
#endif
#endif
