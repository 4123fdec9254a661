// b200_binding.cpp -- TEST INFRASTRUCTURE (CPU host, built only by oracle/build_ref.py against oracle/_ref/lib/libpbrt.a):
// the reference-side binding of INTEGRATION.md, compiled for real.  The reference's own tokenizer, parser and BasicSceneBuilder
// read the .pbrt file into a BasicScene; RenderB200() walks that scene and drives the engine through the C ABI of
// include/b200pt.h (loaded with dlopen, $BPT_LIB or <repo>/pbrt-v4_b200/libb200pt.so); the image is written by the reference's
// own Image::Write.  It is what a maintainer would add as src/pbrt/wavefront/b200.cpp next to RenderWavefront
// (wavefront/wavefront.cpp:14-70, cmd/pbrt.cpp:285-288).
//
// Entities reach the engine in RENDERING space (bpt_scene_set_transform_space(BPT_SPACE_RENDER)): ShapeSceneEntity::renderFromObject,
// InstanceSceneEntity::renderFromInstance (scene.h:88-115,187-225) are passed as the reference built them.  BasicScene does not
// keep the light and medium entities (it hands them to asynchronous creation jobs, scene.cpp AddLight / AddMedium), and the
// builder's option entities are private, so the binding derives from BasicSceneBuilder, records lights and media at their
// directives, and reads the private members through the "friend access" below -- a maintainer would add a friend declaration.
//
//   pbrt_b200 [--seed n] [--spp n] [--outfile f] [--quiet] scene.pbrt       render through the engine
//   BPT_LIB=tests/tools/build/libabi_record.so BPT_RECORD_OUT=calls.txt pbrt_b200 scene.pbrt   (CPU only) dump the ABI call stream
#include <dlfcn.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include <pbrt/pbrt.h>
#include <pbrt/cameras.h>
#include <pbrt/cpu/primitive.h>
#include <pbrt/options.h>
#include <pbrt/paramdict.h>
#include <pbrt/parser.h>
#include <pbrt/util/args.h>
#include <pbrt/util/colorspace.h>
#include <pbrt/util/containers.h>
#include <pbrt/util/error.h>
#include <pbrt/util/file.h>
#include <pbrt/util/image.h>
#include <pbrt/util/memory.h>
#include <pbrt/util/mesh.h>
#include <pbrt/util/parallel.h>
#include <pbrt/util/pstd.h>
#include <pbrt/util/string.h>
#include <pbrt/util/transform.h>
// friend access to BasicSceneBuilder / BasicScene internals (see the header comment)
#define private public
#define protected public
#include <pbrt/scene.h>
#undef private
#undef protected

#include "../include/b200pt.h"

namespace pbrt {

// ---- the C ABI, resolved at run time so that the binding links without CUDA
struct Api {
    void *h = nullptr;
#define BPT_FN(name) decltype(&::name) name = nullptr;
    BPT_FN(bpt_create) BPT_FN(bpt_destroy) BPT_FN(bpt_last_error) BPT_FN(bpt_scene_begin) BPT_FN(bpt_scene_set_transform_space)
    BPT_FN(bpt_scene_set_camera) BPT_FN(bpt_scene_set_film) BPT_FN(bpt_scene_set_filter) BPT_FN(bpt_scene_set_sampler)
    BPT_FN(bpt_scene_set_integrator) BPT_FN(bpt_scene_set_materials) BPT_FN(bpt_scene_set_textures) BPT_FN(bpt_scene_set_area_lights) BPT_FN(bpt_scene_add_mesh) BPT_FN(bpt_scene_add_quadric) BPT_FN(bpt_scene_add_bilinear_mesh)
    BPT_FN(bpt_scene_add_instance) BPT_FN(bpt_scene_add_light) BPT_FN(bpt_scene_add_medium) BPT_FN(bpt_scene_set_camera_medium)
    BPT_FN(bpt_scene_commit) BPT_FN(bpt_render) BPT_FN(bpt_film_resolve) BPT_FN(bpt_film_resolve_gbuffer)
#undef BPT_FN
    void Load(const std::string &path) {
        h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) ErrorExit("%s", dlerror());
#define BPT_FN(name) name = (decltype(name))dlsym(h, #name); if (!name) ErrorExit("%s: symbol %s missing", path, #name);
        BPT_FN(bpt_create) BPT_FN(bpt_destroy) BPT_FN(bpt_last_error) BPT_FN(bpt_scene_begin) BPT_FN(bpt_scene_set_transform_space)
        BPT_FN(bpt_scene_set_camera) BPT_FN(bpt_scene_set_film) BPT_FN(bpt_scene_set_filter) BPT_FN(bpt_scene_set_sampler)
        BPT_FN(bpt_scene_set_integrator) BPT_FN(bpt_scene_set_materials) BPT_FN(bpt_scene_set_textures) BPT_FN(bpt_scene_set_area_lights) BPT_FN(bpt_scene_add_mesh) BPT_FN(bpt_scene_add_quadric) BPT_FN(bpt_scene_add_bilinear_mesh)
        BPT_FN(bpt_scene_add_instance) BPT_FN(bpt_scene_add_light) BPT_FN(bpt_scene_add_medium) BPT_FN(bpt_scene_set_camera_medium)
        BPT_FN(bpt_scene_commit) BPT_FN(bpt_render) BPT_FN(bpt_film_resolve) BPT_FN(bpt_film_resolve_gbuffer)
#undef BPT_FN
    }
};

static bpt_mat4 ToB(const SquareMatrix<4> &m) { bpt_mat4 r; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = m[i][j]; return r; }

// ---- the builder: BasicSceneBuilder plus a record of what BasicScene does not keep
struct LightRecord { std::string type; ParameterDictionary params; Transform renderFromLight; };
struct MediumRecord { std::string name; ParameterDictionary params; Transform renderFromMedium; };
struct TextureRecord { std::string name, type, texname; ParameterDictionary params; FileLoc loc; pbrt::Transform renderFromTexture; };

class B200SceneBuilder : public BasicSceneBuilder {
  public:
    B200SceneBuilder(BasicScene *scene) : BasicSceneBuilder(scene) {}
    void Camera(const std::string &name, ParsedParameterVector params, FileLoc loc) override {
        cameraFromWorld = graphicsState.ctm[0];     // the CTM the reference builds CameraTransform from (scene.cpp:139-156)
        BasicSceneBuilder::Camera(name, std::move(params), loc);
    }
    void LightSource(const std::string &name, ParsedParameterVector params, FileLoc loc) override {
        // (the dictionary is built as BasicSceneBuilder::LightSource builds it: parameters + light attributes, current colour space)
        ParsedParameterVector copy = params;
        lights.push_back(LightRecord{name, ParameterDictionary(std::move(copy), graphicsState.lightAttributes, graphicsState.colorSpace), RenderFromObject(0)});
        BasicSceneBuilder::LightSource(name, std::move(params), loc);
    }
    void MakeNamedMedium(const std::string &name, ParsedParameterVector params, FileLoc loc) override {
        ParsedParameterVector copy = params;
        media.push_back(MediumRecord{NormalizeUTF8(name), ParameterDictionary(std::move(copy), graphicsState.mediumAttributes, graphicsState.colorSpace), RenderFromObject(0)});
        BasicSceneBuilder::MakeNamedMedium(name, std::move(params), loc);
    }
    void Texture(const std::string &name, const std::string &type, const std::string &texname, ParsedParameterVector params, FileLoc loc) override {
        // BasicScene hands image textures to loading jobs and keeps no entity for them (scene.cpp:912-1010): record the directive
        ParsedParameterVector copy = params;
        textures.push_back(TextureRecord{NormalizeUTF8(name), type, texname, ParameterDictionary(std::move(copy), graphicsState.textureAttributes, graphicsState.colorSpace), loc, RenderFromObject(0)});
        BasicSceneBuilder::Texture(name, type, texname, std::move(params), loc);
    }
    class Transform cameraFromWorld;
    std::vector<TextureRecord> textures;
    std::vector<LightRecord> lights;
    std::vector<MediumRecord> media;
};

// ---- parameter helpers
struct Stores {
    std::vector<std::unique_ptr<std::vector<float>>> floats;
    std::vector<std::unique_ptr<std::vector<int>>> ints;
    float *F(std::vector<float> v) { floats.emplace_back(new std::vector<float>(std::move(v))); return floats.back()->data(); }
    int *I(std::vector<int> v) { ints.emplace_back(new std::vector<int>(std::move(v))); return ints.back()->data(); }
};

// How a spectrum-valued parameter was written (the engine restates ParameterDictionary::GetOneSpectrum itself, paramdict.cpp)
static bpt_spectrum GetSpectrum(const ParameterDictionary &dict, const std::string &name, Stores &st) {
    bpt_spectrum s; std::memset(&s, 0, sizeof(s));
    for (const ParsedParameter *p : dict.GetParameterVector()) {
        if (p->name != name) continue;
        p->lookedUp = true;
        if (p->type == "rgb" && p->floats.size() == 3) { s.kind = BPT_SPEC_RGB; for (int k = 0; k < 3; ++k) s.rgb[k] = p->floats[k]; return s; }
        if (p->type == "float" && p->floats.size() == 1) { s.kind = BPT_SPEC_CONSTANT; s.value = p->floats[0]; return s; }
        if (p->type == "blackbody" && p->floats.size() == 1) { s.kind = BPT_SPEC_BLACKBODY; s.value = p->floats[0]; return s; }
        if (p->type == "spectrum") {
            if (!p->strings.empty()) { s.kind = BPT_SPEC_NAMED; std::strncpy(s.name, p->strings[0].c_str(), sizeof(s.name) - 1); return s; }
            if (p->floats.size() >= 2 && p->floats.size() % 2 == 0) {
                s.kind = BPT_SPEC_SAMPLED; s.n_samples = (int)p->floats.size() / 2;
                s.samples = st.F(std::vector<float>(p->floats.begin(), p->floats.end()));
                return s;
            }
        }
        ErrorExit(&p->loc, "parameter \"%s\": spectrum specification of type \"%s\" is not supported by the engine (textures: f1)", name, p->type);
    }
    s.kind = BPT_SPEC_NONE;
    return s;
}

// Material parameters that name a texture ("texture reflectance" "name"): the id (index + 1) the texture got in bpt_scene_set_textures
struct TexIds { std::map<std::string, int> floats, spectra; std::function<int(const std::string &)> normalMap; };
static const ParsedParameter *FindParam(const ParameterDictionary &dict, const std::string &name) {
    for (const ParsedParameter *p : dict.GetParameterVector()) if (p->name == name) return p;
    return nullptr;
}
static bpt_spectrum GetSpectrumOrTexture(const ParameterDictionary &d, const std::string &name, Stores &st, const TexIds &tx, int *tex, const FileLoc *loc) {
    *tex = 0;
    const ParsedParameter *p = FindParam(d, name);
    if (p && p->type == "texture") {
        p->lookedUp = true;
        auto it = p->strings.empty() ? tx.spectra.end() : tx.spectra.find(p->strings[0]);
        if (it == tx.spectra.end()) ErrorExit(loc, "parameter \"%s\": unknown spectrum texture", name);
        *tex = it->second;
        bpt_spectrum s; std::memset(&s, 0, sizeof(s));
        return s;
    }
    return GetSpectrum(d, name, st);
}
static float GetFloatOrTexture(const ParameterDictionary &d, const std::string &name, float dflt, const TexIds &tx, int *tex, const FileLoc *loc) {
    *tex = 0;
    const ParsedParameter *p = FindParam(d, name);
    if (p && p->type == "texture") {
        p->lookedUp = true;
        auto it = p->strings.empty() ? tx.floats.end() : tx.floats.find(p->strings[0]);
        if (it == tx.floats.end()) ErrorExit(loc, "parameter \"%s\": unknown float texture", name);
        *tex = it->second;
        return dflt;
    }
    return d.GetOneFloat(name, dflt);
}
// "roughness" / "uroughness" / "vroughness" as the materials' Create() functions resolve them (materials.cpp)
static void GetRoughness(const ParameterDictionary &d, const std::string &prefix, const TexIds &tx, float *rough, float *u, float *v, int *utex, int *vtex, const FileLoc *loc) {
    int rt = 0, ut = 0, vt = 0;
    *rough = GetFloatOrTexture(d, prefix + "roughness", 0.f, tx, &rt, loc);
    const bool haveU = FindParam(d, prefix + "uroughness") != nullptr, haveV = FindParam(d, prefix + "vroughness") != nullptr;
    *u = GetFloatOrTexture(d, prefix + "uroughness", *rough, tx, &ut, loc);
    *v = GetFloatOrTexture(d, prefix + "vroughness", *rough, tx, &vt, loc);
    if (utex) { *utex = haveU ? ut : rt; *vtex = haveV ? vt : rt; }
    else if (rt || ut || vt) ErrorExit(loc, "parameter \"%sroughness\": textures are not implemented by the engine here", prefix);
}

static bpt_material MakeMaterial(const std::string &type, const ParameterDictionary &d, Stores &st, const TexIds &tx, const FileLoc *loc) {
    bpt_material m; std::memset(&m, 0, sizeof(m));
    m.uroughness = m.vroughness = m.iface_uroughness = m.iface_vroughness = -1;
    if (std::string nm = ResolveFilename(d.GetOneString("normalmap", "")); !nm.empty()) m.normal_map_tex = tx.normalMap(nm);
    if (const ParsedParameter *p = FindParam(d, "displacement")) {
        auto it = (p->type != "texture" || p->strings.empty()) ? tx.floats.end() : tx.floats.find(p->strings[0]);
        if (it == tx.floats.end()) ErrorExit(loc, "\"displacement\" must name a float texture");
        p->lookedUp = true;
        m.displacement_tex = it->second;
    }
    if (type == "diffuse") { m.type = BPT_MAT_DIFFUSE; m.reflectance = GetSpectrumOrTexture(d, "reflectance", st, tx, &m.reflectance_tex, loc); return m; }
    if (type == "interface" || type == "none" || type == "") { m.type = BPT_MAT_INTERFACE; return m; }
    if (type == "conductor") {
        m.type = BPT_MAT_CONDUCTOR;
        m.eta = GetSpectrum(d, "eta", st); m.k = GetSpectrum(d, "k", st); m.reflectance = GetSpectrumOrTexture(d, "reflectance", st, tx, &m.reflectance_tex, loc);
        GetRoughness(d, "", tx, &m.roughness, &m.uroughness, &m.vroughness, &m.uroughness_tex, &m.vroughness_tex, loc);
        m.remap_roughness = d.GetOneBool("remaproughness", true);
        return m;
    }
    if (type == "dielectric") {
        m.type = BPT_MAT_DIELECTRIC;
        m.eta = GetSpectrum(d, "eta", st);
        GetRoughness(d, "", tx, &m.roughness, &m.uroughness, &m.vroughness, &m.uroughness_tex, &m.vroughness_tex, loc);
        m.remap_roughness = d.GetOneBool("remaproughness", true);
        return m;
    }
    if (type == "coatedconductor") {
        m.type = BPT_MAT_COATED_CONDUCTOR;
        m.iface_eta = GetSpectrum(d, "interface.eta", st); m.conductor_eta = GetSpectrum(d, "conductor.eta", st);
        m.conductor_k = GetSpectrum(d, "conductor.k", st); m.reflectance = GetSpectrumOrTexture(d, "reflectance", st, tx, &m.reflectance_tex, loc); m.albedo = GetSpectrum(d, "albedo", st);
        GetRoughness(d, "interface.", tx, &m.iface_roughness, &m.iface_uroughness, &m.iface_vroughness, nullptr, nullptr, loc);
        GetRoughness(d, "conductor.", tx, &m.roughness, &m.uroughness, &m.vroughness, &m.uroughness_tex, &m.vroughness_tex, loc);
        m.thickness = d.GetOneFloat("thickness", .01f); m.g = d.GetOneFloat("g", 0.f);
        m.max_depth = d.GetOneInt("maxdepth", 10); m.n_samples = d.GetOneInt("nsamples", 1);
        m.remap_roughness = d.GetOneBool("remaproughness", true);
        return m;
    }
    if (type == "coateddiffuse") {
        m.type = BPT_MAT_COATED_DIFFUSE;
        m.reflectance = GetSpectrumOrTexture(d, "reflectance", st, tx, &m.reflectance_tex, loc); m.eta = GetSpectrum(d, "eta", st); m.albedo = GetSpectrum(d, "albedo", st);
        GetRoughness(d, "", tx, &m.roughness, &m.uroughness, &m.vroughness, &m.uroughness_tex, &m.vroughness_tex, loc);
        m.thickness = d.GetOneFloat("thickness", .01f); m.g = d.GetOneFloat("g", 0.f);
        m.max_depth = d.GetOneInt("maxdepth", 10); m.n_samples = d.GetOneInt("nsamples", 1);
        m.remap_roughness = d.GetOneBool("remaproughness", true);
        return m;
    }
    if (type == "diffusetransmission") {
        m.type = BPT_MAT_DIFFUSE_TRANSMISSION;
        m.reflectance = GetSpectrumOrTexture(d, "reflectance", st, tx, &m.reflectance_tex, loc);
        m.transmittance = GetSpectrumOrTexture(d, "transmittance", st, tx, &m.transmittance_tex, loc);
        m.scale = d.GetOneFloat("scale", 1.f);
        return m;
    }
    if (type == "thindielectric") { m.type = BPT_MAT_THIN_DIELECTRIC; m.eta = GetSpectrum(d, "eta", st); return m; }
    ErrorExit(loc, "material \"%s\" is not implemented by the engine", type);
    return m;
}

static void Check(Api &api, bpt_engine *e, int rc, const char *what) {
    if (rc != BPT_OK) ErrorExit("%s failed (%d): %s", what, rc, api.bpt_last_error(e));
}

// ---- the driver a maintainer would call next to RenderWavefront(scene)
void RenderB200(BasicScene &scene, B200SceneBuilder &builder, Api &api) {
    Stores st;
    bpt_engine *e = nullptr;
    if (api.bpt_create(Options->gpuDevice ? *Options->gpuDevice : 0, nullptr, &e) != BPT_OK) ErrorExit("bpt_create: %s", api.bpt_last_error(nullptr));
    Check(api, e, api.bpt_scene_begin(e), "bpt_scene_begin");
    Check(api, e, api.bpt_scene_set_transform_space(e, BPT_SPACE_RENDER), "bpt_scene_set_transform_space");

    // camera (cameras.cpp CameraBaseParameters + PerspectiveCamera::Create)
    {
        const CameraSceneEntity &ce = builder.camera;
        if (ce.name != "perspective" && ce.name != "orthographic") ErrorExit(&ce.loc, "camera \"%s\" is not implemented by the engine", ce.name.ToString());
        const ParameterDictionary &d = ce.parameters;
        bpt_camera c; std::memset(&c, 0, sizeof(c));
        c.type = ce.name == "perspective" ? BPT_CAMERA_PERSPECTIVE : BPT_CAMERA_ORTHOGRAPHIC;
        c.camera_from_world = ToB(builder.cameraFromWorld.GetMatrix()); c.world_from_camera = ToB(builder.cameraFromWorld.GetInverseMatrix());
        c.fov = d.GetOneFloat("fov", 90.f); c.lens_radius = d.GetOneFloat("lensradius", 0.f); c.focal_distance = d.GetOneFloat("focaldistance", 1e6f);
        c.frame_aspect_ratio = d.GetOneFloat("frameaspectratio", -1.f);
        std::vector<Float> sw = d.GetFloatArray("screenwindow");
        if (sw.size() == 4) { c.has_screen_window = 1; for (int k = 0; k < 4; ++k) c.screen_window[k] = sw[k]; }
        c.shutter_open = d.GetOneFloat("shutteropen", 0.f); c.shutter_close = d.GetOneFloat("shutterclose", 1.f);
        Check(api, e, api.bpt_scene_set_camera(e, &c), "bpt_scene_set_camera");
    }
    // film (film.cpp FilmBaseParameters, RGBFilm::Create, GBufferFilm::Create)
    bpt_film film; std::memset(&film, 0, sizeof(film));
    std::string outName;
    {
        const SceneEntity &fe = builder.film;
        const ParameterDictionary &d = fe.parameters;
        if (fe.name != "rgb" && fe.name != "gbuffer") ErrorExit(&fe.loc, "film \"%s\" is not implemented by the engine", fe.name.ToString());
        film.type = fe.name == "rgb" ? BPT_FILM_RGB : BPT_FILM_GBUFFER;
        film.xres = d.GetOneInt("xresolution", 1280); film.yres = d.GetOneInt("yresolution", 720);
        if (Options->quickRender) { film.xres = std::max(1, film.xres / 4); film.yres = std::max(1, film.yres / 4); }
        std::vector<int> pb = d.GetIntArray("pixelbounds");
        if (pb.size() == 4) for (int k = 0; k < 4; ++k) film.pixel_bounds[k] = pb[k];
        std::vector<Float> cr = d.GetFloatArray("cropwindow");
        if (cr.size() == 4) {   // film.cpp:138-160
            Bounds2f crop;
            crop.pMin.x = Clamp(std::min(cr[0], cr[1]), 0.f, 1.f); crop.pMax.x = Clamp(std::max(cr[0], cr[1]), 0.f, 1.f);
            crop.pMin.y = Clamp(std::min(cr[2], cr[3]), 0.f, 1.f); crop.pMax.y = Clamp(std::max(cr[2], cr[3]), 0.f, 1.f);
            film.pixel_bounds[0] = (int)pstd::ceil(film.xres * crop.pMin.x); film.pixel_bounds[1] = (int)pstd::ceil(film.xres * crop.pMax.x);
            film.pixel_bounds[2] = (int)pstd::ceil(film.yres * crop.pMin.y); film.pixel_bounds[3] = (int)pstd::ceil(film.yres * crop.pMax.y);
        }
        std::string csys = d.GetOneString("coordinatesystem", "camera");
        film.coordinate_system = csys == "world" ? 1 : 0;
        film.iso = d.GetOneFloat("iso", 100.f); film.white_balance = d.GetOneFloat("whitebalance", 0.f); film.exposure_time = -1;
        film.max_component_value = d.GetOneFloat("maxcomponentvalue", -1.f); film.diagonal = d.GetOneFloat("diagonal", 35.f);
        if (d.GetOneString("sensor", "cie1931") != "cie1931") ErrorExit(&fe.loc, "only the cie1931 sensor is implemented by the engine");
        outName = !Options->imageFile.empty() ? Options->imageFile : d.GetOneString("filename", "pbrt.pfm");
        Check(api, e, api.bpt_scene_set_film(e, &film), "bpt_scene_set_film");
    }
    {   // filter (filters.cpp *::Create)
        const SceneEntity &fe = builder.filter;
        const ParameterDictionary &d = fe.parameters;
        bpt_filter f; std::memset(&f, 0, sizeof(f));
        if (fe.name == "gaussian") { f.type = BPT_FILTER_GAUSSIAN; f.xradius = d.GetOneFloat("xradius", 1.5f); f.yradius = d.GetOneFloat("yradius", 1.5f); f.sigma = d.GetOneFloat("sigma", 0.5f); }
        else if (fe.name == "box") { f.type = BPT_FILTER_BOX; f.xradius = d.GetOneFloat("xradius", 0.5f); f.yradius = d.GetOneFloat("yradius", 0.5f); }
        else if (fe.name == "mitchell") { f.type = BPT_FILTER_MITCHELL; f.xradius = d.GetOneFloat("xradius", 2.f); f.yradius = d.GetOneFloat("yradius", 2.f); f.b = d.GetOneFloat("B", 1.f / 3.f); f.c = d.GetOneFloat("C", 1.f / 3.f); }
        else if (fe.name == "sinc") { f.type = BPT_FILTER_SINC; f.xradius = d.GetOneFloat("xradius", 4.f); f.yradius = d.GetOneFloat("yradius", 4.f); f.tau = d.GetOneFloat("tau", 3.f); }
        else if (fe.name == "triangle") { f.type = BPT_FILTER_TRIANGLE; f.xradius = d.GetOneFloat("xradius", 2.f); f.yradius = d.GetOneFloat("yradius", 2.f); }
        else ErrorExit(&fe.loc, "pixel filter \"%s\" unknown", fe.name.ToString());
        Check(api, e, api.bpt_scene_set_filter(e, &f), "bpt_scene_set_filter");
    }
    int spp = 16;
    {   // sampler (samplers.cpp *::Create: --spp / --seed override the file)
        const SceneEntity &se = builder.sampler;
        const ParameterDictionary &d = se.parameters;
        bpt_sampler s; std::memset(&s, 0, sizeof(s));
        static const std::map<std::string, int> types = {{"zsobol", BPT_SAMPLER_ZSOBOL}, {"independent", BPT_SAMPLER_INDEPENDENT}, {"halton", BPT_SAMPLER_HALTON},
            {"sobol", BPT_SAMPLER_SOBOL}, {"paddedsobol", BPT_SAMPLER_PADDED_SOBOL}, {"stratified", BPT_SAMPLER_STRATIFIED}, {"pmj02bn", BPT_SAMPLER_PMJ02BN}};
        auto it = types.find(se.name.ToString());
        if (it == types.end()) ErrorExit(&se.loc, "sampler \"%s\" unknown", se.name.ToString());
        s.type = it->second;
        s.spp = d.GetOneInt("pixelsamples", 16);
        if (Options->pixelSamples) s.spp = *Options->pixelSamples;
        s.seed = d.GetOneInt("seed", Options->seed);
        std::string r = d.GetOneString("randomization", s.type == BPT_SAMPLER_HALTON ? "permutedigits" : "fastowen");
        s.randomize = r == "none" ? BPT_RANDOMIZE_NONE : r == "permutedigits" ? BPT_RANDOMIZE_PERMUTE_DIGITS : r == "owen" ? BPT_RANDOMIZE_OWEN : BPT_RANDOMIZE_FAST_OWEN;
        if (s.type == BPT_SAMPLER_STRATIFIED) {   // StratifiedSampler::Create (samplers.cpp:297-319)
            s.jitter = d.GetOneBool("jitter", true);
            s.x_samples = d.GetOneInt("xsamples", 4); s.y_samples = d.GetOneInt("ysamples", 4);
            if (Options->pixelSamples) {
                int nSamples = *Options->pixelSamples, div = (int)std::sqrt(nSamples);
                while (nSamples % div) --div;
                s.x_samples = nSamples / div; s.y_samples = nSamples / s.x_samples;
            }
            s.spp = s.x_samples * s.y_samples;
        }
        spp = s.spp;
        Check(api, e, api.bpt_scene_set_sampler(e, &s), "bpt_scene_set_sampler");
    }
    {   // integrator (cpu/integrators.cpp VolPathIntegrator::Create)
        const SceneEntity &ie = scene.integrator;
        const ParameterDictionary &d = ie.parameters;
        if (ie.name != "volpath") ErrorExit(&ie.loc, "integrator \"%s\": the engine implements volpath", ie.name.ToString());
        bpt_integrator i; std::memset(&i, 0, sizeof(i));
        i.max_depth = d.GetOneInt("maxdepth", 5); i.regularize = d.GetOneBool("regularize", false);
        std::string ls = d.GetOneString("lightsampler", "bvh");
        i.light_sampler = ls == "uniform" ? BPT_LIGHTSAMPLER_UNIFORM : ls == "power" ? BPT_LIGHTSAMPLER_POWER : ls == "exhaustive" ? BPT_LIGHTSAMPLER_EXHAUSTIVE : BPT_LIGHTSAMPLER_BVH;
        Check(api, e, api.bpt_scene_set_integrator(e, &i), "bpt_scene_set_integrator");
    }
    // textures (FloatImageTexture::Create / SpectrumImageTexture::Create, textures.cpp:410-500), recorded at Texture; the image as
    // Image::Read returns it (MIPMap::CreateFromFile, util/mipmap.cpp:349-381)
    TexIds tx;
    std::vector<bpt_texture> texs;
    std::vector<std::unique_ptr<std::vector<uint8_t>>> texelStore;
    // TextureMapping2D::Create (textures.cpp:49-73)
    auto SetMapping = [&](const TextureRecord &tr, bpt_texture *t) {
        const ParameterDictionary &d = tr.params;
        std::string type = d.GetOneString("mapping", "uv");
        t->su = d.GetOneFloat("uscale", 1.f); t->sv = d.GetOneFloat("vscale", 1.f); t->du = d.GetOneFloat("udelta", 0.f); t->dv = d.GetOneFloat("vdelta", 0.f);
        t->v1[0] = 1; t->v2[1] = 1;
        if (type == "uv") { t->mapping = BPT_TEXMAP_UV; return; }
        if (type == "spherical") t->mapping = BPT_TEXMAP_SPHERICAL;
        else if (type == "cylindrical") t->mapping = BPT_TEXMAP_CYLINDRICAL;
        else if (type == "planar") {
            t->mapping = BPT_TEXMAP_PLANAR;
            Vector3f v1 = d.GetOneVector3f("v1", Vector3f(1, 0, 0)), v2 = d.GetOneVector3f("v2", Vector3f(0, 1, 0));
            for (int k = 0; k < 3; ++k) { t->v1[k] = v1[k]; t->v2[k] = v2[k]; }
        } else ErrorExit(&tr.loc, "2D texture mapping \"%s\" unknown", type);
        t->world_from_texture = ToB(tr.renderFromTexture.GetMatrix()); t->texture_from_world = ToB(tr.renderFromTexture.GetInverseMatrix());
    };
    for (const TextureRecord &tr : builder.textures) {
        const ParameterDictionary &d = tr.params;
        const bool spec = tr.type == "spectrum";
        // a float / spectrum texture parameter of a texture: a named texture, or a constant in its own constant-texture entry
        // (TextureParameterDictionary::GetFloatTexture / GetSpectrumTexture, paramdict.cpp)
        auto floatChild = [&](const std::string &name, float dflt) -> int {
            const ParsedParameter *p = FindParam(d, name);
            if (p && p->type == "texture") {
                p->lookedUp = true;
                auto it = p->strings.empty() ? tx.floats.end() : tx.floats.find(p->strings[0]);
                if (it == tx.floats.end()) ErrorExit(&tr.loc, "parameter \"%s\": unknown float texture", name);
                return it->second;
            }
            bpt_texture c; std::memset(&c, 0, sizeof(c));
            c.kind = BPT_TEXTURE_CONSTANT; c.value_f = d.GetOneFloat(name, dflt);
            texs.push_back(c); return (int)texs.size();
        };
        auto spectrumChild = [&](const std::string &name, float dflt) -> int {
            const ParsedParameter *p = FindParam(d, name);
            if (p && p->type == "texture") {
                p->lookedUp = true;
                auto it = p->strings.empty() ? tx.spectra.end() : tx.spectra.find(p->strings[0]);
                if (it == tx.spectra.end()) ErrorExit(&tr.loc, "parameter \"%s\": unknown spectrum texture", name);
                return it->second;
            }
            bpt_texture c; std::memset(&c, 0, sizeof(c));
            c.kind = BPT_TEXTURE_CONSTANT; c.is_spectrum = 1; c.value_s = GetSpectrum(d, name, st);
            if (c.value_s.kind == BPT_SPEC_NONE) { c.value_s.kind = BPT_SPEC_CONSTANT; c.value_s.value = dflt; }
            texs.push_back(c); return (int)texs.size();
        };
        if (tr.texname == "constant" || tr.texname == "scale" || tr.texname == "mix" || tr.texname == "checkerboard") {
            bpt_texture t; std::memset(&t, 0, sizeof(t));
            t.is_spectrum = spec;
            if (tr.texname == "checkerboard") {
                if (d.GetOneInt("dimension", 2) != 2) ErrorExit(&tr.loc, "only the two-dimensional checkerboard is implemented by the engine");
                t.kind = BPT_TEXTURE_CHECKERBOARD;
                SetMapping(tr, &t);
                t.tex1 = spec ? spectrumChild("tex1", 1.f) : floatChild("tex1", 1.f);
                t.tex2 = spec ? spectrumChild("tex2", 0.f) : floatChild("tex2", 0.f);
            } else
            if (tr.texname == "constant") {
                t.kind = BPT_TEXTURE_CONSTANT;
                if (spec) { t.value_s = GetSpectrum(d, "value", st); if (t.value_s.kind == BPT_SPEC_NONE) { t.value_s.kind = BPT_SPEC_CONSTANT; t.value_s.value = 1.f; } }
                else t.value_f = d.GetOneFloat("value", 1.f);
            } else if (tr.texname == "scale") {
                t.kind = BPT_TEXTURE_SCALE;
                t.tex1 = spec ? spectrumChild("tex", 1.f) : floatChild("tex", 1.f);
                t.tex2 = floatChild("scale", 1.f);
            } else {
                t.kind = BPT_TEXTURE_MIX;
                t.tex1 = spec ? spectrumChild("tex1", 0.f) : floatChild("tex1", 0.f);
                t.tex2 = spec ? spectrumChild("tex2", 1.f) : floatChild("tex2", 1.f);
                t.tex3 = floatChild("amount", 0.5f);
            }
            texs.push_back(t);
            (spec ? tx.spectra : tx.floats)[tr.name] = (int)texs.size();
            continue;
        }
        if (tr.texname != "imagemap") ErrorExit(&tr.loc, "texture class \"%s\" is not implemented by the engine (imagemap, constant, scale, mix, checkerboard)", tr.texname);
        bpt_texture t; std::memset(&t, 0, sizeof(t));
        SetMapping(tr, &t);
        t.max_anisotropy = d.GetOneFloat("maxanisotropy", 8.f);
        pstd::optional<FilterFunction> ff = ParseFilter(d.GetOneString("filter", "bilinear"));
        if (!ff) ErrorExit(&tr.loc, "filter function unknown");
        t.filter = (int)*ff;
        pstd::optional<WrapMode> wm = ParseWrapMode(d.GetOneString("wrap", "repeat").c_str());
        if (!wm || *wm == WrapMode::OctahedralSphere) ErrorExit(&tr.loc, "wrap mode unknown");
        t.wrap = (int)*wm;
        t.scale = d.GetOneFloat("scale", 1.f); t.invert = d.GetOneBool("invert", false);
        std::string filename = ResolveFilename(d.GetOneString("filename", ""));
        std::string enc = d.GetOneString("encoding", HasExtension(filename, "png") ? "sRGB" : "linear");
        if (enc != "sRGB" && enc != "linear") ErrorExit(&tr.loc, "texture encoding \"%s\" is not implemented by the engine", enc);
        t.encoding = enc == "sRGB" ? BPT_ENCODING_SRGB : BPT_ENCODING_LINEAR;
        ImageAndMetadata im = Image::Read(filename, Allocator(), ColorEncoding::Get(enc, Allocator()));
        Image &image = im.image;
        if (image.NChannels() != 1) {
            ImageChannelDesc rgba = image.GetChannelDesc({"R", "G", "B", "A"}), rgb = image.GetChannelDesc({"R", "G", "B"});
            if (rgba) image = image.SelectChannels(rgba);
            else if (rgb) image = image.SelectChannels(rgb);
            else ErrorExit(&tr.loc, "%s: image doesn't have R, G, and B channels", filename);
        }
        t.xres = image.Resolution().x; t.yres = image.Resolution().y; t.n_channels = image.NChannels();
        const size_t count = (size_t)t.xres * t.yres * t.n_channels;
        if (image.Format() == PixelFormat::U256) {
            t.format = BPT_IMAGE_U8;
            texelStore.emplace_back(new std::vector<uint8_t>((const uint8_t *)image.RawPointer({0, 0}), (const uint8_t *)image.RawPointer({0, 0}) + count));
        } else if (image.Format() == PixelFormat::Float) {
            t.format = BPT_IMAGE_F32;
            texelStore.emplace_back(new std::vector<uint8_t>((const uint8_t *)image.RawPointer({0, 0}), (const uint8_t *)image.RawPointer({0, 0}) + count * sizeof(float)));
        } else ErrorExit(&tr.loc, "%s: half-precision texture images are not implemented by the engine", filename);
        t.pixels = texelStore.back()->data();
        t.is_spectrum = spec;
        texs.push_back(t);
        (tr.type == "float" ? tx.floats : tx.spectra)[tr.name] = (int)texs.size();
    }
    // normal maps (BasicScene::startLoadingNormalMaps, scene.cpp:886-910): the file read with the linear encoding, R G B selected
    std::map<std::string, int> normalMapIds;
    tx.normalMap = [&](const std::string &filename) -> int {
        auto it = normalMapIds.find(filename);
        if (it != normalMapIds.end()) return it->second;
        ImageAndMetadata im = Image::Read(filename, Allocator(), ColorEncoding::Linear);
        ImageChannelDesc rgb = im.image.GetChannelDesc({"R", "G", "B"});
        if (!rgb) ErrorExit("%s: normal map image must contain R, G, and B channels", filename);
        Image image = im.image.SelectChannels(rgb);
        bpt_texture t; std::memset(&t, 0, sizeof(t));
        t.kind = BPT_TEXTURE_NORMALMAP; t.encoding = BPT_ENCODING_LINEAR; t.wrap = BPT_TEXWRAP_REPEAT; t.filter = BPT_TEXFILTER_BILINEAR; t.scale = 1; t.su = t.sv = 1;
        t.xres = image.Resolution().x; t.yres = image.Resolution().y; t.n_channels = 3;
        const size_t count = (size_t)t.xres * t.yres * 3;
        if (image.Format() == PixelFormat::U256) { t.format = BPT_IMAGE_U8; texelStore.emplace_back(new std::vector<uint8_t>((const uint8_t *)image.RawPointer({0, 0}), (const uint8_t *)image.RawPointer({0, 0}) + count)); }
        else if (image.Format() == PixelFormat::Float) { t.format = BPT_IMAGE_F32; texelStore.emplace_back(new std::vector<uint8_t>((const uint8_t *)image.RawPointer({0, 0}), (const uint8_t *)image.RawPointer({0, 0}) + count * sizeof(float))); }
        else ErrorExit("%s: half-precision images are not implemented by the engine", filename);
        t.pixels = texelStore.back()->data();
        texs.push_back(t);
        return normalMapIds[filename] = (int)texs.size();
    };
    // materials: the indexed ones (scene.materials) followed by the named ones (scene.namedMaterials)
    std::vector<bpt_material> mats;
    std::map<std::string, int> namedIndex;
    for (const SceneEntity &me : scene.materials) mats.push_back(MakeMaterial(me.name.ToString(), me.parameters, st, tx, &me.loc));
    for (const auto &nm : scene.namedMaterials) {
        namedIndex[nm.first] = (int)mats.size();
        mats.push_back(MakeMaterial(nm.second.parameters.GetOneString("type", ""), nm.second.parameters, st, tx, &nm.second.loc));
    }
    if (!texs.empty()) Check(api, e, api.bpt_scene_set_textures(e, texs.data(), (int)texs.size()), "bpt_scene_set_textures");
    Check(api, e, api.bpt_scene_set_materials(e, mats.data(), (int)mats.size()), "bpt_scene_set_materials");
    // area lights (lights.cpp DiffuseAreaLight::Create)
    std::vector<bpt_area_light> als;
    for (const SceneEntity &ae : scene.areaLights) {
        if (ae.name != "diffuse") ErrorExit(&ae.loc, "area light \"%s\" unknown", ae.name.ToString());
        const ParameterDictionary &d = ae.parameters;
        bpt_area_light a; std::memset(&a, 0, sizeof(a));
        a.L = GetSpectrum(d, "L", st); a.scale = d.GetOneFloat("scale", 1.f); a.two_sided = d.GetOneBool("twosided", false); a.power = d.GetOneFloat("power", -1.f);
        if (!d.GetOneString("filename", "").empty()) ErrorExit(&ae.loc, "image area lights are not implemented by the engine");
        als.push_back(a);
    }
    Check(api, e, api.bpt_scene_set_area_lights(e, als.data(), (int)als.size()), "bpt_scene_set_area_lights");
    // media (media.cpp HomogeneousMedium::Create / GridMedium::Create), recorded at MakeNamedMedium
    std::map<std::string, int> mediumIndex;
    for (const MediumRecord &mr : builder.media) {
        const ParameterDictionary &d = mr.params;
        bpt_medium m; std::memset(&m, 0, sizeof(m));
        std::string type = d.GetOneString("type", "");
        if (type == "uniformgrid") m.type = BPT_MEDIUM_UNIFORMGRID;
        else if (type == "homogeneous") m.type = BPT_MEDIUM_HOMOGENEOUS;
        else if (type == "rgbgrid") m.type = BPT_MEDIUM_RGBGRID;
        else ErrorExit("medium type \"%s\" is not implemented by the engine", type);
        if (!d.GetOneString("preset", "").empty()) ErrorExit("medium presets are not implemented by the engine");
        if (m.type == BPT_MEDIUM_RGBGRID) {   // RGBGridMedium::Create (media.cpp:380-453)
            m.nx = d.GetOneInt("nx", 1); m.ny = d.GetOneInt("ny", 1); m.nz = d.GetOneInt("nz", 1);
            auto grid = [&](const char *name) -> const float * {
                std::vector<RGB> v = d.GetRGBArray(name);
                if (v.empty()) return nullptr;
                if (v.size() != (size_t)m.nx * m.ny * m.nz) ErrorExit("rgbgrid medium: \"%s\" needs nx*ny*nz RGB values", name);
                std::vector<float> f;
                for (RGB c : v) { f.push_back(c.r); f.push_back(c.g); f.push_back(c.b); }
                return st.F(std::move(f));
            };
            m.sigma_a_rgb = grid("sigma_a"); m.sigma_s_rgb = grid("sigma_s"); m.Le_rgb = grid("Le");
            m.scale = d.GetOneFloat("scale", 1.f); m.g = d.GetOneFloat("g", 0.f); m.Le_scale = d.GetOneFloat("Lescale", 1.f);
            m.world_from_medium = ToB(mr.renderFromMedium.GetMatrix()); m.medium_from_world = ToB(mr.renderFromMedium.GetInverseMatrix());
            Point3f p0 = d.GetOnePoint3f("p0", Point3f(0, 0, 0)), p1 = d.GetOnePoint3f("p1", Point3f(1, 1, 1));
            m.bounds[0] = p0.x; m.bounds[1] = p0.y; m.bounds[2] = p0.z; m.bounds[3] = p1.x; m.bounds[4] = p1.y; m.bounds[5] = p1.z;
            mediumIndex[mr.name] = (int)mediumIndex.size();
            Check(api, e, api.bpt_scene_add_medium(e, &m), "bpt_scene_add_medium");
            continue;
        }
        m.sigma_a = GetSpectrum(d, "sigma_a", st); m.sigma_s = GetSpectrum(d, "sigma_s", st); m.Le = GetSpectrum(d, "Le", st);
        m.scale = d.GetOneFloat("scale", 1.f); m.g = d.GetOneFloat("g", 0.f);
        m.Le_scale = m.type == BPT_MEDIUM_HOMOGENEOUS ? d.GetOneFloat("Lescale", 1.f) : 1.f;
        if (!d.GetFloatArray("temperature").empty()) ErrorExit("temperature grids are not implemented by the engine");
        m.world_from_medium = ToB(mr.renderFromMedium.GetMatrix()); m.medium_from_world = ToB(mr.renderFromMedium.GetInverseMatrix());
        m.bounds[3] = m.bounds[4] = m.bounds[5] = 1;
        if (m.type == BPT_MEDIUM_UNIFORMGRID) {
            std::vector<Float> dens = d.GetFloatArray("density");
            m.nx = d.GetOneInt("nx", 1); m.ny = d.GetOneInt("ny", 1); m.nz = d.GetOneInt("nz", 1);
            if (dens.empty() || (size_t)m.nx * m.ny * m.nz != dens.size()) ErrorExit("uniformgrid medium: density count != nx*ny*nz");
            m.density = st.F(std::vector<float>(dens.begin(), dens.end()));
            std::vector<Float> les = d.GetFloatArray("Lescale");
            if (!les.empty()) {
                if (les.size() != dens.size()) ErrorExit("uniformgrid medium: \"Lescale\" count != nx*ny*nz");
                m.Le_scale_grid = st.F(std::vector<float>(les.begin(), les.end()));
            }
            Point3f p0 = d.GetOnePoint3f("p0", Point3f(0, 0, 0)), p1 = d.GetOnePoint3f("p1", Point3f(1, 1, 1));
            m.bounds[0] = p0.x; m.bounds[1] = p0.y; m.bounds[2] = p0.z; m.bounds[3] = p1.x; m.bounds[4] = p1.y; m.bounds[5] = p1.z;
        }
        mediumIndex[mr.name] = (int)mediumIndex.size();
        Check(api, e, api.bpt_scene_add_medium(e, &m), "bpt_scene_add_medium");
    }
    auto mediumOf = [&](const std::string &name) { if (name.empty()) return -1; auto it = mediumIndex.find(name); if (it == mediumIndex.end()) ErrorExit("medium \"%s\" undefined", name); return it->second; };
    if (!builder.camera.medium.empty()) Check(api, e, api.bpt_scene_set_camera_medium(e, mediumOf(builder.camera.medium)), "bpt_scene_set_camera_medium");

    // shapes (shapes.cpp Triangle::CreateMesh / util/mesh.cpp TriQuadMesh::ReadPLY), ShapeSceneEntity -> bpt_mesh
    auto addShape = [&](const ShapeSceneEntity &sh, int objectId) {
        const ParameterDictionary &d = sh.parameters;
        std::vector<float> P, N, UV; std::vector<int> idx;
        if (sh.name == "sphere" || sh.name == "disk" || sh.name == "cylinder") {
            // Sphere / Disk / Cylinder::Create (shapes.cpp:72-83, 107-117, 135-146)
            bpt_quadric q; std::memset(&q, 0, sizeof(q));
            q.radius = d.GetOneFloat("radius", 1.f); q.phimax = d.GetOneFloat("phimax", 360.f);
            if (sh.name == "sphere") { q.type = BPT_SHAPE_SPHERE; q.zmin = d.GetOneFloat("zmin", -q.radius); q.zmax = d.GetOneFloat("zmax", q.radius); }
            else if (sh.name == "cylinder") { q.type = BPT_SHAPE_CYLINDER; q.zmin = d.GetOneFloat("zmin", -1.f); q.zmax = d.GetOneFloat("zmax", 1.f); }
            else { q.type = BPT_SHAPE_DISK; q.height = d.GetOneFloat("height", 0.f); q.innerradius = d.GetOneFloat("innerradius", 0.f); }
            if (!d.GetTexture("alpha").empty() || d.GetOneFloat("alpha", 1.f) < 1.f) ErrorExit(&sh.loc, "alpha on a %s is not implemented by the engine", sh.name.ToString());
            q.world_from_object = ToB(sh.renderFromObject->GetMatrix()); q.object_from_world = ToB(sh.renderFromObject->GetInverseMatrix());
            q.reverse_orientation = sh.reverseOrientation;
            int qmi = sh.materialName.empty() ? sh.materialIndex : namedIndex.at(sh.materialName);
            q.material = mats[qmi].type == BPT_MAT_INTERFACE ? -1 : qmi;
            q.area_light = sh.lightIndex;
            q.medium_inside = mediumOf(sh.insideMedium); q.medium_outside = mediumOf(sh.outsideMedium);
            q.object_id = objectId;
            Check(api, e, api.bpt_scene_add_quadric(e, &q), "bpt_scene_add_quadric");
            return;
        }
        std::vector<int> quadIdx;   // bilinear patches: "bilinearmesh", or the quad faces of a PLY file (a second mesh over the same vertices)
        if (sh.name == "bilinearmesh") {   // BilinearPatch::CreateMesh (shapes.cpp:910-999)
            for (Point3f p : d.GetPoint3fArray("P")) { P.push_back(p.x); P.push_back(p.y); P.push_back(p.z); }
            quadIdx = d.GetIntArray("indices");
            if (quadIdx.empty() && P.size() == 12) quadIdx = {0, 1, 2, 3};
            while (quadIdx.size() % 4) quadIdx.pop_back();
            for (Normal3f n : d.GetNormal3fArray("N")) { N.push_back(n.x); N.push_back(n.y); N.push_back(n.z); }
            for (Point2f u : d.GetPoint2fArray("uv")) { UV.push_back(u.x); UV.push_back(u.y); }
            if (!d.GetOneString("emissionfilename", "").empty()) ErrorExit(&sh.loc, "\"emissionfilename\" is not implemented by the engine");
        } else if (sh.name == "trianglemesh") {
            for (Point3f p : d.GetPoint3fArray("P")) { P.push_back(p.x); P.push_back(p.y); P.push_back(p.z); }
            idx = d.GetIntArray("indices");
            if (idx.empty() && P.size() == 9) idx = {0, 1, 2};
            for (Normal3f n : d.GetNormal3fArray("N")) { N.push_back(n.x); N.push_back(n.y); N.push_back(n.z); }
            for (Point2f u : d.GetPoint2fArray("uv")) { UV.push_back(u.x); UV.push_back(u.y); }
        } else if (sh.name == "plymesh") {
            std::string fn = ResolveFilename(d.GetOneString("filename", ""));
            TriQuadMesh mesh = TriQuadMesh::ReadPLY(fn);
            quadIdx.assign(mesh.quadIndices.begin(), mesh.quadIndices.end());
            for (Point3f p : mesh.p) { P.push_back(p.x); P.push_back(p.y); P.push_back(p.z); }
            for (Normal3f n : mesh.n) { N.push_back(n.x); N.push_back(n.y); N.push_back(n.z); }
            for (Point2f u : mesh.uv) { UV.push_back(u.x); UV.push_back(u.y); }
            idx.assign(mesh.triIndices.begin(), mesh.triIndices.end());
        } else ErrorExit(&sh.loc, "shape \"%s\" is not implemented by the engine (triangle meshes, bilinear patches, sphere, disk, cylinder)", sh.name.ToString());
        if (!d.GetTexture("alpha").empty()) ErrorExit(&sh.loc, "texture-valued alpha is not implemented by the engine");
        bpt_mesh m; std::memset(&m, 0, sizeof(m));
        if (Float alpha = d.GetOneFloat("alpha", 1.f); alpha < 1.f) { m.has_alpha = 1; m.alpha = alpha; }   // getAlphaTexture (scene.cpp)
        m.n_vertices = (int)P.size() / 3; m.n_triangles = (int)idx.size() / 3;
        const bool haveTris = !idx.empty();
        m.P = st.F(std::move(P)); m.N = N.empty() ? nullptr : st.F(std::move(N)); m.UV = UV.empty() ? nullptr : st.F(std::move(UV)); m.indices = st.I(std::move(idx));
        m.world_from_object = ToB(sh.renderFromObject->GetMatrix()); m.object_from_world = ToB(sh.renderFromObject->GetInverseMatrix());
        m.reverse_orientation = sh.reverseOrientation;
        int mi = sh.materialName.empty() ? sh.materialIndex : namedIndex.at(sh.materialName);
        m.material = mats[mi].type == BPT_MAT_INTERFACE ? -1 : mi;
        m.area_light = sh.lightIndex;
        m.medium_inside = mediumOf(sh.insideMedium); m.medium_outside = mediumOf(sh.outsideMedium);
        m.object_id = objectId;
        if (haveTris || quadIdx.empty()) Check(api, e, api.bpt_scene_add_mesh(e, &m), "bpt_scene_add_mesh");
        if (!quadIdx.empty()) {   // Shape::Create: the triangles of a PLY file first, then its quads (shapes.cpp:1435-1461)
            if (m.has_alpha) ErrorExit(&sh.loc, "alpha on bilinear patches is not implemented by the engine");
            m.n_triangles = (int)quadIdx.size() / 4; m.indices = st.I(std::move(quadIdx));
            Check(api, e, api.bpt_scene_add_bilinear_mesh(e, &m), "bpt_scene_add_bilinear_mesh");
        }
    };
    if (!scene.animatedShapes.empty()) ErrorExit("animated shapes are not implemented by the engine");
    for (const ShapeSceneEntity &sh : scene.shapes) addShape(sh, -1);
    std::map<std::string, int> objectIds;
    for (const auto &def : scene.instanceDefinitions) {
        int id = (int)objectIds.size();
        objectIds[def.first.ToString()] = id;
        if (!def.second->animatedShapes.empty()) ErrorExit("animated shapes are not implemented by the engine");
        for (const ShapeSceneEntity &sh : def.second->shapes) addShape(sh, id);
    }
    for (const InstanceSceneEntity &in : scene.instances) {
        if (in.renderFromInstanceAnim) ErrorExit(&in.loc, "animated instances are not implemented by the engine");
        auto it = objectIds.find(in.name.ToString());
        if (it == objectIds.end()) ErrorExit(&in.loc, "instance of an undefined object");
        bpt_instance bi; std::memset(&bi, 0, sizeof(bi));
        bi.object_id = it->second;
        bi.world_from_instance = ToB(in.renderFromInstance->GetMatrix()); bi.instance_from_world = ToB(in.renderFromInstance->GetInverseMatrix());
        Check(api, e, api.bpt_scene_add_instance(e, &bi), "bpt_scene_add_instance");
    }
    // lights (lights.cpp Light::Create "infinite"), recorded at LightSource
    for (const LightRecord &lr : builder.lights) {
        if (lr.type != "infinite" && lr.type != "point" && lr.type != "spot" && lr.type != "distant")
            ErrorExit("light \"%s\" is not implemented by the engine (infinite, point, spot, distant, diffuse area)", lr.type);
        const ParameterDictionary &d = lr.params;
        bpt_light l; std::memset(&l, 0, sizeof(l));
        l.world_from_light = ToB(lr.renderFromLight.GetMatrix()); l.light_from_world = ToB(lr.renderFromLight.GetInverseMatrix());
        l.scale = d.GetOneFloat("scale", 1.f);
        if (lr.type != "infinite") {
            // the parameters PointLight / SpotLight / DistantLight::Create read (lights.cpp:192-213, 1433-1464, 246-276)
            l.type = lr.type == "point" ? BPT_LIGHT_POINT : lr.type == "spot" ? BPT_LIGHT_SPOT : BPT_LIGHT_DISTANT;
            l.L = GetSpectrum(d, lr.type == "distant" ? "L" : "I", st);
            Point3f from = d.GetOnePoint3f("from", Point3f(0, 0, 0));
            Point3f to = lr.type == "point" ? Point3f(0, 0, 1) : d.GetOnePoint3f("to", Point3f(0, 0, 1));
            for (int k = 0; k < 3; ++k) { l.from[k] = from[k]; l.to[k] = to[k]; }
            l.power = lr.type == "distant" ? -1.f : d.GetOneFloat("power", -1.f);
            l.illuminance = lr.type == "distant" ? d.GetOneFloat("illuminance", -1.f) : -1.f;
            if (lr.type == "spot") { l.cone_angle = d.GetOneFloat("coneangle", 30.f); l.cone_delta = d.GetOneFloat("conedeltaangle", 5.f); }
            Check(api, e, api.bpt_scene_add_light(e, &l), "bpt_scene_add_light");
            continue;
        }
        l.L = GetSpectrum(d, "L", st); l.illuminance = d.GetOneFloat("illuminance", -1.f);
        std::string fn = ResolveFilename(d.GetOneString("filename", ""));
        if (fn.empty()) l.type = BPT_LIGHT_UNIFORM_INFINITE;
        else {
            l.type = BPT_LIGHT_IMAGE_INFINITE;
            ImageAndMetadata im = Image::Read(fn);
            ImageChannelDesc rgb = im.image.GetChannelDesc({"R", "G", "B"});
            if (!rgb || im.image.Resolution().x != im.image.Resolution().y) ErrorExit("%s: square RGB environment map expected", fn);
            int res = im.image.Resolution().x;
            std::vector<float> px(3 * (size_t)res * res);
            for (int y = 0; y < res; ++y)
                for (int x = 0; x < res; ++x) {
                    ImageChannelValues v = im.image.GetChannels({x, y}, rgb);
                    for (int c = 0; c < 3; ++c) px[3 * ((size_t)y * res + x) + c] = v[c];
                }
            l.image_rgb = st.F(std::move(px)); l.image_res = res;
        }
        Check(api, e, api.bpt_scene_add_light(e, &l), "bpt_scene_add_light");
    }
    Check(api, e, api.bpt_scene_commit(e), "bpt_scene_commit");
    if (getenv("BPT_RECORD_OUT")) { api.bpt_destroy(e); return; }   // recorder run: the call stream has been written

    // render all samples, like WavefrontPathIntegrator::Render()'s sample loop (wavefront/integrator.cpp:290-493)
    Check(api, e, api.bpt_render(e, 0, spp, nullptr, 0), "bpt_render");
    if (film.type == BPT_FILM_RGB) {
        std::vector<float> rgb(3 * (size_t)film.xres * film.yres);
        Check(api, e, api.bpt_film_resolve(e, rgb.data()), "bpt_film_resolve");   // == RGBFilm::GetPixelRGB per pixel (film.h:258-274)
        Image image(PixelFormat::Float, {film.xres, film.yres}, {"R", "G", "B"});
        for (int y = 0; y < film.yres; ++y)
            for (int x = 0; x < film.xres; ++x)
                for (int c = 0; c < 3; ++c) image.SetChannel({x, y}, c, rgb[3 * ((size_t)y * film.xres + x) + c]);
        ImageMetadata md;
        if (!image.Write(outName, md)) ErrorExit("%s: could not write", outName);
    } else {
        static const char *names[25] = {"R", "G", "B", "Albedo.R", "Albedo.G", "Albedo.B", "P.X", "P.Y", "P.Z", "dzdx", "dzdy", "N.X", "N.Y", "N.Z", "Ns.X", "Ns.Y",
                                        "Ns.Z", "u", "v", "Variance.R", "Variance.G", "Variance.B", "RelativeVariance.R", "RelativeVariance.G", "RelativeVariance.B"};
        std::vector<float> g(25 * (size_t)film.xres * film.yres);
        Check(api, e, api.bpt_film_resolve_gbuffer(e, g.data()), "bpt_film_resolve_gbuffer");
        Image image(PixelFormat::Float, {film.xres, film.yres}, std::vector<std::string>(names, names + 25));
        for (int y = 0; y < film.yres; ++y)
            for (int x = 0; x < film.xres; ++x)
                for (int c = 0; c < 25; ++c) image.SetChannel({x, y}, c, g[25 * ((size_t)y * film.xres + x) + c]);
        ImageMetadata md;
        if (!image.Write(outName, md)) ErrorExit("%s: could not write", outName);
    }
    api.bpt_destroy(e);
}

}  // namespace pbrt

int main(int argc, char *argv[]) {
    using namespace pbrt;
    std::vector<std::string> args = GetCommandLineArguments(argv);
    PBRTOptions options;
    std::vector<std::string> filenames;
    auto onError = [](const std::string &err) { fprintf(stderr, "pbrt_b200: %s\n", err.c_str()); exit(1); };
    for (auto iter = args.begin(); iter != args.end(); ++iter) {
        if ((*iter)[0] != '-') { filenames.push_back(*iter); continue; }
        if (ParseArg(&iter, args.end(), "seed", &options.seed, onError) || ParseArg(&iter, args.end(), "outfile", &options.imageFile, onError) ||
            ParseArg(&iter, args.end(), "quiet", &options.quiet, onError) || ParseArg(&iter, args.end(), "nthreads", &options.nThreads, onError) ||
            ParseArg(&iter, args.end(), "quick", &options.quickRender, onError)) continue;
        int spp;
        if (ParseArg(&iter, args.end(), "spp", &spp, onError)) { options.pixelSamples = spp; continue; }
        int dev;
        if (ParseArg(&iter, args.end(), "gpu-device", &dev, onError)) { options.gpuDevice = dev; continue; }
        onError("unknown argument " + *iter);
    }
    if (filenames.empty()) onError("usage: pbrt_b200 [--seed n] [--spp n] [--outfile f] [--quiet] scene.pbrt");
    InitPBRT(options);
    Api api;
    const char *lib = getenv("BPT_LIB");
    std::string path = lib ? lib : std::string(BPT_REPO_ROOT) + "/pbrt-v4_b200/libb200pt.so";
    api.Load(path);
    {
        BasicScene scene;
        B200SceneBuilder builder(&scene);
        ParseFiles(&builder, filenames);
        RenderB200(scene, builder, api);
    }
    // (no CleanupPBRT(): the reference's asynchronous light / medium / texture creation jobs were never consumed; exit ends them)
    fflush(stdout);
    _exit(0);
}
