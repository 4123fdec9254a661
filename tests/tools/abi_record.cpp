// abi_record.cpp -- TEST INFRASTRUCTURE: a recording implementation of the scene half of the C ABI (include/b200pt.h).
//
// Built together with the product's .pbrt front end (csrc/pbrt_frontend.cpp, compiled for the host) into libabi_record.so.  Every
// bpt_scene_* call is stored; bpt_scene_commit writes a canonical text dump to $BPT_RECORD_OUT: cameras / films / samplers with
// the engine's defaults filled in, every transform RESOLVED TO RENDERING SPACE with the engine's own code (csrc/scene_xforms.h),
// materials / area lights / media inlined by content where they are referenced, geometry and images by FNV-1a hash, object
// definitions keyed by the hash of their meshes.  The dump is therefore independent of index numbering and of the transform space
// the caller used, and two routes into the engine can be compared byte for byte without a GPU:
//   * bpt_load_pbrt_ex               -- this repo's front end (world-space CTMs),
//   * oracle/_ref/bin/pbrt_b200      -- the reference's own parser + BasicScene walked by oracle/b200_binding.cpp (render space).
// Nothing here renders: bpt_render* return BPT_ERR_UNSUPPORTED.
#include <algorithm>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/b200pt.h"
#include "../../pbrt-v4_b200/csrc/scene_xforms.h"

namespace bpt { uint32_t h_sobol32[104]; }
using namespace bpt;

namespace {
struct Spec { bpt_spectrum s; std::vector<float> samples; };
struct Mesh { bpt_mesh m; std::vector<float> P, N, UV; std::vector<int> idx; bool isQuadric = false; bpt_quadric q; bool isBilinear = false; };
struct Light { bpt_light l; std::vector<float> img; Spec L; };
struct Medium { bpt_medium m; std::vector<float> density, leScale; Spec sa, ss, Le; uint64_t rgbHash[3]; };
struct Material { bpt_material m; Spec sp[8]; };
struct Texture { bpt_texture t; uint64_t hash; Spec value; std::string mapStr; };
struct AreaLight { bpt_area_light a; Spec L; };
}  // namespace

struct bpt_engine {
    std::string err;
    bpt_camera cam{}; bpt_film film{}; bpt_filter filter{}; bpt_sampler sampler{}; bpt_integrator integ{};
    bool haveCam = false, haveFilm = false, haveFilter = false, haveSampler = false, haveInteg = false;
    std::vector<Material> materials;
    std::vector<Texture> textures; std::vector<AreaLight> areaLights; std::vector<Mesh> meshes;
    std::vector<bpt_instance> instances; std::vector<Light> lights; std::vector<Medium> media;
    int cameraMedium = -1, space = BPT_SPACE_WORLD;
};

static Spec Own(const bpt_spectrum &s) {
    Spec o; o.s = s;
    if (s.kind == BPT_SPEC_SAMPLED && s.samples) o.samples.assign(s.samples, s.samples + 2 * (size_t)s.n_samples);
    o.s.samples = nullptr;
    return o;
}
static uint64_t Fnv(const void *p, size_t n, uint64_t h = 1469598103934665603ull) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
static std::string F(float v) { char b[32]; uint32_t u; std::memcpy(&u, &v, 4); snprintf(b, sizeof(b), "%08x", u); return b; }
static std::string M(const M4 &m) { std::string s; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { s += F(m.m[i][j]); s += (j == 3 ? (i == 3 ? "" : "|") : ","); } return s; }
static std::string X(const Xform &t) { return "m=" + M(t.m) + " inv=" + M(t.mInv); }
static std::string S(const Spec &sp) {
    const bpt_spectrum &s = sp.s;
    char b[160];
    switch (s.kind) {
    case BPT_SPEC_NONE: return "none";
    case BPT_SPEC_RGB: return "rgb(" + F(s.rgb[0]) + "," + F(s.rgb[1]) + "," + F(s.rgb[2]) + ")";
    case BPT_SPEC_CONSTANT: return "const(" + F(s.value) + ")";
    case BPT_SPEC_NAMED: return std::string("named(") + s.name + ")";
    case BPT_SPEC_BLACKBODY: return "blackbody(" + F(s.value) + ")";
    case BPT_SPEC_SAMPLED: snprintf(b, sizeof(b), "sampled(%d,%016" PRIx64 ")", s.n_samples, Fnv(sp.samples.data(), sp.samples.size() * 4)); return b;
    }
    return "?";
}
static std::string Tex(const std::vector<Texture> &tx, const char *what, int id) {
    if (id == 0) return "";
    if (id < 0 || id > (int)tx.size()) return std::string(" ") + what + "=BAD";
    const bpt_texture &t = tx[id - 1].t;
    char b[320];
    if (t.kind == BPT_TEXTURE_CONSTANT) return std::string(" ") + what + "=const{" + (t.is_spectrum ? S(tx[id - 1].value) : F(t.value_f)) + "}";
    if (t.kind == BPT_TEXTURE_CHECKERBOARD) {
        if (t.tex1 >= id || t.tex2 >= id) return std::string(" ") + what + "=checkerboard{BAD}";
        return std::string(" ") + what + "=checkerboard{" + (t.is_spectrum ? "spectrum" : "float") + " uv=" + F(t.su) + "," + F(t.sv) + "," + F(t.du) + "," + F(t.dv) + tx[id - 1].mapStr + Tex(tx, "tex1", t.tex1) +
               Tex(tx, "tex2", t.tex2) + "}";
    }
    if (t.kind == BPT_TEXTURE_SCALE || t.kind == BPT_TEXTURE_MIX) {
        const bool bad = t.tex1 >= id || t.tex2 >= id || t.tex3 >= id;   // children come first
        std::string r = std::string(" ") + what + (t.kind == BPT_TEXTURE_SCALE ? "=scale{" : "=mix{") + (t.is_spectrum ? "spectrum" : "float");
        if (bad) return r + " BAD}";
        r += Tex(tx, t.kind == BPT_TEXTURE_SCALE ? "tex" : "tex1", t.tex1) + Tex(tx, t.kind == BPT_TEXTURE_SCALE ? "scale" : "tex2", t.tex2);
        if (t.kind == BPT_TEXTURE_MIX) r += Tex(tx, "amount", t.tex3);
        return r + "}";
    }
    if (t.kind == BPT_TEXTURE_NORMALMAP) {   // image only; an RGBA file counts as its RGB channels
        snprintf(b, sizeof(b), " %s=image{%dx%d fmt=%d px=%016" PRIx64 "}", what, t.xres, t.yres, t.format, tx[id - 1].hash);
        return b;
    }
    snprintf(b, sizeof(b), " %s=tex{%s %dx%dx%d fmt=%d enc=%d px=%016" PRIx64 " filter=%d aniso=%s wrap=%d scale=%s invert=%d uv=%s,%s,%s,%s}", what, t.is_spectrum ? "spectrum" : "float", t.xres, t.yres, t.n_channels, t.format,
             t.format == BPT_IMAGE_U8 ? t.encoding : 0, tx[id - 1].hash, t.filter, F(t.max_anisotropy).c_str(), t.wrap, F(t.scale).c_str(), t.invert != 0, F(t.su).c_str(), F(t.sv).c_str(),
             F(t.du).c_str(), F(t.dv).c_str());
    return std::string(b) + tx[id - 1].mapStr;
}
static std::string MatBody(const Material &mt);
static std::string Mat(const Material &mt, const std::vector<Texture> &tx) {
    const bpt_material &m = mt.m;
    std::string s = MatBody(mt);
    if (m.type == BPT_MAT_INTERFACE) return s;
    s.pop_back();   // the closing brace
    return s + Tex(tx, "refl_tex", m.reflectance_tex) + Tex(tx, "trans_tex", m.transmittance_tex) + Tex(tx, "urough_tex", m.uroughness_tex) + Tex(tx, "vrough_tex", m.vroughness_tex) +
           Tex(tx, "displacement", m.displacement_tex) + Tex(tx, "normalmap", m.normal_map_tex) + "}";
}
static std::string MatBody(const Material &mt) {
    const bpt_material &m = mt.m;
    std::string s = "mat{type=" + std::to_string(m.type);
    if (m.type == BPT_MAT_INTERFACE) return s + "}";
    if (m.type == BPT_MAT_THIN_DIELECTRIC) return s + " eta=" + S(mt.sp[1]) + "}";
    if (m.type == BPT_MAT_DIFFUSE_TRANSMISSION) return s + " refl=" + S(mt.sp[0]) + " trans=" + S(mt.sp[7]) + " scale=" + F(m.scale > 0 ? m.scale : 1.f) + "}";
    if (m.type == BPT_MAT_COATED_DIFFUSE)
        return s + " refl=" + S(mt.sp[0]) + " eta=" + S(mt.sp[1]) + " rough=" + F(m.roughness) + "," + F(m.uroughness) + "," + F(m.vroughness) + " remap=" +
               std::to_string(m.remap_roughness != 0) + " thick=" + F(m.thickness) + " g=" + F(m.g) + " albedo=" + S(mt.sp[6]) + " depth=" + std::to_string(m.max_depth) +
               " ns=" + std::to_string(m.n_samples) + "}";
    s += " refl=" + S(mt.sp[0]) + " eta=" + S(mt.sp[1]) + " k=" + S(mt.sp[2]);
    if (m.type != BPT_MAT_DIFFUSE)
        s += " rough=" + F(m.roughness) + "," + F(m.uroughness) + "," + F(m.vroughness) + " remap=" + std::to_string(m.remap_roughness != 0);
    if (m.type == BPT_MAT_COATED_CONDUCTOR)
        s += " iface=" + F(m.iface_roughness) + "," + F(m.iface_uroughness) + "," + F(m.iface_vroughness) + " thick=" + F(m.thickness) + " g=" + F(m.g) +
             " ieta=" + S(mt.sp[3]) + " ceta=" + S(mt.sp[4]) + " ck=" + S(mt.sp[5]) + " albedo=" + S(mt.sp[6]) + " depth=" + std::to_string(m.max_depth) +
             " ns=" + std::to_string(m.n_samples);
    return s + "}";
}

extern "C" {
int bpt_create(int, const char *, bpt_engine **out) { *out = new bpt_engine; return BPT_OK; }
void bpt_destroy(bpt_engine *e) { delete e; }
const char *bpt_last_error(const bpt_engine *e) { return e ? e->err.c_str() : ""; }
int bpt_scene_begin(bpt_engine *e) { std::string keep; *e = bpt_engine(); return BPT_OK; }
int bpt_scene_set_transform_space(bpt_engine *e, int s) { e->space = s; return BPT_OK; }
int bpt_scene_set_camera(bpt_engine *e, const bpt_camera *c) { e->cam = *c; e->haveCam = true; return BPT_OK; }
int bpt_scene_set_film(bpt_engine *e, const bpt_film *f) { e->film = *f; e->haveFilm = true; return BPT_OK; }
int bpt_scene_set_filter(bpt_engine *e, const bpt_filter *f) { e->filter = *f; e->haveFilter = true; return BPT_OK; }
int bpt_scene_set_sampler(bpt_engine *e, const bpt_sampler *s) { e->sampler = *s; e->haveSampler = true; return BPT_OK; }
int bpt_scene_set_integrator(bpt_engine *e, const bpt_integrator *i) { e->integ = *i; e->haveInteg = true; return BPT_OK; }
int bpt_scene_set_materials(bpt_engine *e, const bpt_material *m, int n) {
    e->materials.resize(n);
    for (int i = 0; i < n; ++i) {
        Material &o = e->materials[i]; o.m = m[i];
        const bpt_spectrum *sp[8] = {&m[i].reflectance, &m[i].eta, &m[i].k, &m[i].iface_eta, &m[i].conductor_eta, &m[i].conductor_k, &m[i].albedo, &m[i].transmittance};
        for (int k = 0; k < 8; ++k) o.sp[k] = Own(*sp[k]);
    }
    return BPT_OK;
}
int bpt_scene_set_textures(bpt_engine *e, const bpt_texture *t, int n) {
    e->textures.resize(n);
    for (int i = 0; i < n; ++i) {
        e->textures[i].t = t[i];
        if (t[i].kind == BPT_TEXTURE_CONSTANT || t[i].kind == BPT_TEXTURE_SCALE || t[i].kind == BPT_TEXTURE_MIX || t[i].kind == BPT_TEXTURE_CHECKERBOARD) { e->textures[i].value = Own(t[i].value_s); e->textures[i].hash = 0; continue; }
        const size_t es = t[i].format == BPT_IMAGE_U8 ? 1 : 4, bytes = (size_t)t[i].xres * t[i].yres * t[i].n_channels * es;
        if (t[i].kind == BPT_TEXTURE_NORMALMAP && t[i].n_channels == 4) {
            uint64_t h = 1469598103934665603ull;
            for (size_t px = 0; px < (size_t)t[i].xres * t[i].yres; ++px) h = Fnv((const char *)t[i].pixels + px * 4 * es, 3 * es, h);
            e->textures[i].hash = h;
        } else
            e->textures[i].hash = Fnv(t[i].pixels, bytes);
        e->textures[i].t.pixels = nullptr;
    }
    return BPT_OK;
}
int bpt_scene_set_area_lights(bpt_engine *e, const bpt_area_light *a, int n) {
    e->areaLights.resize(n);
    for (int i = 0; i < n; ++i) { e->areaLights[i].a = a[i]; e->areaLights[i].L = Own(a[i].L); }
    return BPT_OK;
}
int bpt_scene_add_mesh(bpt_engine *e, const bpt_mesh *m) {
    e->meshes.emplace_back(); Mesh &o = e->meshes.back(); o.m = *m;
    o.P.assign(m->P, m->P + 3 * (size_t)m->n_vertices);
    if (m->N) o.N.assign(m->N, m->N + 3 * (size_t)m->n_vertices);
    if (m->UV) o.UV.assign(m->UV, m->UV + 2 * (size_t)m->n_vertices);
    o.idx.assign(m->indices, m->indices + 3 * (size_t)m->n_triangles);
    return BPT_OK;
}
int bpt_scene_add_bilinear_mesh(bpt_engine *e, const bpt_mesh *m) {
    e->meshes.emplace_back(); Mesh &o = e->meshes.back(); o.m = *m; o.isBilinear = true;
    o.P.assign(m->P, m->P + 3 * (size_t)m->n_vertices);
    if (m->N) o.N.assign(m->N, m->N + 3 * (size_t)m->n_vertices);
    if (m->UV) o.UV.assign(m->UV, m->UV + 2 * (size_t)m->n_vertices);
    o.idx.assign(m->indices, m->indices + 4 * (size_t)m->n_triangles);
    return BPT_OK;
}
int bpt_scene_add_quadric(bpt_engine *e, const bpt_quadric *q) {
    e->meshes.emplace_back(); Mesh &o = e->meshes.back(); std::memset(&o.m, 0, sizeof(o.m)); o.isQuadric = true; o.q = *q;
    o.m.world_from_object = q->world_from_object; o.m.object_from_world = q->object_from_world; o.m.reverse_orientation = q->reverse_orientation;
    o.m.material = q->material; o.m.area_light = q->area_light; o.m.medium_inside = q->medium_inside; o.m.medium_outside = q->medium_outside;
    o.m.object_id = q->object_id;
    return BPT_OK;
}
int bpt_scene_add_instance(bpt_engine *e, const bpt_instance *i) { e->instances.push_back(*i); return BPT_OK; }
int bpt_scene_add_light(bpt_engine *e, const bpt_light *l) {
    e->lights.emplace_back(); Light &o = e->lights.back(); o.l = *l; o.L = Own(l->L);
    if (l->image_rgb) o.img.assign(l->image_rgb, l->image_rgb + 3 * (size_t)l->image_res * l->image_res);
    return BPT_OK;
}
int bpt_scene_add_medium(bpt_engine *e, const bpt_medium *m) {
    e->media.emplace_back(); Medium &o = e->media.back(); o.m = *m; o.sa = Own(m->sigma_a); o.ss = Own(m->sigma_s); o.Le = Own(m->Le);
    if (m->density) o.density.assign(m->density, m->density + (size_t)m->nx * m->ny * m->nz);
    if (m->Le_scale_grid) o.leScale.assign(m->Le_scale_grid, m->Le_scale_grid + (size_t)m->nx * m->ny * m->nz);
    const float *g3[3] = {m->sigma_a_rgb, m->sigma_s_rgb, m->Le_rgb};
    for (int k = 0; k < 3; ++k) o.rgbHash[k] = (m->type == BPT_MEDIUM_RGBGRID && g3[k]) ? Fnv(g3[k], 12 * (size_t)m->nx * m->ny * m->nz) : 0;
    return BPT_OK;
}
int bpt_scene_set_camera_medium(bpt_engine *e, int m) { e->cameraMedium = m; return BPT_OK; }

int bpt_scene_commit(bpt_engine *e) {
    const char *path = std::getenv("BPT_RECORD_OUT");
    FILE *f = path ? fopen(path, "w") : stdout;
    if (!f) { e->err = "cannot open BPT_RECORD_OUT"; return BPT_ERR_IO; }
    CameraSpaces cs;
    if (!e->haveCam || !ResolveCameraSpaces(e->cam, &cs)) { e->err = "no / singular camera"; return BPT_ERR_INVALID; }
    for (Texture &t : e->textures) {   // texture mappings that look at the shading point: resolved to rendering space like every other transform
        t.mapStr.clear();
        if (t.t.mapping == BPT_TEXMAP_UV || (t.t.kind != BPT_TEXTURE_IMAGEMAP && t.t.kind != BPT_TEXTURE_CHECKERBOARD)) continue;
        Xform rft = ResolveRenderFromObject(e->space, cs, t.t.world_from_texture, t.t.texture_from_world);
        t.mapStr = " mapping=" + std::to_string(t.t.mapping) + "{" + X(rft) + " v1=" + F(t.t.v1[0]) + "," + F(t.t.v1[1]) + "," + F(t.t.v1[2]) + " v2=" + F(t.t.v2[0]) + "," + F(t.t.v2[1]) + "," +
                   F(t.t.v2[2]) + "}";
    }
    const bpt_camera &c = e->cam;
    fprintf(f, "camera type=%d renderFromCamera{%s} worldFromRender{%s} fov=%s lens=%s focal=%s aspect=%s sw=%d:%s,%s,%s,%s shutter=%s,%s medium=%s\n", c.type,
            X(cs.renderFromCamera).c_str(), X(cs.worldFromRender).c_str(), F(c.fov).c_str(), F(c.lens_radius).c_str(), F(c.focal_distance).c_str(),
            F(c.frame_aspect_ratio > 0 ? c.frame_aspect_ratio : -1.f).c_str(), c.has_screen_window, F(c.screen_window[0]).c_str(), F(c.screen_window[1]).c_str(),
            F(c.screen_window[2]).c_str(), F(c.screen_window[3]).c_str(), F(c.shutter_open).c_str(), F(c.shutter_close).c_str(),
            e->cameraMedium >= 0 ? "yes" : "none");
    bpt_film fl = e->film;
    if (fl.pixel_bounds[0] == 0 && fl.pixel_bounds[1] == 0 && fl.pixel_bounds[2] == 0 && fl.pixel_bounds[3] == 0) { fl.pixel_bounds[1] = fl.xres; fl.pixel_bounds[3] = fl.yres; }
    fprintf(f, "film type=%d res=%dx%d bounds=%d,%d,%d,%d iso=%s wb=%s maxc=%s diag=%s csys=%d\n", fl.type, fl.xres, fl.yres, fl.pixel_bounds[0], fl.pixel_bounds[1],
            fl.pixel_bounds[2], fl.pixel_bounds[3], F(fl.iso > 0 ? fl.iso : 100.f).c_str(), F(fl.white_balance).c_str(),
            F(fl.max_component_value > 0 ? fl.max_component_value : -1.f).c_str(), F(fl.diagonal > 0 ? fl.diagonal : 35.f).c_str(), fl.coordinate_system);
    bpt_filter ft = e->haveFilter ? e->filter : bpt_filter{BPT_FILTER_GAUSSIAN, 1.5f, 1.5f, 0.5f, 0, 0, 0};
    fprintf(f, "filter type=%d r=%s,%s sigma=%s b=%s c=%s tau=%s\n", ft.type, F(ft.xradius).c_str(), F(ft.yradius).c_str(),
            F(ft.type == BPT_FILTER_GAUSSIAN ? ft.sigma : 0.f).c_str(), F(ft.type == BPT_FILTER_MITCHELL ? ft.b : 0.f).c_str(),
            F(ft.type == BPT_FILTER_MITCHELL ? ft.c : 0.f).c_str(), F(ft.type == BPT_FILTER_SINC ? ft.tau : 0.f).c_str());
    bpt_sampler sm = e->haveSampler ? e->sampler : bpt_sampler{BPT_SAMPLER_ZSOBOL, 16, 0, BPT_RANDOMIZE_FAST_OWEN, 0, 0, 0};
    const bool strat = sm.type == BPT_SAMPLER_STRATIFIED, sob = sm.type == BPT_SAMPLER_ZSOBOL || sm.type == BPT_SAMPLER_PADDED_SOBOL || sm.type == BPT_SAMPLER_SOBOL || sm.type == BPT_SAMPLER_HALTON;
    fprintf(f, "sampler type=%d spp=%d seed=%d randomize=%d strata=%dx%d jitter=%d\n", sm.type, strat ? sm.x_samples * sm.y_samples : sm.spp, sm.seed,
            sob ? sm.randomize : 0, strat ? sm.x_samples : 0, strat ? sm.y_samples : 0, strat ? (sm.jitter != 0) : 0);
    bpt_integrator ig = e->haveInteg ? e->integ : bpt_integrator{5, 0, BPT_LIGHTSAMPLER_BVH};
    fprintf(f, "integrator maxdepth=%d regularize=%d lightsampler=%d\n", ig.max_depth, ig.regularize != 0, ig.light_sampler);

    auto mediumStr = [&](int id) -> std::string {
        if (id < 0) return "none";
        if (id >= (int)e->media.size()) return "BAD";
        const Medium &md = e->media[id];
        Xform rfm = ResolveRenderFromObject(e->space, cs, md.m.world_from_medium, md.m.medium_from_world);
        char b[96]; snprintf(b, sizeof(b), " grid=%dx%dx%d:%016" PRIx64, md.m.nx, md.m.ny, md.m.nz, Fnv(md.density.data(), md.density.size() * 4));
        std::string s = "medium{type=" + std::to_string(md.m.type) + " sa=" + S(md.sa) + " ss=" + S(md.ss) + " Le=" + S(md.Le) + " scale=" + F(md.m.scale) + " g=" + F(md.m.g) +
                        " Lescale=" + F(md.Le.s.kind == BPT_SPEC_NONE ? 1.f : md.m.Le_scale) + ":" + std::to_string(md.leScale.empty() ? 0ull : (unsigned long long)Fnv(md.leScale.data(), md.leScale.size() * 4));
        if (md.m.type == BPT_MEDIUM_RGBGRID) {
            char rb[200]; snprintf(rb, sizeof(rb), " rgbgrid=%dx%dx%d:%016" PRIx64 ",%016" PRIx64 ",%016" PRIx64, md.m.nx, md.m.ny, md.m.nz, md.rgbHash[0], md.rgbHash[1], md.rgbHash[2]);
            std::string r = "medium{type=2 scale=" + F(md.m.scale) + " g=" + F(md.m.g) + " Lescale=" + F(md.m.Le_scale) + rb + " bounds=";
            for (int k = 0; k < 6; ++k) r += F(md.m.bounds[k]) + (k < 5 ? "," : "");
            return r + " " + X(rfm) + "}";
        }
        if (md.m.type == BPT_MEDIUM_UNIFORMGRID) {
            s += b; s += " bounds=";
            for (int k = 0; k < 6; ++k) s += F(md.m.bounds[k]) + (k < 5 ? "," : "");
            s += " " + X(rfm);
        }
        return s + "}";
    };
    auto meshStr = [&](const Mesh &ms) {
        const bpt_mesh &m = ms.m;
        Xform rfo = ResolveRenderFromObject(e->space, cs, m.world_from_object, m.object_from_world);
        std::string s = std::string(ms.isBilinear ? "bilinearmesh" : "mesh") + " nv=" + std::to_string(m.n_vertices) + " nt=" + std::to_string(m.n_triangles);
        char b[200];
        if (ms.isQuadric)
            s = "quadric type=" + std::to_string(ms.q.type) + " radius=" + F(ms.q.radius) + " zmin=" + F(ms.q.zmin) + " zmax=" + F(ms.q.zmax) + " phimax=" + F(ms.q.phimax) +
                " height=" + F(ms.q.height) + " innerradius=" + F(ms.q.innerradius);
        else {
        snprintf(b, sizeof(b), " P=%016" PRIx64 " N=%016" PRIx64 " UV=%016" PRIx64 " idx=%016" PRIx64, Fnv(ms.P.data(), ms.P.size() * 4),
                 ms.N.empty() ? 0 : Fnv(ms.N.data(), ms.N.size() * 4), ms.UV.empty() ? 0 : Fnv(ms.UV.data(), ms.UV.size() * 4), Fnv(ms.idx.data(), ms.idx.size() * 4));
        s += b;
        }
        s += " renderFromObject{" + X(rfo) + "} reverse=" + std::to_string(m.reverse_orientation != 0) + " alpha=" + (m.has_alpha ? F(m.alpha) : std::string("none"));
        Material iface; std::memset(&iface.m, 0, sizeof(iface.m)); iface.m.type = BPT_MAT_INTERFACE;
        const Material &mt = m.material < 0 ? iface : e->materials[m.material];
        s += " " + Mat(mt, e->textures);
        if (m.area_light >= 0 && mt.m.type != BPT_MAT_INTERFACE) {
            const AreaLight &al = e->areaLights[m.area_light];
            s += " arealight{L=" + S(al.L) + " scale=" + F(al.a.scale) + " two=" + std::to_string(al.a.two_sided != 0) + " power=" + F(al.a.power > 0 ? al.a.power : -1.f) + "}";
        } else s += " arealight{none}";
        s += " inside=" + mediumStr(m.medium_inside) + " outside=" + mediumStr(m.medium_outside);
        return s;
    };
    fprintf(f, "cameraMedium %s\n", mediumStr(e->cameraMedium).c_str());
    // top-level meshes in scene order; object definitions keyed by the hash of their (ordered) mesh records
    std::map<int, std::vector<std::string>> objMeshes;
    for (const Mesh &ms : e->meshes) {
        if (ms.m.object_id < 0) fprintf(f, "%s\n", meshStr(ms).c_str());
        else objMeshes[ms.m.object_id].push_back(meshStr(ms));
    }
    std::map<int, uint64_t> objHash;
    std::vector<std::pair<uint64_t, int>> order;
    for (auto &it : objMeshes) {
        uint64_t h = 1469598103934665603ull;
        for (const std::string &s : it.second) h = Fnv(s.data(), s.size(), h);
        objHash[it.first] = h; order.push_back({h, it.first});
    }
    std::sort(order.begin(), order.end());
    for (auto &o : order) {
        fprintf(f, "object %016" PRIx64 "\n", o.first);
        for (const std::string &s : objMeshes[o.second]) fprintf(f, "  %s\n", s.c_str());
    }
    for (const bpt_instance &in : e->instances) {
        Xform rfi = ResolveRenderFromInstance(e->space, cs, in.world_from_instance, in.instance_from_world);
        fprintf(f, "instance object=%016" PRIx64 " renderFromInstance{%s}\n", objHash.count(in.object_id) ? objHash[in.object_id] : 0, X(rfi).c_str());
    }
    for (const Light &l : e->lights) {
        Xform rfl = ResolveRenderFromObject(e->space, cs, l.l.world_from_light, l.l.light_from_world);
        fprintf(f, "light type=%d renderFromLight{%s} L=%s scale=%s illuminance=%s image=%d:%016" PRIx64, l.l.type, X(rfl).c_str(), S(l.L).c_str(), F(l.l.scale).c_str(),
                F(l.l.illuminance > 0 ? l.l.illuminance : -1.f).c_str(), l.l.image_res, l.img.empty() ? 0 : Fnv(l.img.data(), l.img.size() * 4));
        if (l.l.type == BPT_LIGHT_POINT || l.l.type == BPT_LIGHT_SPOT || l.l.type == BPT_LIGHT_DISTANT)
            fprintf(f, " from=%s,%s,%s to=%s,%s,%s power=%s cone=%s,%s", F(l.l.from[0]).c_str(), F(l.l.from[1]).c_str(), F(l.l.from[2]).c_str(), F(l.l.to[0]).c_str(),
                    F(l.l.to[1]).c_str(), F(l.l.to[2]).c_str(), F(l.l.power > 0 ? l.l.power : -1.f).c_str(), F(l.l.cone_angle).c_str(), F(l.l.cone_delta).c_str());
        fprintf(f, "\n");
    }
    if (path) fclose(f);
    return BPT_OK;
}

static int Unsupported(bpt_engine *e) { if (e) e->err = "libabi_record.so records scene calls; it does not render"; return BPT_ERR_UNSUPPORTED; }
int bpt_render(bpt_engine *e, int, int, const int *, int) { return Unsupported(e); }
int bpt_film_dims(const bpt_engine *e, int *x, int *y) { if (x) *x = e->film.xres; if (y) *y = e->film.yres; return BPT_OK; }
int bpt_film_resolve(bpt_engine *e, float *) { return Unsupported(e); }
int bpt_film_resolve_gbuffer(bpt_engine *e, float *) { return Unsupported(e); }
int bpt_sampler_spp(const bpt_engine *e) { return e->haveSampler ? e->sampler.spp : 16; }
}
