// pbrt_frontend.cpp -- ".pbrt scene in" front end above the C ABI: a from-scratch tokenizer / directive interpreter
// for the subset of the pbrt-v4 scene format that the hot path consumes.  It plays the role of
// parser.cpp + BasicSceneBuilder (scene.cpp:105-737) for these directives and feeds bpt_scene_* exactly what
// BasicScene would hold (CTMs with tracked inverses, parameter lists resolved to POD records).
//
// Supported: LookAt Translate Scale Rotate Identity Transform ConcatTransform CoordinateSystem CoordSysTransform
//            Camera("perspective") Film("rgb"|"gbuffer") Sampler("zsobol"|"independent") PixelFilter("gaussian"|"box")
//            Integrator("volpath"|"path") Option ColorSpace("srgb") WorldBegin AttributeBegin/End TransformBegin/End
//            ReverseOrientation Material MakeNamedMaterial NamedMaterial AreaLightSource("diffuse")
//            LightSource("infinite") Shape("trianglemesh"|"plymesh") ObjectBegin/End ObjectInstance
//            MakeNamedMedium MediumInterface Include
// Anything else is an error (never silently ignored), matching the reference's ErrorExit behaviour.
#include <thread>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>
#include "../../include/b200pt.h"
#include "xform.h"

using namespace bpt;

#ifdef BPT_FRONTEND_IN_LIBRARY
// (engine.cu) moves the mesh arrays into the engine; the stand-alone build of this file for the ABI recorder uses the public call
int bpt_internal_add_mesh_move(bpt_engine *e, const bpt_mesh *m, std::vector<float> &&P, std::vector<float> &&N, std::vector<float> &&UV, std::vector<int> &&idx);
#endif
namespace {

struct Param {
    std::string type, name;
    std::vector<double> nums;
    std::vector<std::string> strs;
    std::vector<bool> bools;
    bool used = false;
};
struct ParamList {
    std::vector<Param> p;
    Param *find(const std::string &name) { for (auto &q : p) if (q.name == name) { q.used = true; return &q; } return nullptr; }
    float f(const std::string &n, float d) { Param *q = find(n); return (q && !q->nums.empty()) ? (float)q->nums[0] : d; }
    int i(const std::string &n, int d) { Param *q = find(n); return (q && !q->nums.empty()) ? (int)q->nums[0] : d; }
    bool b(const std::string &n, bool d) { Param *q = find(n); return (q && !q->bools.empty()) ? (bool)q->bools[0] : d; }
    std::string s(const std::string &n, const std::string &d) { Param *q = find(n); return (q && !q->strs.empty()) ? q->strs[0] : d; }
    std::string unused() const { for (auto &q : p) if (!q.used) return q.type + " " + q.name; return ""; }
};

struct Tokenizer {
    std::string buf; size_t pos = 0; std::string file; int line = 1;
    bool load(const std::string &path) {
        std::ifstream in(path, std::ios::binary);
        if (!in) return false;
        std::stringstream ss; ss << in.rdbuf(); buf = ss.str(); file = path; return true;
    }
    // returns false at EOF.  quoted: token was a "string"
    bool next(std::string *tok, bool *quoted) {
        while (pos < buf.size()) {
            char c = buf[pos];
            if (c == '\n') { ++line; ++pos; }
            else if (c == ' ' || c == '\t' || c == '\r') ++pos;
            else if (c == '#') { while (pos < buf.size() && buf[pos] != '\n') ++pos; }
            else break;
        }
        if (pos >= buf.size()) return false;
        char c = buf[pos];
        *quoted = false;
        if (c == '"') {
            size_t e = buf.find('"', pos + 1);
            if (e == std::string::npos) e = buf.size();
            *tok = buf.substr(pos + 1, e - pos - 1); pos = e + 1; *quoted = true; return true;
        }
        if (c == '[' || c == ']') { *tok = std::string(1, c); ++pos; return true; }
        size_t s = pos;
        while (pos < buf.size() && !strchr(" \t\r\n\"[]#", buf[pos])) ++pos;
        *tok = buf.substr(s, pos - s);
        return true;
    }
};

struct GState {
    Xform ctm;
    bool reverseOrientation = false;
    int material = 0;
    int areaLight = -1;
    std::string mediumInside, mediumOutside;
};

struct MeshRec {
    std::vector<float> P, N, UV; std::vector<int> idx;
    std::vector<int> quadIdx;   // bilinear patches (Shape "bilinearmesh", or the quad faces of a PLY file): p00, p10, p01, p11 per patch
    bpt_mesh m;
    bool isQuadric = false;   // Shape "sphere" / "disk" / "cylinder": q carries the parameters, m the members common to all shapes
    bpt_quadric q;
};

struct Frontend {
    bpt_engine *e;
    std::string err;
    std::string baseDir;
    GState gs;
    std::vector<GState> stack;
    std::vector<Xform> xstack;
    std::map<std::string, Xform> namedCS;
    bool inWorld = false;
    bpt_camera cam{}; bool haveCam = false;
    bpt_film film{}; bpt_filter filter{}; bpt_sampler sampler{}; bpt_integrator integ{};
    bool haveFilter = false, haveSampler = false, haveInteg = false;
    std::vector<bpt_material> materials;
    std::vector<std::unique_ptr<std::vector<float>>> sampleStore;
    std::map<std::string, int> namedMaterials;
    std::vector<bpt_area_light> areaLights;
    std::vector<std::unique_ptr<MeshRec>> meshes;
    std::vector<bpt_instance> instances;
    std::vector<bpt_light> lights;
    std::vector<std::unique_ptr<std::vector<float>>> imageStore;
    std::map<std::string, int> objectIds;
    int curObject = -1;
    std::vector<bpt_medium> media;
    std::vector<std::unique_ptr<std::vector<float>>> densityStore;
    std::map<std::string, int> namedMedia;
    std::string cameraMediumName;
    int seedOption = 0;
    int sppOverride = 0;
    // Texture "name" "float" | "spectrum" "imagemap": one bpt_texture per directive; material parameters of type "texture" refer to them
    std::vector<bpt_texture> textures;
    std::vector<std::unique_ptr<std::vector<unsigned char>>> texelStore;
    std::map<std::string, int> floatTextures, spectrumTextures;   // name -> index + 1
    std::map<std::string, int> normalMaps;                        // resolved file name -> index + 1

    bool fail(const Tokenizer &t, const std::string &m) { err = t.file + ":" + std::to_string(t.line) + ": " + m; return false; }

    static bpt_mat4 ToB(const M4 &a) { bpt_mat4 r; std::memcpy(r.m, a.m, sizeof(r.m)); return r; }

    bool getSpectrum(ParamList &pl, const std::string &name, bpt_spectrum *out) {
        std::memset(out, 0, sizeof(*out));
        Param *q = pl.find(name);
        if (!q) { out->kind = BPT_SPEC_NONE; return true; }
        if (q->type == "rgb" && q->nums.size() == 3) { out->kind = BPT_SPEC_RGB; for (int k = 0; k < 3; ++k) out->rgb[k] = (float)q->nums[k]; return true; }
        if (q->type == "float" && q->nums.size() == 1) { out->kind = BPT_SPEC_CONSTANT; out->value = (float)q->nums[0]; return true; }
        if (q->type == "spectrum") {
            if (!q->strs.empty()) { out->kind = BPT_SPEC_NAMED; std::strncpy(out->name, q->strs[0].c_str(), sizeof(out->name) - 1); return true; }
            if (q->nums.size() >= 2 && q->nums.size() % 2 == 0) {
                sampleStore.emplace_back(new std::vector<float>(q->nums.begin(), q->nums.end()));
                out->kind = BPT_SPEC_SAMPLED; out->samples = sampleStore.back()->data(); out->n_samples = (int)q->nums.size() / 2; return true;
            }
        }
        if (q->type == "blackbody" && q->nums.size() == 1) { out->kind = BPT_SPEC_BLACKBODY; out->value = (float)q->nums[0]; return true; }
        err = "parameter \"" + name + "\": unsupported spectrum specification (type " + q->type + ")";
        return false;
    }

    // "texture <name>" in place of a spectrum / float parameter: the id (index + 1) of the named texture, the constant left unset
    bool getSpectrumOrTexture(ParamList &pl, const std::string &name, bpt_spectrum *out, int *tex) {
        *tex = 0;
        Param *q = pl.find(name);
        if (q && q->type == "texture") {
            std::memset(out, 0, sizeof(*out));
            out->kind = BPT_SPEC_NONE;
            if (q->strs.empty() || !spectrumTextures.count(q->strs[0])) { err = "parameter \"" + name + "\": unknown spectrum texture"; return false; }
            *tex = spectrumTextures[q->strs[0]];
            return true;
        }
        return getSpectrum(pl, name, out);
    }
    bool getFloatOrTexture(ParamList &pl, const std::string &name, float dflt, float *v, int *tex) {
        *tex = 0; *v = dflt;
        Param *q = pl.find(name);
        if (!q) return true;
        if (q->type == "texture") {
            if (q->strs.empty() || !floatTextures.count(q->strs[0])) { err = "parameter \"" + name + "\": unknown float texture"; return false; }
            *tex = floatTextures[q->strs[0]];
            return true;
        }
        if (!q->nums.empty()) *v = (float)q->nums[0];
        return true;
    }
    // "roughness" / "uroughness" / "vroughness" as ConductorMaterial::Create and friends resolve them (materials.cpp): the u / v
    // parameter wins, otherwise both take "roughness" (constant or texture)
    bool getRoughness(ParamList &pl, const std::string &prefix, float *rough, float *u, float *v, int *utex, int *vtex) {
        int rt = 0, ut = 0, vt = 0;
        float r = 0, uu = 0, vv = 0;
        if (!getFloatOrTexture(pl, prefix + "roughness", 0.f, &r, &rt)) return false;
        bool haveU = pl.find(prefix + "uroughness") != nullptr, haveV = pl.find(prefix + "vroughness") != nullptr;
        if (!getFloatOrTexture(pl, prefix + "uroughness", r, &uu, &ut) || !getFloatOrTexture(pl, prefix + "vroughness", r, &vv, &vt)) return false;
        *rough = r; *u = uu; *v = vv;
        if (utex) { *utex = haveU ? ut : rt; *vtex = haveV ? vt : rt; }
        else if (rt || ut || vt) { err = "parameter \"" + prefix + "roughness\": textures are not supported here"; return false; }
        return true;
    }
    bool getDisplacement(ParamList &pl, bpt_material *m) {
        // "string normalmap": the image is read with the linear encoding (BasicScene::startLoadingNormalMaps); one entry per file
        std::string nm = pl.s("normalmap", "");
        if (!nm.empty()) {
            if (nm[0] != '/') nm = baseDir + "/" + nm;
            if (!normalMaps.count(nm)) {
                bpt_texture t; std::memset(&t, 0, sizeof(t));
                t.kind = BPT_TEXTURE_NORMALMAP; t.encoding = BPT_ENCODING_LINEAR; t.wrap = BPT_TEXWRAP_REPEAT; t.filter = BPT_TEXFILTER_BILINEAR; t.scale = 1; t.su = t.sv = 1;
                const bool png = nm.size() > 4 && (nm.substr(nm.size() - 4) == ".png" || nm.substr(nm.size() - 4) == ".PNG");
                if (png ? !readPNG(nm, &t) : !readPFMTexture(nm, &t)) return false;
                if (t.n_channels < 3) { err = nm + ": normal map image must contain R, G, and B channels"; return false; }
                textures.push_back(t);
                normalMaps[nm] = (int)textures.size();
            }
            m->normal_map_tex = normalMaps[nm];
        }
        Param *q = pl.find("displacement");
        if (!q) return true;
        if (q->type != "texture" || q->strs.empty() || !floatTextures.count(q->strs[0])) { err = "\"displacement\" must name a float texture"; return false; }
        m->displacement_tex = floatTextures[q->strs[0]];
        return true;
    }

    // ---- image files for textures: what Image::Read returns for .png (8-bit) and .pfm (util/image.cpp:1260-1365, ReadPFM)
    static uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
    bool readPNG(const std::string &path, bpt_texture *t) {
        std::ifstream in(path, std::ios::binary);
        if (!in) { err = "cannot open " + path; return false; }
        std::string buf((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
        const unsigned char *d = (const unsigned char *)buf.data();
        static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
        if (buf.size() < 33 || std::memcmp(d, sig, 8) != 0) { err = path + ": not a PNG file"; return false; }
        uint32_t w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
        std::string idat;
        for (size_t pos = 8; pos + 12 <= buf.size();) {
            uint32_t len = be32(d + pos);
            if (pos + 12 + len > buf.size()) { err = path + ": truncated PNG"; return false; }
            const unsigned char *ty = d + pos + 4, *body = d + pos + 8;
            if (!std::memcmp(ty, "IHDR", 4)) { w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
            else if (!std::memcmp(ty, "IDAT", 4)) idat.append((const char *)body, len);
            else if (!std::memcmp(ty, "IEND", 4)) break;
            pos += 12 + len;
        }
        if (depth != 8 || interlace != 0 || (ctype != 0 && ctype != 2 && ctype != 4 && ctype != 6)) {
            err = path + ": only non-interlaced 8-bit grey / grey+alpha / RGB / RGBA PNG files are supported"; return false;
        }
        const int fileCh = ctype == 0 ? 1 : ctype == 4 ? 2 : ctype == 2 ? 3 : 4;
        const size_t stride = (size_t)w * fileCh;
        std::vector<unsigned char> raw((stride + 1) * h);
        uLongf rawLen = (uLongf)raw.size();
        if (uncompress(raw.data(), &rawLen, (const Bytef *)idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) { err = path + ": PNG data does not inflate"; return false; }
        std::vector<unsigned char> img(stride * h);
        for (uint32_t y = 0; y < h; ++y) {
            const unsigned char *src = raw.data() + (stride + 1) * y;
            unsigned char *dst = img.data() + stride * y;
            const unsigned char *up = y ? dst - stride : nullptr;
            const int ft = src[0];
            for (size_t i = 0; i < stride; ++i) {
                int a = i >= (size_t)fileCh ? dst[i - fileCh] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)fileCh) ? up[i - fileCh] : 0, x = src[1 + i], v;
                switch (ft) {
                case 0: v = x; break;
                case 1: v = x + a; break;
                case 2: v = x + b; break;
                case 3: v = x + ((a + b) >> 1); break;
                case 4: { int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c); v = x + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)); break; }
                default: err = path + ": bad PNG filter type"; return false;
                }
                dst[i] = (unsigned char)v;
            }
        }
        // grey+alpha is decoded as grey (the reference asks lodepng for LCT_GREY)
        int nc = fileCh == 2 ? 1 : fileCh;
        texelStore.emplace_back(new std::vector<unsigned char>((size_t)w * h * nc));
        std::vector<unsigned char> &px = *texelStore.back();
        if (fileCh == 2) for (size_t i = 0; i < (size_t)w * h; ++i) px[i] = img[2 * i];
        else px = img;
        t->format = BPT_IMAGE_U8; t->xres = (int)w; t->yres = (int)h; t->n_channels = nc; t->pixels = px.data();
        return true;
    }
    bool readPFMTexture(const std::string &path, bpt_texture *t) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) { err = "cannot open " + path; return false; }
        char hdr[3]; int w, h; float sc;
        if (std::fscanf(f, "%2s %d %d %f", hdr, &w, &h, &sc) != 4 || (std::strcmp(hdr, "PF") != 0 && std::strcmp(hdr, "Pf") != 0)) { std::fclose(f); err = path + ": PFM header expected"; return false; }
        std::fgetc(f);
        if (sc > 0) { std::fclose(f); err = path + ": big-endian PFM not supported"; return false; }
        const int nc = std::strcmp(hdr, "PF") == 0 ? 3 : 1;
        texelStore.emplace_back(new std::vector<unsigned char>((size_t)w * h * nc * sizeof(float)));
        float *px = (float *)texelStore.back()->data();
        std::vector<float> row((size_t)w * nc);
        for (int y = h - 1; y >= 0; --y) {   // bottom-to-top rows, scaled by |scale| (ReadPFM)
            if (std::fread(row.data(), 4, row.size(), f) != row.size()) { std::fclose(f); err = path + ": truncated"; return false; }
            for (size_t k = 0; k < row.size(); ++k) px[(size_t)y * w * nc + k] = row[k] * std::abs(sc);
        }
        std::fclose(f);
        t->format = BPT_IMAGE_F32; t->xres = w; t->yres = h; t->n_channels = nc; t->pixels = px;
        return true;
    }
    // FloatImageTexture::Create / SpectrumImageTexture::Create (textures.cpp:410-500) with TextureMapping2D "uv"
    bool makeTexture(const std::vector<std::string> &a, ParamList &pl) {
        if (a.size() < 3) { err = "Texture \"name\" \"float|spectrum\" \"imagemap\" expected"; return false; }
        if (a[1] != "float" && a[1] != "spectrum") { err = "texture value type must be float or spectrum"; return false; }
        const bool spec = a[1] == "spectrum";
        auto &names = spec ? spectrumTextures : floatTextures;
        // a parameter of a texture that is itself a float / spectrum texture: "texture <name>", or a constant wrapped in its own
        // constant-texture entry as TextureParameterDictionary::GetFloatTexture / GetSpectrumTexture do (paramdict.cpp)
        auto floatChild = [&](const std::string &name, float dflt, int *id) -> bool {
            Param *q = pl.find(name);
            if (q && q->type == "texture") {
                if (q->strs.empty() || !floatTextures.count(q->strs[0])) { err = "parameter \"" + name + "\": unknown float texture"; return false; }
                *id = floatTextures[q->strs[0]]; return true;
            }
            bpt_texture c; std::memset(&c, 0, sizeof(c));
            c.kind = BPT_TEXTURE_CONSTANT; c.value_f = (q && !q->nums.empty()) ? (float)q->nums[0] : dflt;
            textures.push_back(c); *id = (int)textures.size(); return true;
        };
        auto spectrumChild = [&](const std::string &name, float dflt, int *id) -> bool {
            Param *q = pl.find(name);
            if (q && q->type == "texture") {
                if (q->strs.empty() || !spectrumTextures.count(q->strs[0])) { err = "parameter \"" + name + "\": unknown spectrum texture"; return false; }
                *id = spectrumTextures[q->strs[0]]; return true;
            }
            bpt_texture c; std::memset(&c, 0, sizeof(c));
            c.kind = BPT_TEXTURE_CONSTANT; c.is_spectrum = 1;
            if (!getSpectrum(pl, name, &c.value_s)) return false;
            if (c.value_s.kind == BPT_SPEC_NONE) { c.value_s.kind = BPT_SPEC_CONSTANT; c.value_s.value = dflt; }
            textures.push_back(c); *id = (int)textures.size(); return true;
        };
        if (a[2] == "constant" || a[2] == "scale" || a[2] == "mix" || a[2] == "checkerboard") {
            bpt_texture t; std::memset(&t, 0, sizeof(t));
            t.is_spectrum = spec;
            if (a[2] == "checkerboard") {   // Float / SpectrumCheckerboardTexture::Create (textures.cpp:219-280), two-dimensional
                if (pl.i("dimension", 2) != 2) { err = "only the two-dimensional checkerboard is supported"; return false; }
                t.kind = BPT_TEXTURE_CHECKERBOARD;
                if (!textureMapping(pl, &t)) return false;
                if (!(spec ? spectrumChild("tex1", 1.f, &t.tex1) : floatChild("tex1", 1.f, &t.tex1)) || !(spec ? spectrumChild("tex2", 0.f, &t.tex2) : floatChild("tex2", 0.f, &t.tex2))) return false;
            } else
            if (a[2] == "constant") {
                t.kind = BPT_TEXTURE_CONSTANT;
                if (spec) { if (!getSpectrum(pl, "value", &t.value_s)) return false; if (t.value_s.kind == BPT_SPEC_NONE) { t.value_s.kind = BPT_SPEC_CONSTANT; t.value_s.value = 1.f; } }
                else t.value_f = pl.f("value", 1.f);
            } else if (a[2] == "scale") {   // Float / SpectrumScaledTexture::Create (textures.cpp:892-957)
                t.kind = BPT_TEXTURE_SCALE;
                if (!(spec ? spectrumChild("tex", 1.f, &t.tex1) : floatChild("tex", 1.f, &t.tex1)) || !floatChild("scale", 1.f, &t.tex2)) return false;
            } else {                        // Float / SpectrumMixTexture::Create (textures.cpp:533-551)
                t.kind = BPT_TEXTURE_MIX;
                if (!(spec ? spectrumChild("tex1", 0.f, &t.tex1) : floatChild("tex1", 0.f, &t.tex1)) ||
                    !(spec ? spectrumChild("tex2", 1.f, &t.tex2) : floatChild("tex2", 1.f, &t.tex2)) || !floatChild("amount", 0.5f, &t.tex3)) return false;
            }
            textures.push_back(t);
            names[a[0]] = (int)textures.size();
            return true;
        }
        if (a[2] != "imagemap") { err = "texture class \"" + a[2] + "\" is not supported by this front end (imagemap, constant, scale, mix, checkerboard)"; return false; }
        bpt_texture t; std::memset(&t, 0, sizeof(t));
        if (!textureMapping(pl, &t)) return false;
        t.max_anisotropy = pl.f("maxanisotropy", 8.f);
        std::string filter = pl.s("filter", "bilinear");
        if (filter == "ewa" || filter == "EWA") t.filter = BPT_TEXFILTER_EWA;
        else if (filter == "trilinear") t.filter = BPT_TEXFILTER_TRILINEAR;
        else if (filter == "bilinear") t.filter = BPT_TEXFILTER_BILINEAR;
        else if (filter == "point") t.filter = BPT_TEXFILTER_POINT;
        else { err = filter + ": filter function unknown"; return false; }
        std::string wrap = pl.s("wrap", "repeat");
        if (wrap == "clamp") t.wrap = BPT_TEXWRAP_CLAMP; else if (wrap == "repeat") t.wrap = BPT_TEXWRAP_REPEAT; else if (wrap == "black") t.wrap = BPT_TEXWRAP_BLACK;
        else { err = wrap + ": wrap mode unknown"; return false; }
        t.scale = pl.f("scale", 1.f); t.invert = pl.b("invert", false);
        std::string fn = pl.s("filename", "");
        if (fn.empty()) { err = "imagemap texture without a filename"; return false; }
        if (fn[0] != '/') fn = baseDir + "/" + fn;
        auto hasExt = [&](const char *e) { size_t n = std::strlen(e); if (fn.size() < n) return false; for (size_t i = 0; i < n; ++i) if (std::tolower((unsigned char)fn[fn.size() - n + i]) != e[i]) return false; return true; };
        const bool png = hasExt(".png");
        std::string enc = pl.s("encoding", png ? "sRGB" : "linear");
        if (enc == "sRGB") t.encoding = BPT_ENCODING_SRGB; else if (enc == "linear") t.encoding = BPT_ENCODING_LINEAR;
        else { err = "texture encoding \"" + enc + "\" is not supported (sRGB, linear)"; return false; }
        if (png) { if (!readPNG(fn, &t)) return false; }
        else if (hasExt(".pfm")) { if (!readPFMTexture(fn, &t)) return false; }
        else { err = fn + ": texture images are read from .png (8-bit) and .pfm files"; return false; }
        t.is_spectrum = spec;
        textures.push_back(t);
        names[a[0]] = (int)textures.size();
        return true;
    }

    bool makeMaterial(const std::string &type, ParamList &pl, bpt_material *m) {
        std::memset(m, 0, sizeof(*m));
        m->uroughness = m->vroughness = m->iface_uroughness = m->iface_vroughness = -1;
        if (!getDisplacement(pl, m)) return false;
        if (type == "diffuse") {
            m->type = BPT_MAT_DIFFUSE;
            return getSpectrumOrTexture(pl, "reflectance", &m->reflectance, &m->reflectance_tex);
        }
        if (type == "interface" || type == "none" || type == "") { m->type = BPT_MAT_INTERFACE; return true; }
        if (type == "conductor") {
            m->type = BPT_MAT_CONDUCTOR;
            if (!getSpectrum(pl, "eta", &m->eta) || !getSpectrum(pl, "k", &m->k) || !getSpectrumOrTexture(pl, "reflectance", &m->reflectance, &m->reflectance_tex)) return false;
            if (!getRoughness(pl, "", &m->roughness, &m->uroughness, &m->vroughness, &m->uroughness_tex, &m->vroughness_tex)) return false;
            m->remap_roughness = pl.b("remaproughness", true);
            return true;
        }
        if (type == "dielectric") {
            m->type = BPT_MAT_DIELECTRIC;
            if (!getSpectrum(pl, "eta", &m->eta)) return false;
            if (!getRoughness(pl, "", &m->roughness, &m->uroughness, &m->vroughness, &m->uroughness_tex, &m->vroughness_tex)) return false;
            m->remap_roughness = pl.b("remaproughness", true);
            return true;
        }
        if (type == "coatedconductor") {
            m->type = BPT_MAT_COATED_CONDUCTOR;
            if (!getSpectrum(pl, "interface.eta", &m->iface_eta) || !getSpectrum(pl, "conductor.eta", &m->conductor_eta) ||
                !getSpectrum(pl, "conductor.k", &m->conductor_k) || !getSpectrumOrTexture(pl, "reflectance", &m->reflectance, &m->reflectance_tex) ||
                !getSpectrum(pl, "albedo", &m->albedo)) return false;
            if (!getRoughness(pl, "interface.", &m->iface_roughness, &m->iface_uroughness, &m->iface_vroughness, nullptr, nullptr)) return false;
            if (!getRoughness(pl, "conductor.", &m->roughness, &m->uroughness, &m->vroughness, &m->uroughness_tex, &m->vroughness_tex)) return false;
            m->thickness = pl.f("thickness", .01f);
            m->g = pl.f("g", 0.f);
            m->max_depth = pl.i("maxdepth", 10); m->n_samples = pl.i("nsamples", 1);
            m->remap_roughness = pl.b("remaproughness", true);
            return true;
        }
        if (type == "coateddiffuse") {
            m->type = BPT_MAT_COATED_DIFFUSE;
            if (!getSpectrumOrTexture(pl, "reflectance", &m->reflectance, &m->reflectance_tex) || !getSpectrum(pl, "eta", &m->eta) || !getSpectrum(pl, "albedo", &m->albedo)) return false;
            if (!getRoughness(pl, "", &m->roughness, &m->uroughness, &m->vroughness, &m->uroughness_tex, &m->vroughness_tex)) return false;
            m->thickness = pl.f("thickness", .01f); m->g = pl.f("g", 0.f);
            m->max_depth = pl.i("maxdepth", 10); m->n_samples = pl.i("nsamples", 1);
            m->remap_roughness = pl.b("remaproughness", true);
            return true;
        }
        if (type == "diffusetransmission") {
            m->type = BPT_MAT_DIFFUSE_TRANSMISSION;
            if (!getSpectrumOrTexture(pl, "reflectance", &m->reflectance, &m->reflectance_tex) ||
                !getSpectrumOrTexture(pl, "transmittance", &m->transmittance, &m->transmittance_tex)) return false;
            m->scale = pl.f("scale", 1.f);
            return true;
        }
        if (type == "thindielectric") {
            m->type = BPT_MAT_THIN_DIELECTRIC;
            return getSpectrum(pl, "eta", &m->eta);
        }
        err = "material \"" + type + "\" is not supported by this front end";
        return false;
    }

    bool readPLY(const std::string &path, MeshRec *mr) {
        std::ifstream in(path, std::ios::binary);
        if (!in) { err = "cannot open PLY file " + path; return false; }
        std::string line;
        std::getline(in, line);
        if (line.substr(0, 3) != "ply") { err = path + ": not a PLY file"; return false; }
        bool binary = false;
        int nVerts = 0, nFaces = 0;
        struct Prop { std::string name, type, listCount, listItem; bool isList; };
        std::vector<Prop> vprops, fprops;
        int cur = 0;
        while (std::getline(in, line)) {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::istringstream ls(line);
            std::string w; ls >> w;
            if (w == "format") { std::string f; ls >> f; if (f == "binary_little_endian") binary = true; else if (f == "ascii") binary = false; else { err = path + ": unsupported PLY format " + f; return false; } }
            else if (w == "element") { std::string n; int c; ls >> n >> c; if (n == "vertex") { nVerts = c; cur = 1; } else if (n == "face") { nFaces = c; cur = 2; } else cur = 3; }
            else if (w == "property") {
                Prop p; std::string t; ls >> t;
                if (t == "list") { p.isList = true; ls >> p.listCount >> p.listItem >> p.name; } else { p.isList = false; p.type = t; ls >> p.name; }
                if (cur == 1) vprops.push_back(p); else if (cur == 2) fprops.push_back(p);
            } else if (w == "end_header") break;
        }
        auto tsize = [](const std::string &t) { if (t == "float" || t == "float32" || t == "int" || t == "int32" || t == "uint" || t == "uint32") return 4; if (t == "double" || t == "float64") return 8; if (t == "uchar" || t == "uint8" || t == "char" || t == "int8") return 1; if (t == "short" || t == "int16" || t == "ushort" || t == "uint16") return 2; return 0; };
        auto readNum = [&](const std::string &t, double *out) -> bool {
            if (!binary) { in >> *out; return (bool)in; }
            char b[8]; int sz = tsize(t); if (!sz) return false; in.read(b, sz); if (!in) return false;
            if (t == "float" || t == "float32") { float v; memcpy(&v, b, 4); *out = v; }
            else if (t == "double" || t == "float64") { double v; memcpy(&v, b, 8); *out = v; }
            else if (t == "int" || t == "int32") { int32_t v; memcpy(&v, b, 4); *out = v; }
            else if (t == "uint" || t == "uint32") { uint32_t v; memcpy(&v, b, 4); *out = v; }
            else if (t == "uchar" || t == "uint8") { *out = (unsigned char)b[0]; }
            else if (t == "char" || t == "int8") { *out = (signed char)b[0]; }
            else if (t == "short" || t == "int16") { int16_t v; memcpy(&v, b, 2); *out = v; }
            else { uint16_t v; memcpy(&v, b, 2); *out = v; }
            return true;
        };
        bool hasN = false, hasUV = false;
        for (auto &p : vprops) { if (p.name == "nx") hasN = true; if (p.name == "u" || p.name == "s") hasUV = true; }
        mr->P.resize(3 * (size_t)nVerts); if (hasN) mr->N.resize(3 * (size_t)nVerts); if (hasUV) mr->UV.resize(2 * (size_t)nVerts);
        if (binary) {
            // fast path: the body is read in one go and decoded with precomputed per-property offsets (the generic
            // stream reader below costs ~0.25 s per million triangles, which is half a bench step)
            // the body is decoded straight out of a read-only mapping of the file (no 19 MB staging vector to allocate, zero and fill)
            std::streampos here = in.tellg();
            struct Mapping {
                void *base = MAP_FAILED; size_t len = 0; int fd = -1;
                ~Mapping() { if (base != MAP_FAILED) munmap(base, len); if (fd >= 0) close(fd); }
            } map;
            map.fd = open(path.c_str(), O_RDONLY);
            struct stat sb;
            if (map.fd < 0 || fstat(map.fd, &sb) != 0 || (size_t)sb.st_size < (size_t)here) { err = path + ": cannot map PLY body"; return false; }
            map.len = (size_t)sb.st_size;
            map.base = mmap(nullptr, map.len, PROT_READ, MAP_PRIVATE | MAP_POPULATE, map.fd, 0);
            if (map.base == MAP_FAILED) { err = path + ": cannot map PLY body"; return false; }
            size_t nBytes = map.len - (size_t)here;
            const unsigned char *q = (const unsigned char *)map.base + (size_t)here, *end = q + nBytes;
            auto tcode = [&](const std::string &t) { if (t == "float" || t == "float32") return 0; if (t == "double" || t == "float64") return 1; if (t == "int" || t == "int32") return 2;
                                                      if (t == "uint" || t == "uint32") return 3; if (t == "uchar" || t == "uint8") return 4; if (t == "char" || t == "int8") return 5;
                                                      if (t == "short" || t == "int16") return 6; if (t == "ushort" || t == "uint16") return 7; return -1; };
            static const int tsz[8] = {4, 8, 4, 4, 1, 1, 2, 2};
            auto load = [](const unsigned char *b, int t) -> double {
                switch (t) {
                case 0: { float v; memcpy(&v, b, 4); return v; }
                case 1: { double v; memcpy(&v, b, 8); return v; }
                case 2: { int32_t v; memcpy(&v, b, 4); return v; }
                case 3: { uint32_t v; memcpy(&v, b, 4); return v; }
                case 4: return (unsigned char)b[0];
                case 5: return (signed char)b[0];
                case 6: { int16_t v; memcpy(&v, b, 2); return v; }
                default: { uint16_t v; memcpy(&v, b, 2); return v; }
                }
            };
            struct VP { int off, type, dest; };
            std::vector<VP> vp;
            int stride = 0;
            for (auto &pp : vprops) {
                if (pp.isList) { err = path + ": list property on vertices is not supported"; return false; }
                int t = tcode(pp.type); if (t < 0) { err = path + ": unknown PLY type " + pp.type; return false; }
                int dest = pp.name == "x" ? 0 : pp.name == "y" ? 1 : pp.name == "z" ? 2 : pp.name == "nx" ? 3 : pp.name == "ny" ? 4 : pp.name == "nz" ? 5 :
                           (pp.name == "u" || pp.name == "s") ? 6 : (pp.name == "v" || pp.name == "t") ? 7 : -1;
                if (dest >= 0) vp.push_back({stride, t, dest});
                stride += tsz[t];
            }
            if ((size_t)nVerts * stride > nBytes) { err = path + ": truncated vertex data"; return false; }
            const bool xyzOnly = stride == 12 && vp.size() == 3 && vp[0].off == 0 && vp[1].off == 4 && vp[2].off == 8 && vp[0].type == 0 &&
                                 vp[1].type == 0 && vp[2].type == 0 && vp[0].dest == 0 && vp[1].dest == 1 && vp[2].dest == 2;
            if (xyzOnly) {   // float x y z records: the vertex block is the P array
                memcpy(mr->P.data(), q, 12 * (size_t)nVerts);
                q += 12 * (size_t)nVerts;
            } else
            for (int v = 0; v < nVerts; ++v, q += stride)
                for (const VP &a : vp) {
                    float f = (float)load(q + a.off, a.type);
                    if (a.dest < 3) mr->P[3 * (size_t)v + a.dest] = f;
                    else if (a.dest < 6) mr->N[3 * (size_t)v + a.dest - 3] = f;
                    else mr->UV[2 * (size_t)v + a.dest - 6] = f;
                }
            struct FP { bool isList, isIdx; int type, cntType; };
            std::vector<FP> fp;
            for (auto &pp : fprops) {
                FP a; a.isList = pp.isList; a.isIdx = pp.name == "vertex_indices" || pp.name == "vertex_index";
                a.type = tcode(pp.isList ? pp.listItem : pp.type); a.cntType = pp.isList ? tcode(pp.listCount) : 0;
                if (a.type < 0 || a.cntType < 0) { err = path + ": unknown PLY type in face element"; return false; }
                fp.push_back(a);
            }
            if (fp.size() == 1 && fp[0].isList && fp[0].isIdx && fp[0].cntType == 4 && (fp[0].type == 2 || fp[0].type == 3)) {
                // the common layout "list uchar int vertex_indices": 13-byte records, decoded without the generic per-property walk
                if ((size_t)(end - q) < 13 * (size_t)nFaces) { err = path + ": truncated face data"; return false; }
                mr->idx.resize(3 * (size_t)nFaces);
                // (a million 13-byte records are ~3 ms on one core: split over a few threads for big meshes)
                const int nThreads = nFaces >= (1 << 18) ? 4 : 1;
                std::vector<int> bad(nThreads, 0);
                auto decode = [&](int t) {
                    const int f0 = (int)((long long)nFaces * t / nThreads), f1 = (int)((long long)nFaces * (t + 1) / nThreads);
                    const unsigned char *qq = q + 13 * (size_t)f0;
                    int *dst = mr->idx.data() + 3 * (size_t)f0;
                    for (int f = f0; f < f1; ++f, qq += 13, dst += 3) {
                        if (qq[0] != 3) { bad[t] = 1; return; }
                        memcpy(dst, qq + 1, 12);
                    }
                };
                if (nThreads == 1) decode(0);
                else {
                    std::vector<std::thread> pool;
                    for (int t = 1; t < nThreads; ++t) pool.emplace_back(decode, t);
                    decode(0);
                    for (auto &th : pool) th.join();
                }
                bool allTriangles = true;
                for (int b : bad) if (b) allTriangles = false;
                if (allTriangles) return true;
                mr->idx.clear();   // a face that is not a triangle: the generic walk below sorts triangles and quads
            }
            mr->idx.reserve(3 * (size_t)nFaces);
            for (int f = 0; f < nFaces; ++f)
                for (const FP &a : fp) {
                    if (!a.isList) { q += tsz[a.type]; continue; }
                    if (q + tsz[a.cntType] > end) { err = path + ": truncated face data"; return false; }
                    int cnt = (int)load(q, a.cntType); q += tsz[a.cntType];
                    if (q + (size_t)cnt * tsz[a.type] > end) { err = path + ": truncated face data"; return false; }
                    if (a.isIdx) {
                        if (cnt != 3 && cnt != 4) { err = path + ": only triangle and quad faces are supported"; return false; }
                        int v4[4];
                        for (int k = 0; k < cnt; ++k) v4[k] = (int)load(q + (size_t)k * tsz[a.type], a.type);
                        if (cnt == 3) { mr->idx.push_back(v4[0]); mr->idx.push_back(v4[1]); mr->idx.push_back(v4[2]); }
                        else { mr->quadIdx.push_back(v4[0]); mr->quadIdx.push_back(v4[1]); mr->quadIdx.push_back(v4[3]); mr->quadIdx.push_back(v4[2]); }   // util/mesh.cpp:302-305
                    }
                    q += (size_t)cnt * tsz[a.type];
                }
            return true;
        }
        for (int v = 0; v < nVerts; ++v)
            for (auto &p : vprops) {
                double d; if (!readNum(p.type, &d)) { err = path + ": truncated vertex data"; return false; }
                float f = (float)d;
                if (p.name == "x") mr->P[3 * v] = f; else if (p.name == "y") mr->P[3 * v + 1] = f; else if (p.name == "z") mr->P[3 * v + 2] = f;
                else if (p.name == "nx") mr->N[3 * v] = f; else if (p.name == "ny") mr->N[3 * v + 1] = f; else if (p.name == "nz") mr->N[3 * v + 2] = f;
                else if (p.name == "u" || p.name == "s") mr->UV[2 * v] = f; else if (p.name == "v" || p.name == "t") mr->UV[2 * v + 1] = f;
            }
        for (int f = 0; f < nFaces; ++f)
            for (auto &p : fprops) {
                if (!p.isList) { double d; if (!readNum(p.type, &d)) { err = path + ": truncated face data"; return false; } continue; }
                double cnt; if (!readNum(p.listCount, &cnt)) { err = path + ": truncated face data"; return false; }
                std::vector<int> v((int)cnt);
                for (int k = 0; k < (int)cnt; ++k) { double d; if (!readNum(p.listItem, &d)) { err = path + ": truncated face data"; return false; } v[k] = (int)d; }
                if (p.name == "vertex_indices" || p.name == "vertex_index") {
                    if (cnt == 3) mr->idx.insert(mr->idx.end(), v.begin(), v.end());
                    else if (cnt == 4) { mr->quadIdx.push_back(v[0]); mr->quadIdx.push_back(v[1]); mr->quadIdx.push_back(v[3]); mr->quadIdx.push_back(v[2]); }
                    else { err = path + ": only triangle and quad faces are supported"; return false; }
                }
            }
        return true;
    }

    // TextureMapping2D::Create (textures.cpp:49-73)
    bool textureMapping(ParamList &pl, bpt_texture *t) {
        std::string type = pl.s("mapping", "uv");
        t->su = pl.f("uscale", 1.f); t->sv = pl.f("vscale", 1.f); t->du = pl.f("udelta", 0.f); t->dv = pl.f("vdelta", 0.f);
        t->v1[0] = 1; t->v2[1] = 1;
        if (type == "uv") { t->mapping = BPT_TEXMAP_UV; return true; }
        if (type == "spherical") t->mapping = BPT_TEXMAP_SPHERICAL;
        else if (type == "cylindrical") t->mapping = BPT_TEXMAP_CYLINDRICAL;
        else if (type == "planar") {
            t->mapping = BPT_TEXMAP_PLANAR;
            if (Param *v = pl.find("v1")) { if (v->nums.size() != 3) { err = "planar mapping: bad \"v1\""; return false; } for (int k = 0; k < 3; ++k) t->v1[k] = (float)v->nums[k]; }
            if (Param *v = pl.find("v2")) { if (v->nums.size() != 3) { err = "planar mapping: bad \"v2\""; return false; } for (int k = 0; k < 3; ++k) t->v2[k] = (float)v->nums[k]; }
        } else { err = "2D texture mapping \"" + type + "\" unknown"; return false; }
        t->world_from_texture = ToB(gs.ctm.m); t->texture_from_world = ToB(gs.ctm.mInv);
        return true;
    }

    bool shape(const std::string &type, ParamList &pl) {
        std::unique_ptr<MeshRec> mr(new MeshRec);
        if (type == "trianglemesh") {
            Param *P = pl.find("P");
            if (!P || P->nums.size() % 3) { err = "trianglemesh: missing or bad \"P\""; return false; }
            mr->P.assign(P->nums.begin(), P->nums.end());
            Param *I = pl.find("indices");
            if (I) mr->idx.assign(I->nums.begin(), I->nums.end());
            else if (mr->P.size() == 9) mr->idx = {0, 1, 2};
            else { err = "trianglemesh: \"indices\" required"; return false; }
            if (Param *N = pl.find("N")) mr->N.assign(N->nums.begin(), N->nums.end());
            if (Param *U = pl.find("uv")) mr->UV.assign(U->nums.begin(), U->nums.end());
            if (!mr->N.empty() && mr->N.size() != mr->P.size()) { err = "trianglemesh: N size mismatch"; return false; }
            if (!mr->UV.empty() && mr->UV.size() / 2 != mr->P.size() / 3) { err = "trianglemesh: uv size mismatch"; return false; }
        } else if (type == "plymesh") {
            std::string fn = pl.s("filename", "");
            if (fn.empty()) { err = "plymesh: missing filename"; return false; }
            if (fn[0] != '/') fn = baseDir + "/" + fn;
            if (!readPLY(fn, mr.get())) return false;
        } else if (type == "bilinearmesh") {   // BilinearPatch::CreateMesh (shapes.cpp:910-999)
            Param *P = pl.find("P");
            if (!P || P->nums.size() % 3) { err = "bilinearmesh: missing or bad \"P\""; return false; }
            mr->P.assign(P->nums.begin(), P->nums.end());
            Param *I = pl.find("indices");
            if (I) mr->quadIdx.assign(I->nums.begin(), I->nums.end());
            else if (mr->P.size() == 12) mr->quadIdx = {0, 1, 2, 3};
            else { err = "bilinearmesh: \"indices\" required"; return false; }
            while (mr->quadIdx.size() % 4) mr->quadIdx.pop_back();
            if (Param *N = pl.find("N")) mr->N.assign(N->nums.begin(), N->nums.end());
            if (Param *U = pl.find("uv")) mr->UV.assign(U->nums.begin(), U->nums.end());
            if (!mr->N.empty() && mr->N.size() != mr->P.size()) { err = "bilinearmesh: N size mismatch"; return false; }
            if (!mr->UV.empty() && mr->UV.size() / 2 != mr->P.size() / 3) { err = "bilinearmesh: uv size mismatch"; return false; }
            if (pl.find("emissionfilename")) { err = "bilinearmesh: \"emissionfilename\" is not supported"; return false; }
        } else if (type == "sphere" || type == "disk" || type == "cylinder") {
            // Sphere / Disk / Cylinder::Create (shapes.cpp:72-83, 107-117, 135-146): the defaults are theirs
            mr->isQuadric = true;
            std::memset(&mr->q, 0, sizeof(mr->q));
            bpt_quadric &q = mr->q;
            q.radius = pl.f("radius", 1.f);
            q.phimax = pl.f("phimax", 360.f);
            if (type == "sphere") { q.type = BPT_SHAPE_SPHERE; q.zmin = pl.f("zmin", -q.radius); q.zmax = pl.f("zmax", q.radius); }
            else if (type == "cylinder") { q.type = BPT_SHAPE_CYLINDER; q.zmin = pl.f("zmin", -1.f); q.zmax = pl.f("zmax", 1.f); }
            else { q.type = BPT_SHAPE_DISK; q.height = pl.f("height", 0.f); q.innerradius = pl.f("innerradius", 0.f); }
            if (pl.find("alpha")) { err = "alpha on a " + type + " is not supported (triangle meshes only)"; return false; }
        } else { err = "shape \"" + type + "\" is not supported (trianglemesh, plymesh, bilinearmesh, sphere, disk, cylinder)"; return false; }
        std::memset(&mr->m, 0, sizeof(mr->m));
        if (Param *al = pl.find("alpha")) {   // getAlphaTexture (scene.cpp): "float alpha" < 1 -> FloatConstantTexture; texture-valued alpha: not implemented
            if (al->type != "float" || al->nums.size() != 1) { err = "texture-valued alpha is not supported (constant \"float alpha\" only)"; return false; }
            if ((float)al->nums[0] < 1.f) { mr->m.has_alpha = 1; mr->m.alpha = (float)al->nums[0]; }
            if (mr->m.has_alpha && !mr->quadIdx.empty()) { err = "alpha on bilinear patches is not supported"; return false; }
        }
        mr->m.n_vertices = (int)mr->P.size() / 3; mr->m.n_triangles = (int)mr->idx.size() / 3;
        mr->m.world_from_object = ToB(gs.ctm.m); mr->m.object_from_world = ToB(gs.ctm.mInv);
        mr->m.reverse_orientation = gs.reverseOrientation;
        mr->m.material = (materials[gs.material].type == BPT_MAT_INTERFACE) ? -1 : gs.material;
        mr->m.area_light = gs.areaLight;
        mr->m.medium_inside = mr->m.medium_outside = -1;
        if (!gs.mediumInside.empty()) { if (!namedMedia.count(gs.mediumInside)) { err = "unknown medium \"" + gs.mediumInside + "\""; return false; } mr->m.medium_inside = namedMedia[gs.mediumInside]; }
        if (!gs.mediumOutside.empty()) { if (!namedMedia.count(gs.mediumOutside)) { err = "unknown medium \"" + gs.mediumOutside + "\""; return false; } mr->m.medium_outside = namedMedia[gs.mediumOutside]; }
        if (curObject >= 0 && (mr->m.medium_inside >= 0 || mr->m.medium_outside >= 0)) { err = "media are not supported with object instancing"; return false; }
        mr->m.object_id = curObject;
        if (curObject >= 0 && gs.areaLight >= 0) { err = "area lights are not supported with object instancing"; return false; }
        meshes.push_back(std::move(mr));
        return true;
    }

    bool directive(Tokenizer &t, const std::string &d, std::vector<std::string> &args, ParamList &pl);
    bool parseFile(const std::string &path);
    bool finish();
};

bool Frontend::parseFile(const std::string &path) {
    Tokenizer t;
    if (!t.load(path)) { err = "cannot open " + path; return false; }
    std::string tok; bool q;
    bool have = t.next(&tok, &q);
    while (have) {
        if (q) return fail(t, "unexpected string \"" + tok + "\"");
        std::string d = tok;
        std::vector<std::string> args;   // positional args (numbers or the type string)
        ParamList pl;
        // collect until next bare identifier that starts with an uppercase letter
        have = t.next(&tok, &q);
        // positional part
        while (have) {
            if (q) {
                // either the directive's type string (first quoted, no space) or a param declaration "type name"
                if (tok.find(' ') == std::string::npos && pl.p.empty() && args.empty() && d != "Include" && d != "CoordinateSystem" && d != "CoordSysTransform" &&
                    d != "NamedMaterial" && d != "ObjectBegin" && d != "ObjectInstance" && d != "MediumInterface" && d != "ColorSpace" && d != "Option") { args.push_back(tok); }
                else if (tok.find(' ') == std::string::npos) args.push_back(tok);
                else {
                    Param p;
                    std::istringstream ss(tok); ss >> p.type >> p.name;
                    std::string v; bool vq;
                    if (!t.next(&v, &vq)) return fail(t, "premature EOF in parameter list");
                    auto addVal = [&](const std::string &s, bool sq) {
                        if (sq) p.strs.push_back(s);
                        else if (s == "true") p.bools.push_back(true);
                        else if (s == "false") p.bools.push_back(false);
                        else p.nums.push_back(std::strtod(s.c_str(), nullptr));
                    };
                    if (!vq && v == "[") {
                        while (true) {
                            if (!t.next(&v, &vq)) return fail(t, "premature EOF in parameter list");
                            if (!vq && v == "]") break;
                            addVal(v, vq);
                        }
                    } else addVal(v, vq);
                    if (p.type == "bool" && p.bools.empty() && !p.strs.empty()) { for (auto &s : p.strs) p.bools.push_back(s == "true"); }
                    pl.p.push_back(p);
                }
                have = t.next(&tok, &q);
            } else if (tok == "[") {
                while (true) { if (!(have = t.next(&tok, &q))) return fail(t, "premature EOF"); if (!q && tok == "]") break; args.push_back(tok); }
                have = t.next(&tok, &q);
            } else if (std::isupper((unsigned char)tok[0])) break;
            else { args.push_back(tok); have = t.next(&tok, &q); }
        }
        if (!directive(t, d, args, pl)) { if (err.find(t.file) == std::string::npos) fail(t, d + ": " + err); return false; }
        std::string un = pl.unused();
        if (!un.empty()) return fail(t, d + ": parameter \"" + un + "\" is not supported by this front end");
    }
    return true;
}

bool Frontend::directive(Tokenizer &t, const std::string &d, std::vector<std::string> &a, ParamList &pl) {
    auto num = [&](size_t i) { return i < a.size() ? (float)std::strtod(a[i].c_str(), nullptr) : 0.f; };
    auto need = [&](size_t n) { if (a.size() != n) { err = "expected " + std::to_string(n) + " arguments"; return false; } return true; };
    if (d == "LookAt") {
        if (!need(9)) return false;
        bool ok;
        Xform la = LookAt(Pt3(num(0), num(1), num(2)), Pt3(num(3), num(4), num(5)), Vec3(num(6), num(7), num(8)), &ok);
        if (!ok) { err = "up vector and viewing direction are collinear"; return false; }
        gs.ctm = gs.ctm * la; return true;
    }
    if (d == "Translate") { if (!need(3)) return false; gs.ctm = gs.ctm * Translate(Vec3(num(0), num(1), num(2))); return true; }
    if (d == "Scale") { if (!need(3)) return false; gs.ctm = gs.ctm * Scale(num(0), num(1), num(2)); return true; }
    if (d == "Rotate") { if (!need(4)) return false; gs.ctm = gs.ctm * Rotate(num(0), Vec3(num(1), num(2), num(3))); return true; }
    if (d == "Identity") { gs.ctm = Xform(); return true; }
    if (d == "Transform" || d == "ConcatTransform") {
        if (!need(16)) return false;
        // pbrt: Transpose(SquareMatrix<4>(pstd::MakeSpan(tr, 16)))
        M4 m;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = num(4 * j + i);
        if (d == "Transform") gs.ctm = Xform(m); else gs.ctm = gs.ctm * Xform(m);
        return true;
    }
    if (d == "CoordinateSystem") { if (a.empty()) { err = "name expected"; return false; } namedCS[a[0]] = gs.ctm; return true; }
    if (d == "CoordSysTransform") { if (a.empty() || !namedCS.count(a[0])) { err = "unknown coordinate system"; return false; } gs.ctm = namedCS[a[0]]; return true; }
    if (d == "ReverseOrientation") { gs.reverseOrientation = !gs.reverseOrientation; return true; }
    if (d == "AttributeBegin") { stack.push_back(gs); return true; }
    if (d == "AttributeEnd") { if (stack.empty()) { err = "unmatched AttributeEnd"; return false; } gs = stack.back(); stack.pop_back(); return true; }
    if (d == "TransformBegin") { xstack.push_back(gs.ctm); return true; }
    if (d == "TransformEnd") { if (xstack.empty()) { err = "unmatched TransformEnd"; return false; } gs.ctm = xstack.back(); xstack.pop_back(); return true; }
    if (d == "Option") { if (Param *s = pl.find("seed")) seedOption = (int)s->nums[0]; for (auto &q : pl.p) q.used = true; return true; }
    if (d == "ColorSpace") { if (a.empty() || a[0] != "srgb") { err = "only the srgb colour space is supported"; return false; } return true; }
    if (d == "Include") { if (a.empty()) { err = "filename expected"; return false; } std::string fn = a[0]; if (fn[0] != '/') fn = baseDir + "/" + fn; return parseFile(fn); }
    if (d == "Camera") {
        if (a.empty() || (a[0] != "perspective" && a[0] != "orthographic")) { err = "cameras supported: \"perspective\", \"orthographic\""; return false; }
        std::memset(&cam, 0, sizeof(cam));
        cam.type = a[0] == "perspective" ? BPT_CAMERA_PERSPECTIVE : BPT_CAMERA_ORTHOGRAPHIC;
        cam.camera_from_world = ToB(gs.ctm.m); cam.world_from_camera = ToB(gs.ctm.mInv);
        cam.fov = pl.f("fov", 90.f); cam.lens_radius = pl.f("lensradius", 0.f); cam.focal_distance = pl.f("focaldistance", 1e6f);
        cam.frame_aspect_ratio = pl.f("frameaspectratio", -1.f);
        if (Param *sw = pl.find("screenwindow")) { if (sw->nums.size() != 4) { err = "screenwindow needs four values"; return false; } cam.has_screen_window = 1; for (int k = 0; k < 4; ++k) cam.screen_window[k] = (float)sw->nums[k]; }
        cam.shutter_open = pl.f("shutteropen", 0.f); cam.shutter_close = pl.f("shutterclose", 1.f);
        namedCS["camera"] = Inverse(gs.ctm);
        cameraMediumName = gs.mediumOutside;   // BasicSceneBuilder::Camera: camera.medium = graphicsState.currentOutsideMedium
        haveCam = true; return true;
    }
    if (d == "Film") {
        if (a.empty() || (a[0] != "rgb" && a[0] != "gbuffer")) { err = "only rgb / gbuffer films are supported"; return false; }
        std::memset(&film, 0, sizeof(film));
        film.type = a[0] == "rgb" ? BPT_FILM_RGB : BPT_FILM_GBUFFER;
        film.xres = pl.i("xresolution", 1280); film.yres = pl.i("yresolution", 720);
        if (Param *pb = pl.find("pixelbounds")) { if (pb->nums.size() != 4) { err = "pixelbounds needs four values"; return false; } for (int k = 0; k < 4; ++k) film.pixel_bounds[k] = (int)pb->nums[k]; }
        if (Param *cw = pl.find("cropwindow")) {
            if (cw->nums.size() != 4) { err = "cropwindow needs four values"; return false; }
            // film.cpp:138-160
            float c[4]; for (int k = 0; k < 4; ++k) c[k] = (float)cw->nums[k];
            auto clamp01 = [](float v) { return v < 0 ? 0.f : (v > 1 ? 1.f : v); };
            float x0 = clamp01(std::min(c[0], c[1])), x1 = clamp01(std::max(c[0], c[1])), y0 = clamp01(std::min(c[2], c[3])), y1 = clamp01(std::max(c[2], c[3]));
            film.pixel_bounds[0] = (int)std::ceil(film.xres * x0); film.pixel_bounds[1] = (int)std::ceil(film.xres * x1);
            film.pixel_bounds[2] = (int)std::ceil(film.yres * y0); film.pixel_bounds[3] = (int)std::ceil(film.yres * y1);
        }
        {
            std::string csys = pl.s("coordinatesystem", "camera");
            if (csys != "camera" && csys != "world") { err = "gbuffer coordinatesystem must be \"camera\" or \"world\""; return false; }
            film.coordinate_system = csys == "world" ? 1 : 0;
        }
        film.iso = pl.f("iso", 100.f); film.white_balance = pl.f("whitebalance", 0.f); film.exposure_time = -1;
        film.max_component_value = pl.f("maxcomponentvalue", -1.f); film.diagonal = pl.f("diagonal", 35.f);
        std::string sensor = pl.s("sensor", "cie1931");
        if (sensor != "cie1931") { err = "only the cie1931 sensor is supported"; return false; }
        pl.s("filename", ""); pl.b("savefp16", true);
        return true;
    }
    if (d == "Sampler") {
        std::memset(&sampler, 0, sizeof(sampler));
        std::string ty = a.empty() ? "zsobol" : a[0];
        if (ty == "zsobol") sampler.type = BPT_SAMPLER_ZSOBOL; else if (ty == "independent") sampler.type = BPT_SAMPLER_INDEPENDENT;
        else if (ty == "paddedsobol") sampler.type = BPT_SAMPLER_PADDED_SOBOL; else if (ty == "stratified") sampler.type = BPT_SAMPLER_STRATIFIED;
        else if (ty == "sobol") sampler.type = BPT_SAMPLER_SOBOL; else if (ty == "halton") sampler.type = BPT_SAMPLER_HALTON;
        else if (ty == "pmj02bn") sampler.type = BPT_SAMPLER_PMJ02BN;
        else { err = "sampler \"" + ty + "\" is not supported (zsobol, sobol, paddedsobol, halton, pmj02bn, stratified, independent)"; return false; }
        sampler.spp = pl.i("pixelsamples", 16); sampler.seed = pl.i("seed", seedOption);
        if (sampler.type == BPT_SAMPLER_STRATIFIED) {   // StratifiedSampler::Create (samplers.cpp:297-319)
            sampler.jitter = pl.b("jitter", true);
            sampler.x_samples = pl.i("xsamples", 4); sampler.y_samples = pl.i("ysamples", 4);
            sampler.spp = sampler.x_samples * sampler.y_samples;
        }
        // (HaltonSampler::Create, samplers.cpp:77-88: permutedigits by default, no fastowen)
        std::string r = pl.s("randomization", sampler.type == BPT_SAMPLER_HALTON ? "permutedigits" : "fastowen");
        if (sampler.type == BPT_SAMPLER_HALTON && r == "fastowen") { err = "\"fastowen\" randomization not supported by Halton sampler"; return false; }
        sampler.randomize = r == "none" ? BPT_RANDOMIZE_NONE : r == "permutedigits" ? BPT_RANDOMIZE_PERMUTE_DIGITS : r == "owen" ? BPT_RANDOMIZE_OWEN : BPT_RANDOMIZE_FAST_OWEN;
        haveSampler = true; return true;
    }
    if (d == "PixelFilter") {
        std::memset(&filter, 0, sizeof(filter));
        std::string ty = a.empty() ? "gaussian" : a[0];
        if (ty == "gaussian") { filter.type = BPT_FILTER_GAUSSIAN; filter.xradius = pl.f("xradius", 1.5f); filter.yradius = pl.f("yradius", 1.5f); filter.sigma = pl.f("sigma", 0.5f); }
        else if (ty == "box") { filter.type = BPT_FILTER_BOX; filter.xradius = pl.f("xradius", 0.5f); filter.yradius = pl.f("yradius", 0.5f); }
        else if (ty == "mitchell") { filter.type = BPT_FILTER_MITCHELL; filter.xradius = pl.f("xradius", 2.f); filter.yradius = pl.f("yradius", 2.f); filter.b = pl.f("B", 1.f / 3.f); filter.c = pl.f("C", 1.f / 3.f); }
        else if (ty == "sinc") { filter.type = BPT_FILTER_SINC; filter.xradius = pl.f("xradius", 4.f); filter.yradius = pl.f("yradius", 4.f); filter.tau = pl.f("tau", 3.f); }
        else if (ty == "triangle") { filter.type = BPT_FILTER_TRIANGLE; filter.xradius = pl.f("xradius", 2.f); filter.yradius = pl.f("yradius", 2.f); }
        else { err = "pixel filter \"" + ty + "\" is unknown (box, gaussian, mitchell, sinc, triangle)"; return false; }
        haveFilter = true; return true;
    }
    if (d == "Integrator") {
        std::string ty = a.empty() ? "volpath" : a[0];
        // ("path" is a different estimator in the reference -- no media, no transmittance along shadow rays -- and is not substituted)
        if (ty != "volpath") { err = "integrator \"" + ty + "\" is not supported (volpath)"; return false; }
        integ.max_depth = pl.i("maxdepth", 5); integ.regularize = pl.b("regularize", false);
        std::string ls = pl.s("lightsampler", "bvh");
        integ.light_sampler = ls == "uniform" ? BPT_LIGHTSAMPLER_UNIFORM : ls == "bvh" ? BPT_LIGHTSAMPLER_BVH : -1;
        if (integ.light_sampler < 0) { err = "light sampler \"" + ls + "\" is not supported (bvh, uniform)"; return false; }
        haveInteg = true; return true;
    }
    if (d == "WorldBegin") { inWorld = true; gs.ctm = Xform(); namedCS["world"] = gs.ctm; return true; }
    if (d == "Texture") return makeTexture(a, pl);
    if (d == "Material") {
        bpt_material m;
        if (!makeMaterial(a.empty() ? "diffuse" : a[0], pl, &m)) return false;
        materials.push_back(m); gs.material = (int)materials.size() - 1; return true;
    }
    if (d == "MakeNamedMaterial") {
        if (a.empty()) { err = "name expected"; return false; }
        bpt_material m;
        if (!makeMaterial(pl.s("type", ""), pl, &m)) return false;
        materials.push_back(m); namedMaterials[a[0]] = (int)materials.size() - 1; return true;
    }
    if (d == "MediumInterface") {
        // MediumInterface "inside" ["outside"]: one name sets both (parser.cpp)
        if (a.empty()) { err = "medium name expected"; return false; }
        gs.mediumInside = a[0]; gs.mediumOutside = a.size() > 1 ? a[1] : a[0];
        return true;
    }
    if (d == "MakeNamedMedium") {
        if (a.empty()) { err = "name expected"; return false; }
        bpt_medium m; std::memset(&m, 0, sizeof(m));
        std::string type = pl.s("type", "");
        if (type == "uniformgrid") m.type = BPT_MEDIUM_UNIFORMGRID;
        else if (type == "homogeneous") m.type = BPT_MEDIUM_HOMOGENEOUS;
        else if (type == "rgbgrid") m.type = BPT_MEDIUM_RGBGRID;
        else { err = "medium type \"" + type + "\" is not supported (homogeneous, uniformgrid, rgbgrid)"; return false; }
        if (pl.find("preset")) { err = "medium presets are not supported"; return false; }
        if (m.type == BPT_MEDIUM_RGBGRID) {   // RGBGridMedium::Create (media.cpp:380-453): RGB arrays, one triple per voxel
            m.nx = pl.i("nx", 1); m.ny = pl.i("ny", 1); m.nz = pl.i("nz", 1);
            const size_t want = 3 * (size_t)m.nx * m.ny * m.nz;
            auto grid = [&](const char *name, const float **out) -> bool {
                *out = nullptr;
                Param *q = pl.find(name);
                if (!q) return true;
                if (q->type != "rgb" || q->nums.size() != want) { err = std::string("rgbgrid medium: \"") + name + "\" needs nx*ny*nz RGB values"; return false; }
                densityStore.emplace_back(new std::vector<float>(q->nums.begin(), q->nums.end()));
                *out = densityStore.back()->data();
                return true;
            };
            if (!grid("sigma_a", &m.sigma_a_rgb) || !grid("sigma_s", &m.sigma_s_rgb) || !grid("Le", &m.Le_rgb)) return false;
            if (!m.sigma_a_rgb && !m.sigma_s_rgb) { err = "RGB grid requires \"sigma_a\" and/or \"sigma_s\" parameter values"; return false; }
            if (m.Le_rgb && !m.sigma_a_rgb) { err = "RGB grid requires \"sigma_a\" if \"Le\" value provided"; return false; }
            m.scale = pl.f("scale", 1.f); m.g = pl.f("g", 0.f); m.Le_scale = pl.f("Lescale", 1.f);
            m.world_from_medium = ToB(gs.ctm.m); m.medium_from_world = ToB(gs.ctm.mInv);
            m.bounds[0] = m.bounds[1] = m.bounds[2] = 0; m.bounds[3] = m.bounds[4] = m.bounds[5] = 1;
            if (Param *q = pl.find("p0")) { if (q->nums.size() != 3) { err = "p0 needs three values"; return false; } for (int k = 0; k < 3; ++k) m.bounds[k] = (float)q->nums[k]; }
            if (Param *q = pl.find("p1")) { if (q->nums.size() != 3) { err = "p1 needs three values"; return false; } for (int k = 0; k < 3; ++k) m.bounds[3 + k] = (float)q->nums[k]; }
            media.push_back(m); namedMedia[a[0]] = (int)media.size() - 1; return true;
        }
        if (!getSpectrum(pl, "sigma_a", &m.sigma_a) || !getSpectrum(pl, "sigma_s", &m.sigma_s) || !getSpectrum(pl, "Le", &m.Le)) return false;
        m.scale = pl.f("scale", 1.f); m.g = pl.f("g", 0.f);
        m.Le_scale = m.type == BPT_MEDIUM_HOMOGENEOUS ? pl.f("Lescale", 1.f) : 1.f;
        m.world_from_medium = ToB(gs.ctm.m); m.medium_from_world = ToB(gs.ctm.mInv);
        m.bounds[0] = m.bounds[1] = m.bounds[2] = 0; m.bounds[3] = m.bounds[4] = m.bounds[5] = 1;
        if (m.type == BPT_MEDIUM_UNIFORMGRID) {
            Param *dn = pl.find("density");
            if (!dn || dn->nums.empty()) { err = "uniformgrid medium: no \"density\" values"; return false; }
            if (pl.find("temperature")) { err = "temperature grids are not supported"; return false; }
            m.nx = pl.i("nx", 1); m.ny = pl.i("ny", 1); m.nz = pl.i("nz", 1);
            if ((size_t)m.nx * m.ny * m.nz != dn->nums.size()) { err = "uniformgrid medium: density count != nx*ny*nz"; return false; }
            densityStore.emplace_back(new std::vector<float>(dn->nums.begin(), dn->nums.end()));
            m.density = densityStore.back()->data();
            if (Param *ls = pl.find("Lescale")) {   // GridMedium: a float array, one value per density sample (media.cpp:293-306)
                if (ls->nums.size() != dn->nums.size()) { err = "uniformgrid medium: \"Lescale\" count != nx*ny*nz"; return false; }
                densityStore.emplace_back(new std::vector<float>(ls->nums.begin(), ls->nums.end()));
                m.Le_scale_grid = densityStore.back()->data();
            }
            if (Param *q = pl.find("p0")) { if (q->nums.size() != 3) { err = "p0 needs three values"; return false; } for (int k = 0; k < 3; ++k) m.bounds[k] = (float)q->nums[k]; }
            if (Param *q = pl.find("p1")) { if (q->nums.size() != 3) { err = "p1 needs three values"; return false; } for (int k = 0; k < 3; ++k) m.bounds[3 + k] = (float)q->nums[k]; }
        }
        media.push_back(m); namedMedia[a[0]] = (int)media.size() - 1; return true;
    }
    if (d == "NamedMaterial") { if (a.empty() || !namedMaterials.count(a[0])) { err = "unknown named material"; return false; } gs.material = namedMaterials[a[0]]; return true; }
    if (d == "AreaLightSource") {
        if (a.empty() || a[0] != "diffuse") { err = "only \"diffuse\" area lights are supported"; return false; }
        bpt_area_light al; std::memset(&al, 0, sizeof(al));
        if (!getSpectrum(pl, "L", &al.L)) return false;
        al.scale = pl.f("scale", 1.f); al.two_sided = pl.b("twosided", false); al.power = pl.f("power", -1.f);
        if (pl.find("filename")) { err = "image area lights are not supported"; return false; }
        areaLights.push_back(al); gs.areaLight = (int)areaLights.size() - 1; return true;
    }
    if (d == "LightSource") {
        if (a.empty() || (a[0] != "infinite" && a[0] != "point" && a[0] != "spot" && a[0] != "distant")) {
            err = "light sources supported: \"infinite\", \"point\", \"spot\", \"distant\""; return false;
        }
        bpt_light l; std::memset(&l, 0, sizeof(l));
        l.world_from_light = ToB(gs.ctm.m); l.light_from_world = ToB(gs.ctm.mInv);
        l.scale = pl.f("scale", 1.f);
        if (a[0] != "infinite") {
            // PointLight / SpotLight / DistantLight::Create (lights.cpp:192-213, 1433-1464, 246-276)
            auto point = [&](const char *name, float *out, float dx, float dy, float dz) {
                out[0] = dx; out[1] = dy; out[2] = dz;
                Param *q = pl.find(name);
                if (q && q->nums.size() >= 3) for (int k = 0; k < 3; ++k) out[k] = (float)q->nums[k];
            };
            l.type = a[0] == "point" ? BPT_LIGHT_POINT : a[0] == "spot" ? BPT_LIGHT_SPOT : BPT_LIGHT_DISTANT;
            if (!getSpectrum(pl, a[0] == "distant" ? "L" : "I", &l.L)) return false;
            point("from", l.from, 0, 0, 0);
            l.to[2] = 1;
            if (a[0] != "point") point("to", l.to, 0, 0, 1);
            l.power = a[0] == "distant" ? -1.f : pl.f("power", -1.f);
            l.illuminance = a[0] == "distant" ? pl.f("illuminance", -1.f) : -1.f;
            if (a[0] == "spot") { l.cone_angle = pl.f("coneangle", 30.f); l.cone_delta = pl.f("conedeltaangle", 5.f); }
            lights.push_back(l); return true;
        }
        if (!getSpectrum(pl, "L", &l.L)) return false;
        l.illuminance = pl.f("illuminance", -1.f);
        std::string fn = pl.s("filename", "");
        if (fn.empty()) l.type = BPT_LIGHT_UNIFORM_INFINITE;
        else {
            l.type = BPT_LIGHT_IMAGE_INFINITE;
            if (fn[0] != '/') fn = baseDir + "/" + fn;
            // PFM reader (util/image.cpp ReadPFM): RGB, bottom-to-top rows
            FILE *f = std::fopen(fn.c_str(), "rb");
            if (!f) { err = "cannot open " + fn; return false; }
            char hdr[3]; int w, h; float sc;
            if (std::fscanf(f, "%2s %d %d %f", hdr, &w, &h, &sc) != 4 || std::strcmp(hdr, "PF") != 0 || w != h) { std::fclose(f); err = fn + ": square RGB PFM expected"; return false; }
            std::fgetc(f);
            imageStore.emplace_back(new std::vector<float>(3 * (size_t)w * h));
            std::vector<float> row(3 * (size_t)w);
            for (int y = h - 1; y >= 0; --y) {
                if (std::fread(row.data(), 4, row.size(), f) != row.size()) { std::fclose(f); err = fn + ": truncated"; return false; }
                for (size_t k = 0; k < row.size(); ++k) (*imageStore.back())[(size_t)y * 3 * w + k] = row[k] * std::abs(sc);
            }
            std::fclose(f);
            if (sc > 0) { err = fn + ": big-endian PFM not supported"; return false; }
            l.image_rgb = imageStore.back()->data(); l.image_res = w;
        }
        lights.push_back(l); return true;
    }
    if (d == "Shape") { if (a.empty()) { err = "shape type expected"; return false; } return shape(a[0], pl); }
    if (d == "ObjectBegin") { if (a.empty()) { err = "name expected"; return false; } if (curObject >= 0) { err = "nested ObjectBegin"; return false; } stack.push_back(gs); int id = (int)objectIds.size(); objectIds[a[0]] = id; curObject = id; return true; }
    if (d == "ObjectEnd") { if (curObject < 0) { err = "unmatched ObjectEnd"; return false; } curObject = -1; gs = stack.back(); stack.pop_back(); return true; }
    if (d == "ObjectInstance") {
        if (a.empty() || !objectIds.count(a[0])) { err = "unknown object"; return false; }
        bpt_instance in; in.object_id = objectIds[a[0]]; in.world_from_instance = ToB(gs.ctm.m); in.instance_from_world = ToB(gs.ctm.mInv); instances.push_back(in); return true;
    }
    err = "directive is not supported by this front end";
    (void)t;
    return false;
}

bool Frontend::finish() {
    if (!haveCam) { err = "scene has no Camera"; return false; }
    if (film.xres == 0) { film.type = BPT_FILM_RGB; film.xres = 1280; film.yres = 720; film.iso = 100; film.diagonal = 35; film.exposure_time = -1; film.max_component_value = -1; }
    auto ck = [&](int rc) { if (rc != BPT_OK) { err = bpt_last_error(e); return false; } return true; };
    if (!ck(bpt_scene_begin(e))) return false;
    if (!ck(bpt_scene_set_camera(e, &cam)) || !ck(bpt_scene_set_film(e, &film))) return false;
    if (haveFilter && !ck(bpt_scene_set_filter(e, &filter))) return false;
    if (haveSampler) {
        if (sppOverride > 0) {
            sampler.spp = sppOverride;
            if (sampler.type == BPT_SAMPLER_STRATIFIED) {   // --spp: the most square factorisation (samplers.cpp:302-313)
                int div = (int)std::sqrt((double)sppOverride);
                while (sppOverride % div) --div;
                sampler.x_samples = sppOverride / div; sampler.y_samples = sppOverride / sampler.x_samples;
            }
        }
        if (!ck(bpt_scene_set_sampler(e, &sampler))) return false;
    }
    else if (sppOverride > 0 || seedOption) { bpt_sampler s{BPT_SAMPLER_ZSOBOL, sppOverride > 0 ? sppOverride : 16, seedOption, BPT_RANDOMIZE_FAST_OWEN, 0, 0, 0}; if (!ck(bpt_scene_set_sampler(e, &s))) return false; }
    if (haveInteg && !ck(bpt_scene_set_integrator(e, &integ))) return false;
    if (!textures.empty() && !ck(bpt_scene_set_textures(e, textures.data(), (int)textures.size()))) return false;
    if (!ck(bpt_scene_set_materials(e, materials.data(), (int)materials.size()))) return false;
    if (!ck(bpt_scene_set_area_lights(e, areaLights.data(), (int)areaLights.size()))) return false;
    for (auto &m : media) if (!ck(bpt_scene_add_medium(e, &m))) return false;
    if (!cameraMediumName.empty()) {
        if (!namedMedia.count(cameraMediumName)) { err = "unknown camera medium"; return false; }
        if (!ck(bpt_scene_set_camera_medium(e, namedMedia[cameraMediumName]))) return false;
    }
    for (auto &m : meshes) {
        bpt_mesh bm = m->m;
        if (m->isQuadric) {
            bpt_quadric q = m->q;
            q.world_from_object = bm.world_from_object; q.object_from_world = bm.object_from_world; q.reverse_orientation = bm.reverse_orientation;
            q.material = bm.material; q.area_light = bm.area_light; q.medium_inside = bm.medium_inside; q.medium_outside = bm.medium_outside;
            q.object_id = bm.object_id;
            if (!ck(bpt_scene_add_quadric(e, &q))) return false;
            continue;
        }
        if (!m->quadIdx.empty()) {
            // Shape::Create: the triangles of a PLY file first, then its quads as a bilinear patch mesh over the same vertices (shapes.cpp:1435-1461)
            bm.P = m->P.data(); bm.N = m->N.empty() ? nullptr : m->N.data(); bm.UV = m->UV.empty() ? nullptr : m->UV.data();
            if (!m->idx.empty()) { bm.indices = m->idx.data(); if (!ck(bpt_scene_add_mesh(e, &bm))) return false; }
            bm.n_triangles = (int)m->quadIdx.size() / 4; bm.indices = m->quadIdx.data();
            if (!ck(bpt_scene_add_bilinear_mesh(e, &bm))) return false;
            continue;
        }
#ifdef BPT_FRONTEND_IN_LIBRARY
        if (!ck(bpt_internal_add_mesh_move(e, &bm, std::move(m->P), std::move(m->N), std::move(m->UV), std::move(m->idx)))) return false;
#else
        bm.P = m->P.data(); bm.N = m->N.empty() ? nullptr : m->N.data(); bm.UV = m->UV.empty() ? nullptr : m->UV.data(); bm.indices = m->idx.data();
        if (!ck(bpt_scene_add_mesh(e, &bm))) return false;
#endif
    }
    for (auto &in : instances) if (!ck(bpt_scene_add_instance(e, &in))) return false;
    for (auto &l : lights) if (!ck(bpt_scene_add_light(e, &l))) return false;
    return ck(bpt_scene_commit(e));
}

}  // namespace

static std::string g_frontendError;

extern "C" int bpt_load_pbrt_ex(bpt_engine *e, const char *path, int spp_override, int seed) {
    if (!e || !path) return BPT_ERR_INVALID;
    Frontend fe;
    fe.e = e;
    std::string p = path;
    size_t s = p.find_last_of('/');
    fe.baseDir = s == std::string::npos ? "." : p.substr(0, s);
    fe.seedOption = seed; fe.sppOverride = spp_override;
    // default material (scene.cpp: graphicsState.currentMaterialIndex = 0 -> "diffuse")
    bpt_material dm; std::memset(&dm, 0, sizeof(dm)); dm.type = BPT_MAT_DIFFUSE;
    fe.materials.push_back(dm);
    fe.integ.max_depth = 5; fe.integ.light_sampler = BPT_LIGHTSAMPLER_BVH;
    const bool timing = std::getenv("BPT_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    if (!fe.parseFile(p)) { g_frontendError = fe.err; return BPT_ERR_INVALID; }
    auto t1 = std::chrono::steady_clock::now();
    if (!fe.finish()) { g_frontendError = fe.err; return BPT_ERR_INVALID; }
    if (timing) fprintf(stderr, "[bpt load] parse %.2f ms, scene intake + commit %.2f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    return BPT_OK;
}
extern "C" int bpt_load_pbrt(bpt_engine *e, const char *path) { return bpt_load_pbrt_ex(e, path, 0, 0); }
extern "C" const char *bpt_frontend_error(void) { return g_frontendError.c_str(); }
