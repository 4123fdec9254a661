// exr_writer.h -- minimal OpenEXR 2 writer (host): single-part scanline file, FLOAT channels, no compression.
//
// What it replaces: Image::WriteEXR (util/image.cpp:1055-1253, which goes through the OpenEXR library) as used by
// RGBFilm::WriteImage / GBufferFilm::WriteImage (film.cpp:533-565, 688-803) for the multi-channel GBuffer hand-off to a denoiser.
// The layout follows the published OpenEXR file format ("OpenEXR File Layout"): magic 20000630, version 2, the header attributes
// every reader requires (channels, compression, dataWindow, displayWindow, lineOrder, pixelAspectRatio, screenWindowCenter,
// screenWindowWidth), the scan-line offset table, then one chunk per scan line {y, byte count, channel-planar pixel data}.  Channels
// are stored in alphabetical order, as the format demands (the reference's 25 GBuffer channel names sort to
// Albedo.B, Albedo.G, Albedo.R, B, G, N.X, ... -- readers look channels up by name).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace bpt {

struct ExrAttr { std::string name, type; std::vector<unsigned char> value; };

inline bool WriteEXR(const char *path, int width, int height, const std::vector<std::string> &names, const float *interleaved,
                     const std::vector<ExrAttr> &extra, std::string *err) {
    const int nc = (int)names.size();
    if (width <= 0 || height <= 0 || nc <= 0 || !interleaved) { if (err) *err = "WriteEXR: empty image"; return false; }
    std::vector<int> order(nc);
    for (int i = 0; i < nc; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return names[a] < names[b]; });
    std::vector<unsigned char> hdr;
    auto put = [&](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; hdr.insert(hdr.end(), b, b + n); };
    auto putStr = [&](const std::string &s) { put(s.c_str(), s.size() + 1); };
    auto putI = [&](int32_t v) { put(&v, 4); };
    auto putF = [&](float v) { put(&v, 4); };
    auto attr = [&](const char *name, const char *type, const std::vector<unsigned char> &v) { putStr(name); putStr(type); putI((int32_t)v.size()); put(v.data(), v.size()); };
    putI(20000630); putI(2);
    {
        std::vector<unsigned char> ch;
        auto cput = [&](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; ch.insert(ch.end(), b, b + n); };
        for (int k = 0; k < nc; ++k) {
            const std::string &nm = names[order[k]];
            cput(nm.c_str(), nm.size() + 1);
            int32_t type = 2 /* FLOAT */, one = 1; unsigned char lin[4] = {0, 0, 0, 0};
            cput(&type, 4); cput(lin, 4); cput(&one, 4); cput(&one, 4);
        }
        unsigned char z = 0; cput(&z, 1);
        attr("channels", "chlist", ch);
    }
    attr("compression", "compression", {0});
    {
        int32_t box[4] = {0, 0, width - 1, height - 1};
        std::vector<unsigned char> b((unsigned char *)box, (unsigned char *)box + 16);
        attr("dataWindow", "box2i", b); attr("displayWindow", "box2i", b);
    }
    attr("lineOrder", "lineOrder", {0});
    { float v = 1; attr("pixelAspectRatio", "float", std::vector<unsigned char>((unsigned char *)&v, (unsigned char *)&v + 4)); }
    { float v[2] = {0, 0}; attr("screenWindowCenter", "v2f", std::vector<unsigned char>((unsigned char *)v, (unsigned char *)v + 8)); }
    { float v = 1; attr("screenWindowWidth", "float", std::vector<unsigned char>((unsigned char *)&v, (unsigned char *)&v + 4)); }
    for (const ExrAttr &a : extra) attr(a.name.c_str(), a.type.c_str(), a.value);
    (void)putF;
    hdr.push_back(0);
    FILE *f = fopen(path, "wb");
    if (!f) { if (err) *err = std::string("WriteEXR: cannot open ") + path; return false; }
    const uint64_t lineBytes = (uint64_t)nc * width * 4;
    const uint64_t tableStart = hdr.size(), dataStart = tableStart + 8ull * height;
    std::vector<uint64_t> table(height);
    for (int y = 0; y < height; ++y) table[y] = dataStart + (uint64_t)y * (8 + lineBytes);
    bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size() && fwrite(table.data(), 8, height, f) == (size_t)height;
    std::vector<float> line((size_t)nc * width);
    for (int y = 0; y < height && ok; ++y) {
        for (int k = 0; k < nc; ++k) {
            const float *src = interleaved + (size_t)y * width * nc + order[k];
            float *dst = line.data() + (size_t)k * width;
            for (int x = 0; x < width; ++x) dst[x] = src[(size_t)x * nc];
        }
        int32_t head[2] = {y, (int32_t)lineBytes};
        ok = fwrite(head, 4, 2, f) == 2 && fwrite(line.data(), 4, line.size(), f) == line.size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok && err) *err = std::string("WriteEXR: write error on ") + path;
    return ok;
}

}  // namespace bpt
