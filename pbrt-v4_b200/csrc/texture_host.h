// texture_host.h -- host side of the image textures: the image as the ABI hands it over (8-bit with a colour encoding, or float),
// and the MIP pyramid the reference builds from it, restated so that every level holds the reference's texel values.
//   Image::GeneratePyramid                      util/image.cpp:313-383
//   Image::FloatResizeUp / ResampleWeights      util/image.cpp:208-311   (images whose sides are not powers of two)
//   Image::GetChannel / CopyRectIn (U256)       util/image.cpp:424-470, 710-737
//   sRGBColorEncoding, LinearToSRGB8            util/color.cpp, util/color.h:497-518
//   MIPMap::CreateFromFile (RGBA with alpha == 1 everywhere drops the alpha channel)  util/mipmap.cpp:349-381
// Levels are returned decoded to float: for 8-bit images every level is first quantised through the encoding exactly as the
// reference stores it (the pyramid keeps the original pixel format), then decoded through the 256-entry table.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "pmath.h"

namespace bpt {

struct HostImage {
    int xres = 0, yres = 0, nc = 0;
    bool is8 = false;             // U256 with the sRGB encoding (encoding "linear" on 8-bit data: srgb = false)
    bool srgb = true;
    std::vector<uint8_t> p8;
    std::vector<float> p32;
};

struct HostMIPLevel { int xres, yres; std::vector<float> texels; };

namespace texhost {

inline float EvalPoly(float t, const float *c, int n) {   // EvaluatePolynomial(t, c0, c1, ...) = FMA(t, EvaluatePolynomial(t, c1, ...), c0)
    float r = c[n - 1];
    for (int i = n - 2; i >= 0; --i) r = FMA(t, r, c[i]);
    return r;
}
inline float LinearToSRGB(float value) {
    if (value <= 0.0031308f) return 12.92f * value;
    float sqrtValue = SafeSqrt(value);
    const float pc[6] = {-0.0016829072605308378f, 0.03453868659826638f, 0.7642611304733891f, 2.0041169284241644f, 0.7551545191665577f,
                         -0.016202083165206348f};
    const float qc[6] = {4.178892964897981e-7f, -0.00004375359692957097f, 0.03467195408529984f, 0.6085338522168684f, 1.8970238036421054f, 1.f};
    float p = EvalPoly(sqrtValue, pc, 6), q = EvalPoly(sqrtValue, qc, 6);
    return p / q * value;
}
inline uint8_t LinearToSRGB8(float value) {
    if (value <= 0) return 0;
    if (value >= 1) return 255;
    float r = std::round(255.f * LinearToSRGB(value) + 0.f);
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
// LinearColorEncoding::FromLinear: uint8_t(Clamp(v * 255.f + .5f, 0, 255))
inline uint8_t LinearTo8(float v) { float r = v * 255.f + .5f; return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r)); }

inline int Mod(int a, int b) { int r = a - (a / b) * b; return r < 0 ? r + b : r; }
// wrap: 0 black, 1 clamp, 2 repeat (WrapMode, util/image.h:53)
inline bool Remap(int *x, int *y, int xres, int yres, int wrap) {
    int *p[2] = {x, y};
    const int res[2] = {xres, yres};
    for (int c = 0; c < 2; ++c) {
        if (*p[c] >= 0 && *p[c] < res[c]) continue;
        if (wrap == 2) *p[c] = Mod(*p[c], res[c]);
        else if (wrap == 1) *p[c] = *p[c] < 0 ? 0 : (*p[c] > res[c] - 1 ? res[c] - 1 : *p[c]);
        else return false;
    }
    return true;
}

struct Pyramid {
    const float *srgbLUT;   // SRGBToLinearLUT (tables blob "srgb8_to_linear")
    float Decode(const HostImage &im, uint8_t v) const { return im.srgb ? srgbLUT[v] : v / 255.f; }
    float GetChannel(const HostImage &im, int x, int y, int c, int wrap) const {
        if (!Remap(&x, &y, im.xres, im.yres, wrap)) return 0;
        size_t o = ((size_t)y * im.xres + x) * im.nc + c;
        return im.is8 ? Decode(im, im.p8[o]) : im.p32[o];
    }
    static float WindowedSinc(float x, float radius, float tau) {
        if (std::abs(x) > radius) return 0;
        auto sinc = [](float v) { v = kPi * v; if (1 - v * v == 1) return 1.f; return std::sin(v) / v; };
        return sinc(x) * sinc(x / tau);
    }
    struct RW { int firstPixel; float weight[4]; };
    static std::vector<RW> ResampleWeights(int oldRes, int newRes) {
        std::vector<RW> wt(newRes);
        float filterRadius = 2, tau = 2;
        for (int i = 0; i < newRes; ++i) {
            float center = (i + .5f) * oldRes / newRes;
            wt[i].firstPixel = (int)std::floor((center - filterRadius) + 0.5f);
            for (int j = 0; j < 4; ++j) {
                float pos = wt[i].firstPixel + j + .5f;
                wt[i].weight[j] = WindowedSinc(pos - center, filterRadius, tau);
            }
            float invSumWts = 1 / (wt[i].weight[0] + wt[i].weight[1] + wt[i].weight[2] + wt[i].weight[3]);
            for (int j = 0; j < 4; ++j) wt[i].weight[j] *= invSumWts;
        }
        return wt;
    }
    // FloatResizeUp: x pass into a buffer of the source rows, then the y pass with max(0, .) -- per output pixel the arithmetic does
    // not depend on the reference's tiling
    std::vector<float> ResizeUp(const HostImage &im, int nx, int ny, int wrap) const {
        std::vector<RW> xw = ResampleWeights(im.xres, nx), yw = ResampleWeights(im.yres, ny);
        const int nc = im.nc;
        const int yLo = yw[0].firstPixel, yHi = yw[ny - 1].firstPixel + 4;   // source rows touched (may lie outside: wrap mode)
        std::vector<float> xbuf((size_t)(yHi - yLo) * nx * nc);
        for (int ys = yLo; ys < yHi; ++ys)
            for (int xo = 0; xo < nx; ++xo)
                for (int c = 0; c < nc; ++c) {
                    const RW &r = xw[xo];
                    xbuf[((size_t)(ys - yLo) * nx + xo) * nc + c] =
                        r.weight[0] * GetChannel(im, r.firstPixel, ys, c, wrap) + r.weight[1] * GetChannel(im, r.firstPixel + 1, ys, c, wrap) +
                        r.weight[2] * GetChannel(im, r.firstPixel + 2, ys, c, wrap) + r.weight[3] * GetChannel(im, r.firstPixel + 3, ys, c, wrap);
                }
        std::vector<float> out((size_t)nx * ny * nc);
        for (int yo = 0; yo < ny; ++yo)
            for (int xo = 0; xo < nx; ++xo)
                for (int c = 0; c < nc; ++c) {
                    const RW &r = yw[yo];
                    auto X = [&](int k) { return xbuf[((size_t)(r.firstPixel + k - yLo) * nx + xo) * nc + c]; };
                    out[((size_t)yo * nx + xo) * nc + c] = std::max(0.f, (r.weight[0] * X(0) + r.weight[1] * X(1) + r.weight[2] * X(2) + r.weight[3] * X(3)));
                }
        return out;
    }
    // What a float scanline becomes once stored in a level of the original pixel format and read back
    void StoreLevel(const HostImage &orig, const std::vector<float> &img, int xres, int yres, HostMIPLevel *lv) const {
        lv->xres = xres; lv->yres = yres;
        lv->texels.resize(img.size());
        if (!orig.is8) { lv->texels = img; return; }
        for (size_t i = 0; i < img.size(); ++i) {
            uint8_t q = orig.srgb ? LinearToSRGB8(img[i]) : LinearTo8(img[i]);
            lv->texels[i] = Decode(orig, q);
        }
    }
    std::vector<HostMIPLevel> Generate(const HostImage &orig, int wrap) const {
        const int nc = orig.nc;
        auto isPow2 = [](int v) { return v && !(v & (v - 1)); };
        auto roundUpPow2 = [](int v) { int p = 1; while (p < v) p <<= 1; return p; };
        int xres = orig.xres, yres = orig.yres;
        std::vector<float> image;
        if (!isPow2(xres) || !isPow2(yres)) {
            int nx = roundUpPow2(xres), ny = roundUpPow2(yres);
            image = ResizeUp(orig, nx, ny, wrap);
            xres = nx; yres = ny;
        } else if (orig.is8) {
            image.resize(orig.p8.size());
            for (size_t i = 0; i < image.size(); ++i) image[i] = Decode(orig, orig.p8[i]);
        } else
            image = orig.p32;
        int nLevels = 1;
        for (int m = std::max(xres, yres); m > 1; m >>= 1) ++nLevels;   // 1 + Log2Int(max(res))
        std::vector<HostMIPLevel> pyr(nLevels);
        for (int i = 0; i < nLevels - 1; ++i) {
            StoreLevel(orig, image, xres, yres, &pyr[i]);
            int nxn = std::max(1, xres / 2), nyn = std::max(1, yres / 2);
            std::vector<float> next((size_t)nxn * nyn * nc);
            int srcDeltas[4] = {0, nc, nc * xres, nc * (xres + 1)};
            if (xres == 1) { srcDeltas[1] = 0; srcDeltas[3] -= nc; }
            if (yres == 1) { srcDeltas[2] = 0; srcDeltas[3] -= nc * xres; }
            for (int y = 0; y < nyn; ++y) {
                size_t srcOffset = (size_t)(2 * y) * xres * nc, nextOffset = (size_t)y * nxn * nc;
                for (int x = 0; x < nxn; ++x, srcOffset += nc)
                    for (int c = 0; c < nc; ++c, ++srcOffset, ++nextOffset)
                        next[nextOffset] = (image[srcOffset] + image[srcOffset + srcDeltas[1]] + image[srcOffset + srcDeltas[2]] + image[srcOffset + srcDeltas[3]]) / 4;
            }
            image.swap(next);
            xres = nxn; yres = nyn;
        }
        StoreLevel(orig, image, 1, 1, &pyr[nLevels - 1]);
        return pyr;
    }
};

}  // namespace texhost

// MIPMap::CreateFromFile's channel selection: RGBA whose alpha is 1 everywhere becomes RGB
inline void DropOpaqueAlpha(HostImage *im) {
    if (im->nc != 4) return;
    size_t n = (size_t)im->xres * im->yres;
    for (size_t i = 0; i < n; ++i) {
        float a = im->is8 ? (im->srgb ? (im->p8[4 * i + 3] == 255 ? 1.f : 0.f) : im->p8[4 * i + 3] / 255.f) : im->p32[4 * i + 3];
        if (a != 1) return;
    }
    if (im->is8) { std::vector<uint8_t> o(3 * n); for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) o[3 * i + c] = im->p8[4 * i + c]; im->p8.swap(o); }
    else { std::vector<float> o(3 * n); for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) o[3 * i + c] = im->p32[4 * i + c]; im->p32.swap(o); }
    im->nc = 3;
}

}  // namespace bpt
