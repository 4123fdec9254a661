// tables.h -- loader for the standard data tables blob (pbrt-v4_b200/data/pbrt_tables.bin): CIE 1931 matching
// functions, D65, sRGB matrices, the sRGB RGB->spectrum coefficient table, Sobol' generator matrices and the
// named metal / glass spectra.  The blob is produced by oracle/ref_dump.cpp from the reference build so the
// values are the reference's own (the rgb2spec table must not be regenerated differently).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace bpt {

struct Tables {
    std::map<std::string, std::vector<float>> f32;
    std::map<std::string, std::vector<uint32_t>> u32;
    bool Load(const std::string &path, std::string *err) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) { *err = "cannot open tables blob " + path; return false; }
        char magic[8];
        if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "BPTTBL1", 7) != 0) { *err = "bad tables blob"; std::fclose(f); return false; }
        while (true) {
            char nm[33]; uint32_t dtype; uint64_t count;
            if (std::fread(nm, 1, 32, f) != 32) break;
            nm[32] = 0;
            if (std::fread(&dtype, 4, 1, f) != 1 || std::fread(&count, 8, 1, f) != 1) break;
            if (dtype == 0) {
                std::vector<float> v(count);
                if (std::fread(v.data(), 4, count, f) != count) { *err = "truncated tables blob"; std::fclose(f); return false; }
                f32[nm] = std::move(v);
            } else {
                std::vector<uint32_t> v(count);
                if (std::fread(v.data(), 4, count, f) != count) { *err = "truncated tables blob"; std::fclose(f); return false; }
                u32[nm] = std::move(v);
            }
        }
        std::fclose(f);
        const char *need[] = {"cie_x", "cie_y", "cie_z", "srgb_illuminant", "srgb_rgb_from_xyz", "rgb2spec_srgb_scale", "rgb2spec_srgb_coeffs"};
        for (const char *n : need) if (!f32.count(n)) { *err = std::string("tables blob lacks ") + n; return false; }
        if (!u32.count("sobol32_2x52")) { *err = "tables blob lacks sobol32_2x52"; return false; }
        return true;
    }
    const std::vector<float> &F(const char *n) const { return f32.at(n); }
};

}  // namespace bpt
