#!/usr/bin/env python3
"""Build recipe for oracle/_ref: the UNMODIFIED reference CPU renderer (pbrt-v4 VolPathIntegrator).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pbrt-v4_b200/) may use what this builds.

What it does (and does not do):
  * compiles the reference's own sources *where they lie* under /root/reference with g++ via a
    ninja file this script writes (the reference's CMake build is NOT run: it needs OpenGL,
    and a network clone of Imath/libdeflate for OpenEXR -- SURVEY.md Appendix B);
  * generated code is produced by the reference's own generators, compiled from their sources:
    `soac` (src/pbrt/cmd/soac.cpp -> pbrt_soa.h, wavefront_workitems_soa.h; CMakeLists.txt:824-841)
    and `rgb2spec_opt` (src/pbrt/cmd/rgb2spec_opt.cpp -> rgbspectrum_*.cpp; CMakeLists.txt:950-975);
  * three translation units cannot be compiled as they are here (no GL, no OpenEXR/Imath).  They are
    patched *on a scratch copy* written to oracle/_ref/patched/ (git-ignored):
      src/pbrt/util/gui.cpp    -> 10-line stub (GUI is interactive-only)
      src/pbrt/util/image.cpp  -> OpenEXR reader/writer block cut to two stubs (PFM/PNG/QOI remain)
      src/pbrt/pbrt.cpp        -> Imf::setGlobalThreadCount line dropped
    No reference source is copied into the tracked repo.
  * float semantics: -ffp-contract=off (gcc would otherwise fuse a*b+c with -mfma; the reference only
    forces this for clang >= 14, CMakeLists.txt:146-149) so that the oracle is plain IEEE + explicit
    std::fma -- the semantics the CUDA engine restates with -fmad=false + explicit fmaf.
  * -march=x86-64-v3 (AVX2+FMA, hardware std::fma) rather than -march=native so the binary also runs
    on the GPU box's host CPU.

Outputs (all under oracle/_ref/): bin/pbrt, bin/imgtool, lib/libpbrt.a + ext libs, gen/*.h, and
bin/ref_dump (oracle/ref_dump.cpp linked against the reference: golden-vector generator).

Usage: python oracle/build_ref.py [-j N] [--only-ninja]
"""
import argparse
import os
import re
import subprocess
import sys

REF = os.environ.get("PBRT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

PBRT_SRC = """
bsdf bssrdf bxdfs cameras film filters interaction lights lightsamplers materials media options
paramdict parser pbrt ray samplers scene shapes textures
cpu/aggregates cpu/integrators cpu/primitive cpu/render
wavefront/aggregate wavefront/camera wavefront/film wavefront/integrator wavefront/media
wavefront/samples wavefront/surfscatter wavefront/subsurface wavefront/wavefront
util/args util/bluenoise util/buffercache util/check util/color util/colorspace util/display
util/error util/file util/float util/gui util/image util/log util/loopsubdiv util/lowdiscrepancy
util/math util/memory util/mesh util/mipmap util/noise util/parallel util/pmj02tables util/primes
util/print util/progressreporter util/pstd util/rng util/sampling util/scattering
util/sobolmatrices util/spectrum util/stats util/stbimage util/string util/transform util/vecmath
""".split()
PATCHED = {"util/gui", "util/image", "pbrt"}

DEFS = ["PBRT_IS_LINUX", "PBRT_HAVE_MMAP", "PBRT_NOINLINE=__attribute__((noinline))",
        "PBRT_RESTRICT=__restrict__", "PBRT_HAVE_POSIX_MEMALIGN", "PTEX_STATIC", "NDEBUG",
        "UTF8PROC_STATIC"]

GUI_STUB = r'''// stub written by oracle/build_ref.py: interactive GUI is unavailable in the oracle build
#define GLFW_INCLUDE_NONE
#include <pbrt/util/gui.h>
#include <pbrt/util/error.h>
namespace pbrt {
GUI::GUI(std::string, Vector2i, Bounds3f) { ErrorExit("GUI unavailable in this build"); }
GUI::~GUI() {}
DisplayState GUI::RefreshDisplay() { return EXIT; }
void GUI::Initialize() {}
Point2i GUI::GetResolution() { return Point2i(0, 0); }
}
'''

EXR_STUB = '''// OpenEXR stubbed by oracle/build_ref.py
static ImageAndMetadata ReadEXR(const std::string &name, Allocator alloc) {
    ErrorExit("%s: EXR unsupported in this build", name);
    return {};
}
// No OpenEXR here: "*.exr" outputs are written as a raw dump instead (the only multi-channel writer the oracle needs,
// for GBufferFilm): text header "BPTRAW1 <width> <height> <nchannels>\\n" + one channel name per line, then float32
// pixels, row-major, channels interleaved.
bool Image::WriteEXR(const std::string &name, const ImageMetadata &metadata) const {
    FILE *f = fopen(name.c_str(), "wb");
    if (!f) { Error("%s: cannot open for writing", name.c_str()); return false; }
    fprintf(f, "BPTRAW1 %d %d %d\\n", resolution.x, resolution.y, NChannels());
    for (int c = 0; c < NChannels(); ++c) fprintf(f, "%s\\n", channelNames[c].c_str());
    for (int y = 0; y < resolution.y; ++y)
        for (int x = 0; x < resolution.x; ++x)
            for (int c = 0; c < NChannels(); ++c) { float v = GetChannel({x, y}, c); fwrite(&v, 4, 1, f); }
    fclose(f);
    return true;
}

'''


def write_if_changed(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path) and open(path).read() == text:
        return
    open(path, "w").write(text)


def make_patched():
    pd = os.path.join(OUT, "patched", "pbrt")
    write_if_changed(os.path.join(pd, "util", "gui.cpp"), GUI_STUB)
    s = open(os.path.join(REF, "src/pbrt/util/image.cpp")).read()
    s = re.sub(r'#include <Imf[A-Za-z]*\.h>\n', '', s)
    a = s.index('// OpenEXR\n#ifndef PBRT_IS_GPU_CODE')
    b = s.index('/' * 75 + '\n// PNG Function Definitions')
    s = s[:a] + EXR_STUB + s[b:]
    tail = '#endif\n\n}  // namespace pbrt'
    i = s.rindex(tail)
    s = s[:i] + s[i + len('#endif\n'):]
    write_if_changed(os.path.join(pd, "util", "image.cpp"), s)
    s = open(os.path.join(REF, "src/pbrt/pbrt.cpp")).read()
    s = s.replace('#include <ImfThreading.h>\n', '')
    s = s.replace('    Imf::setGlobalThreadCount(opt.nThreads ? opt.nThreads : AvailableCores());\n', '')
    write_if_changed(os.path.join(pd, "pbrt.cpp"), s)
    # Ptex version header (ptex/src/ptex/PtexVersion.h.in configured by its CMake; v2.4)
    s = open(os.path.join(REF, "src/ext/ptex/src/ptex/PtexVersion.h.in")).read()
    s = s.replace("@PTEX_MAJOR_VERSION@", "2").replace("@PTEX_MINOR_VERSION@", "4")
    write_if_changed(os.path.join(OUT, "gen", "PtexVersion.h"), s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=os.cpu_count())
    ap.add_argument("--only-ninja", action="store_true")
    ap.add_argument("targets", nargs="*")
    args = ap.parse_args()
    if not os.path.isdir(REF):
        print(f"reference not present at {REF}; using prebuilt oracle/_ref if any", file=sys.stderr)
        return 0 if os.path.exists(os.path.join(OUT, "bin", "pbrt")) else 1
    os.makedirs(os.path.join(OUT, "gen"), exist_ok=True)
    make_patched()

    src = os.path.join(REF, "src")
    ext = os.path.join(src, "ext")
    inc = [os.path.join(OUT, "gen"), src, ext, f"{ext}/stb", f"{ext}/qoi", f"{ext}/libdeflate",
           f"{ext}/filesystem", f"{ext}/ptex/src/ptex", f"{ext}/double-conversion",
           f"{ext}/openvdb/nanovdb", f"{ext}/glfw/include", f"{ext}/glad/include", f"{ext}/utf8proc",
           f"{ext}/flip"]
    cflags = "-O3 -march=x86-64-v3 -ffp-contract=off -fPIC -w " + " ".join(f"'-D{d}'" for d in DEFS) + \
        " " + " ".join(f"-I{i}" for i in inc)
    n = []
    n.append(f"cxxflags = -std=gnu++17 {cflags}")
    n.append(f"cflags = {cflags}")
    n.append("rule cxx\n  command = g++ $cxxflags -MMD -MF $out.d -c $in -o $out\n  depfile = $out.d\n  deps = gcc\n  description = CXX $out")
    n.append("rule cc\n  command = gcc $cflags -MMD -MF $out.d -c $in -o $out\n  depfile = $out.d\n  deps = gcc\n  description = CC $out")
    n.append("rule ar\n  command = rm -f $out && ar rcs $out $in\n  description = AR $out")
    n.append("rule link\n  command = g++ -s -o $out $in $libs -lz -lpthread\n  description = LINK $out")
    n.append("rule soac\n  command = $soac $in > $out\n  description = SOAC $out")
    n.append("rule rgb2spec\n  command = $tool 64 $out $space\n  description = RGB2SPEC $out")

    o = lambda *p: os.path.join(OUT, "obj", *p)
    # generators
    n.append(f"build {o('soac.o')}: cxx {src}/pbrt/cmd/soac.cpp")
    n.append(f"build {OUT}/bin/soac: link {o('soac.o')}")
    n.append(f"build {o('rgb2spec_opt.o')}: cxx {src}/pbrt/cmd/rgb2spec_opt.cpp")
    n.append(f"build {OUT}/bin/rgb2spec_opt: link {o('rgb2spec_opt.o')}")
    gen_h = []
    for soa_in, soa_out in (("pbrt/pbrt.soa", "pbrt_soa.h"),
                            ("pbrt/wavefront/workitems.soa", "wavefront_workitems_soa.h")):
        n.append(f"build {OUT}/gen/{soa_out}: soac {src}/{soa_in} | {OUT}/bin/soac\n  soac = {OUT}/bin/soac")
        gen_h.append(f"{OUT}/gen/{soa_out}")
    objs = []
    for space, name in (("ACES2065_1", "aces"), ("DCI_P3", "dci_p3"), ("REC2020", "rec2020"), ("sRGB", "srgb")):
        cpp = f"{OUT}/gen/rgbspectrum_{name}.cpp"
        n.append(f"build {cpp}: rgb2spec | {OUT}/bin/rgb2spec_opt\n  tool = {OUT}/bin/rgb2spec_opt\n  space = {space}")
        n.append(f"build {o('gen', name + '.o')}: cxx {cpp}")
        objs.append(o('gen', name + '.o'))
    order = " || " + " ".join(gen_h)
    for s in PBRT_SRC:
        path = (os.path.join(OUT, "patched", "pbrt", s + ".cpp") if s in PATCHED
                else f"{src}/pbrt/{s}.cpp")
        n.append(f"build {o('pbrt', s + '.o')}: cxx {path}{order}")
        objs.append(o('pbrt', s + '.o'))
    for s in ("lodepng/lodepng.cpp", "rply/rply.cpp", "gtest/gtest-all.cc"):
        n.append(f"build {o('ext', s + '.o')}: cxx {ext}/{s}")
        objs.append(o('ext', s + '.o'))
    n.append(f"build {OUT}/lib/libpbrt.a: ar {' '.join(objs)}")
    # ext libs
    extobjs = []
    for f in sorted(os.listdir(f"{ext}/ptex/src/ptex")):
        if f.endswith(".cpp") and f != "PtexHalfTableGen.cpp":
            n.append(f"build {o('ptex', f + '.o')}: cxx {ext}/ptex/src/ptex/{f}")
            extobjs.append(o('ptex', f + '.o'))
    for f in sorted(os.listdir(f"{ext}/double-conversion/double-conversion")):
        if f.endswith(".cc"):
            n.append(f"build {o('dc', f + '.o')}: cxx {ext}/double-conversion/double-conversion/{f}")
            extobjs.append(o('dc', f + '.o'))
    n.append(f"build {o('utf8proc.o')}: cc {ext}/utf8proc/utf8proc.c")
    extobjs.append(o('utf8proc.o'))
    for f in ["lib/adler32.c", "lib/crc32.c", "lib/deflate_decompress.c", "lib/deflate_compress.c",
              "lib/gzip_decompress.c", "lib/gzip_compress.c", "lib/zlib_decompress.c",
              "lib/zlib_compress.c", "lib/utils.c", "lib/x86/cpu_features.c"]:
        n.append(f"build {o('deflate', f + '.o')}: cc {ext}/libdeflate/{f}")
        extobjs.append(o('deflate', f + '.o'))
    n.append(f"build {OUT}/lib/libpbrtext.a: ar {' '.join(extobjs)}")
    # one shared library next to the archives: the executables (pbrt, pbrt_b200) then weigh a few hundred kB instead of 48 MB each
    # (the RGB->spectrum tables dominate), which is what travels to the GPU box
    n.append("rule solink\n  command = g++ -s -shared -o $out -Wl,--whole-archive $in -Wl,--no-whole-archive -lz -lpthread\n  description = SOLINK $out")
    n.append(f"build {OUT}/lib/libpbrt_ref.so: solink {OUT}/lib/libpbrt.a {OUT}/lib/libpbrtext.a")
    alibs = f"{OUT}/lib/libpbrt.a {OUT}/lib/libpbrtext.a"
    libs = f"-L{OUT}/lib -lpbrt_ref -Wl,-rpath,'$$ORIGIN/../lib'"
    solib = f"{OUT}/lib/libpbrt_ref.so"
    # executables
    n.append(f"build {o('cmd_pbrt.o')}: cxx {src}/pbrt/cmd/pbrt.cpp{order}")
    n.append(f"build {OUT}/bin/pbrt: link {o('cmd_pbrt.o')} | {solib}\n  libs = {libs}")
    n.append(f"build {o('ref_dump.o')}: cxx {HERE}/ref_dump.cpp{order}")
    n.append(f"build {OUT}/bin/ref_dump: link {o('ref_dump.o')} | {alibs}\n  libs = {alibs}")
    # the reference-side binding of INTEGRATION.md, compiled against the reference (oracle/b200_binding.cpp): reference parser ->
    # BasicScene -> C ABI of include/b200pt.h (dlopen) -> image
    repo = os.path.dirname(HERE)
    n.append("rule cxxbind\n  command = g++ $cxxflags '-DBPT_REPO_ROOT=\"" + repo + "\"' -MMD -MF $out.d -c $in -o $out\n  depfile = $out.d\n  deps = gcc\n  description = CXX $out")
    n.append("rule linkdl\n  command = g++ -s -o $out $in $libs -lz -lpthread -ldl\n  description = LINK $out")
    n.append(f"build {o('b200_binding.o')}: cxxbind {HERE}/b200_binding.cpp{order}")
    n.append(f"build {OUT}/bin/pbrt_b200: linkdl {o('b200_binding.o')} | {solib}\n  libs = {libs}")
    n.append(f"default {OUT}/bin/pbrt" + (f" {OUT}/bin/ref_dump" if os.path.exists(f"{HERE}/ref_dump.cpp") else "") +
             (f" {OUT}/bin/pbrt_b200" if os.path.exists(f"{HERE}/b200_binding.cpp") else ""))
    write_if_changed(os.path.join(OUT, "build.ninja"), "\n".join(n) + "\n")
    if args.only_ninja:
        return 0
    cmd = ["ninja", "-C", OUT, f"-j{args.j}"] + [os.path.join(OUT, t) for t in args.targets]
    return subprocess.call(cmd)


if __name__ == "__main__":
    sys.exit(main())
