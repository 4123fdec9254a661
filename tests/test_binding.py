"""The reference-side binding (oracle/b200_binding.cpp -> oracle/_ref/bin/pbrt_b200: the reference's own parser and BasicScene
driving the engine through the C ABI) against this repo's .pbrt front end (bpt_load_pbrt).

CPU: both routes are run against the ABI recorder (tests/tools/abi_record.cpp: the scene half of include/b200pt.h, dumping every
call with all transforms resolved to rendering space by the engine's own csrc/scene_xforms.h) and the two call streams must be
byte-identical -- every matrix and tracked inverse, every material / light / medium parameter, geometry and image hashes.
GPU: the same scene rendered through both routes must give bit-identical images."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINDING = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt_b200")

pytestmark = pytest.mark.skipif(not os.path.exists(BINDING), reason="oracle/_ref/bin/pbrt_b200 not built (python oracle/build_ref.py)")


@pytest.fixture(scope="module")
def recorder(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("abi_record") / "libabi_record.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-mfma", "-fPIC", "-shared", "-I/usr/local/cuda/include",
                           os.path.join(ROOT, "tests", "tools", "abi_record.cpp"), os.path.join(ROOT, "pbrt-v4_b200", "csrc", "pbrt_frontend.cpp"), "-o", so, "-lz"])
    lib = C.CDLL(so)
    lib.bpt_create.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.bpt_load_pbrt_ex.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    lib.bpt_frontend_error.restype = C.c_char_p
    return so, lib


def _scenes():
    from pbrt_v4_b200 import scenes
    cases = {
        "cornell": lambda d: scenes.cornell(d, res=(64, 64), spp=4),
        "heightfield": lambda d: scenes.diffuse_heightfield(d, n=64, res=(64, 64), spp=4),
        "mixed_envmap": lambda d: scenes.mixed_envmap(d, n=48, res=(64, 64), spp=4, envres=64, nmesh=4),
        "instanced_coated": lambda d: scenes.instanced(d, n_inst=50, subdiv=2, res=(64, 64), spp=4, material="coatedconductor"),
        "smoke_grid": lambda d: scenes.smoke(d, grid=16, res=(64, 64), spp=4),
        "smoke_homogeneous": lambda d: scenes.smoke(d, res=(64, 64), spp=4, homogeneous=True),
        "alpha_tested_panels": lambda d: scenes.cornell_alpha(d, res=(32, 32), spp=1),
        "gbuffer_world": lambda d: scenes.as_gbuffer(scenes.cornell(d, res=(64, 64), spp=4), "world"),
        "smoke_emissive_homogeneous": lambda d: scenes.smoke(d, res=(32, 32), spp=1, homogeneous=True, emission="le"),
        "smoke_emissive_grid": lambda d: scenes.smoke(d, grid=8, res=(32, 32), spp=1, emission="grid"),
        "smoke_rgbgrid": lambda d: scenes.smoke_rgb(d, grid=6, res=(32, 32), spp=1),
    }
    def filtered(d, flt):
        p = scenes.cornell(d, res=(32, 32), spp=1)
        txt = open(p).read().replace("WorldBegin", "PixelFilter " + flt + "\nWorldBegin")
        open(p, "w").write(txt)
        return p
    for k, flt in enumerate(['"box"', '"triangle" "float xradius" 1.5', '"mitchell" "float B" 0.2', '"sinc" "float tau" 2']):
        cases["filter_%d" % k] = (lambda d, flt=flt: filtered(d, flt))
    def with_material(d, mat):
        p = scenes.cornell(d, res=(32, 32), spp=1)
        txt = open(p).read().replace('Material "diffuse" "rgb reflectance" [0.73 0.73 0.73]', "Material " + mat)
        open(p, "w").write(txt)
        return p
    for k, mat in enumerate(['"coateddiffuse" "rgb reflectance" [0.2 0.5 0.7] "float roughness" 0.1 "float thickness" 0.02',
                             '"coateddiffuse" "float g" 0.3 "rgb albedo" [0.4 0.4 0.2] "integer maxdepth" 6 "integer nsamples" 2',
                             '"thindielectric" "float eta" 1.3', '"diffusetransmission" "rgb reflectance" [0.3 0.2 0.1] "rgb transmittance" [0.5 0.6 0.7] "float scale" 1.2']):
        cases["material_%d" % k] = (lambda d, mat=mat: with_material(d, mat))
    def with_sampler(d, smp):
        p = scenes.cornell(d, res=(32, 32), spp=4)
        txt = open(p).read()
        a = txt.index("Sampler "); b = txt.index("\n", a)
        open(p, "w").write(txt[:a] + "Sampler " + smp + txt[b:])
        return p
    for k, smp in enumerate(['"paddedsobol" "integer pixelsamples" 8', '"paddedsobol" "integer pixelsamples" 4 "string randomization" "owen"',
                             '"stratified" "integer xsamples" 3 "integer ysamples" 2', '"stratified" "bool jitter" false', '"independent" "integer pixelsamples" 5',
                             '"sobol" "integer pixelsamples" 8 "string randomization" "owen"', '"halton" "integer pixelsamples" 8', '"halton" "string randomization" "owen" "integer seed" 2', '"pmj02bn" "integer pixelsamples" 4']):
        cases["sampler_%d" % k] = (lambda d, smp=smp: with_sampler(d, smp))
    def with_camera(d, cam):
        p = scenes.cornell(d, res=(32, 32), spp=1)
        txt = open(p).read()
        a = txt.index("Camera "); b = txt.index("\n", a)
        open(p, "w").write(txt[:a] + "Camera " + cam + txt[b:])
        return p
    for k, cam in enumerate(['"orthographic" "float screenwindow" [-1.1 1.1 -1.1 1.1]', '"orthographic" "float lensradius" 0.05 "float focaldistance" 3.2 "float frameaspectratio" 1.2']):
        cases["camera_%d" % k] = (lambda d, cam=cam: with_camera(d, cam))
    for ts in sorted(scenes.TEXTURE_SETS):    # image textures: PNG / PFM readers, filter / wrap / scale / invert / uv mapping, material parameters
        cases["textures_" + ts] = (lambda d, ts=ts: scenes.cornell_textured(d, ts, res=(32, 32), spp=1))
    for ls in sorted(scenes.LIGHT_SETS):   # point / spot / distant lights: parameters, defaults, transforms, power / illuminance
        cases["lights_" + ls] = (lambda d, ls=ls: scenes.cornell_lights(d, ls, area=(ls == "all"), res=(32, 32), spp=1))
    for qs in sorted(scenes.QUADRIC_SETS):   # Shape "sphere" / "disk" / "cylinder": parameters, defaults, transforms, area lights, instances, media
        cases["quadrics_" + qs] = (lambda d, qs=qs: scenes.cornell_quadrics(d, qs, res=(32, 32), spp=1))
    for bs in sorted(scenes.BILINEAR_SETS):   # Shape "bilinearmesh" and the quad faces of PLY files
        cases["bilinear_" + bs] = (lambda d, bs=bs: scenes.cornell_bilinear(d, bs, res=(32, 32), spp=1))
    for parts in ("t", "s", "r", "ts", "tr", "rs", "trs"):   # every combination of translate / rotate / scale in the instance transform
        cases["instances_" + parts] = (lambda d, parts=parts: scenes.instanced(d, n_inst=40, subdiv=1, res=(32, 32), spp=1, material="diffuse", parts=parts))
    return cases


@pytest.mark.parametrize("name", sorted(_scenes()))
def test_front_end_and_reference_parser_binding_issue_identical_abi_calls(recorder, tmp_path, name, monkeypatch):
    so, lib = recorder
    path = _scenes()[name](str(tmp_path))
    fe, bi = str(tmp_path / "frontend.txt"), str(tmp_path / "binding.txt")
    monkeypatch.setenv("BPT_RECORD_OUT", fe)
    h = C.c_void_p()
    assert lib.bpt_create(0, None, C.byref(h)) == 0
    assert lib.bpt_load_pbrt_ex(h, path.encode(), 0, 0) == 0, lib.bpt_frontend_error()
    r = subprocess.run([BINDING, "--quiet", "--seed", "0", path], env=dict(os.environ, BPT_LIB=so, BPT_RECORD_OUT=bi), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    a, b = open(fe).read().splitlines(), open(bi).read().splitlines()
    assert len(a) == len(b) and len(a) >= 6
    for x, y in zip(a, b):
        if x != y:
            xt, yt = x.split(), y.split()
            first = next(((u, v) for u, v in zip(xt, yt) if u != v), (x[-200:], y[-200:]))
            raise AssertionError(f"{name}: record '{xt[0]}' differs: front end {first[0][:200]} vs binding {first[1][:200]}")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cornell", "mixed_envmap", "instanced_coated", "smoke_grid", "gbuffer_world", "lights_all", "textures_bump", "quadrics_instanced", "quadrics_lights", "bilinear_ply"])
def test_binding_and_front_end_render_identical_images(tmp_path, name):
    from pbrt_v4_b200 import engine as E, imageio
    path = _scenes()[name](str(tmp_path))
    gb = name.startswith("gbuffer")
    out = str(tmp_path / ("bind.exr" if gb else "bind.pfm"))
    r = subprocess.run([BINDING, "--quiet", "--seed", "0", "--outfile", out, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, e.sampler_spp())
    if gb:
        names, ref = imageio.read_raw_channels(out)
        mine = e.gbuffer()
        assert names == E.Engine.GBUFFER_CHANNELS
    else:
        ref, mine = imageio.read_pfm(out), e.image()
    e.close()
    assert (np.ascontiguousarray(mine, np.float32).view(np.uint32) == np.ascontiguousarray(ref, np.float32).view(np.uint32)).all()
