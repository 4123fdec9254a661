"""GPU parity tests (B200 box): the CUDA path, called through the C ABI, against
  * the golden vectors / images produced by the reference itself (tests/golden/),
  * the plain-C oracle on seeded inputs (oracle/oracle.c),
  * the reference CPU binary run live when oracle/_ref/bin/pbrt travelled to the box,
  * size-independent properties at the full BASELINE size (determinism, tile and sample-range additivity).
Integer / index results must be bit-exact; images must satisfy per-channel RMSE < 1e-4 (north_star tolerance)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_py as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EXE = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
RMSE_TOL = 1e-4


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def E():
    from pbrt_v4_b200 import engine
    return engine


@pytest.fixture(scope="module")
def cornell_engine(E, tmp_path_factory):
    from pbrt_v4_b200 import scenes
    d = tmp_path_factory.mktemp("cornell")
    path = scenes.cornell(str(d), res=(64, 64), spp=4)
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    yield e, path
    e.close()


def test_native_library_is_the_one_running(E):
    lib = E.load_library()
    assert os.path.samefile(lib._name, os.path.join(ROOT, "pbrt-v4_b200", "libb200pt.so"))
    maps = open("/proc/self/maps").read()
    assert "libb200pt.so" in maps


def test_sampler_on_device_matches_reference_golden(E, tmp_path):
    from pbrt_v4_b200 import scenes
    G = O.golden()
    for cfg in (0, 2):   # fastowen zsobol at two resolutions / spp (the configurations the scenes use)
        spp, rx, ry, seed, rs = [int(v) for v in G[f"sampler{cfg}_params"]]
        p = scenes.cornell(str(tmp_path / f"s{cfg}"), res=(rx, ry), spp=spp)
        e = E.Engine(0)
        e.load_pbrt(p, seed=seed)
        inp = G[f"sampler{cfg}_in"].reshape(-1, 4)
        o1, o2 = e.probe_sampler(inp[:, 0:2], inp[:, 2], inp[:, 3])
        z = G[f"sampler{cfg}_zsobol"].reshape(-1, 3)
        assert (bits(o1) == bits(z[:, 0])).all()
        assert (bits(o2) == bits(z[:, 1:3])).all()
        e.close()


def test_watertight_triangle_test_on_device_matches_reference_golden(cornell_engine):
    e, _ = cornell_engine
    G = O.golden()
    i = G["tri_in"].reshape(-1, 16)
    o = G["tri_out"].reshape(-1, 5)
    out = e.probe_intersect_triangle(i[:, 0:7], i[:, 7:16])
    assert (out[:, 0] >= 0).all(), "hoisted traversal variant disagrees with the restated reference test"
    assert (out[:, 0] == o[:, 0]).all()
    m = o[:, 0] == 1
    assert (bits(out[m, 1:5]) == bits(o[m, 1:5])).all()


def _soup_scene(tmp_path, verts):
    P = verts.reshape(-1, 3)
    idx = np.arange(len(P))
    s = ("LookAt 0 0 0  0 0 -1  0 1 0\nCamera \"perspective\" \"float fov\" 40\nFilm \"rgb\" \"integer xresolution\" 16 \"integer yresolution\" 16\n"
         "WorldBegin\nShape \"trianglemesh\" \"integer indices\" [%s] \"point3 P\" [%s]\n" %
         (" ".join(map(str, idx)), " ".join(repr(float(v)) for v in P.reshape(-1))))
    p = tmp_path / "soup.pbrt"
    p.write_text(s)
    return str(p)


def test_bvh_closest_and_any_hits_match_reference_bvh_and_oracle(E, tmp_path):
    """2000-triangle soup, 20000 rays: (primitive, b0, b1, b2, t) must equal what the reference's BVHAggregate::Intersect
    returned (golden) and what the brute-force C oracle returns; IntersectP likewise."""
    G = O.golden()
    verts = G["bvh_verts"].reshape(-1, 9)
    rays = G["bvh_rays"].reshape(-1, 8)
    hits = G["bvh_hits"].reshape(-1, 12)
    e = E.Engine(0)
    e.load_pbrt(_soup_scene(tmp_path, verts), seed=0)   # camera at the origin looking down -z: render space == world space
    prim, bary, t = e.trace_closest(rays)
    occ = e.trace_any(rays)
    gp = hits[:, 0].astype(np.int32)
    assert (prim == gp).all()
    m = gp >= 0
    assert (bits(t[m]) == bits(hits[m, 4])).all() and (bits(bary[m]) == bits(hits[m, 1:4])).all()
    assert (occ == hits[:, 5].astype(np.int32)).all()
    # oracle on a fresh seeded ray set (edge case: rays starting on / grazing triangles)
    rng = np.random.default_rng(3)
    r2 = np.zeros((4000, 8), np.float32)
    tri = verts[rng.integers(0, len(verts), 4000)]
    r2[:, 0:3] = rng.uniform(-1.5, 1.5, (4000, 3))
    target = tri[:, 0:3] * 0.5 + tri[:, 3:6] * 0.5          # aim at edge midpoints
    d = target - r2[:, 0:3]
    r2[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r2[:, 3] = np.inf
    p1, b1, t1 = e.trace_closest(r2)
    p2, b2, t2 = O.closest_hits(verts, r2)
    assert (p1 == p2).all() and (bits(t1) == bits(t2)).all() and (bits(b1[p2 >= 0]) == bits(b2[p2 >= 0])).all()
    e.close()


def test_empty_and_degenerate_inputs(E, tmp_path):
    # zero rays; a degenerate (zero-area) triangle must never be hit (reference: IntersectTriangle returns {} first thing)
    verts = np.array([[0, 0, -2, 1, 0, -2, 0.5, 0, -2],        # collinear -> degenerate
                      [-1, -1, -3, 1, -1, -3, 0, 1, -3]], np.float32)
    e = E.Engine(0)
    e.load_pbrt(_soup_scene(tmp_path, verts), seed=0)
    prim, bary, t = e.trace_closest(np.zeros((0, 8), np.float32))
    assert len(prim) == 0
    rays = np.array([[0.25, 0, 0, np.inf, 0, 0, -1, 0], [0, 0, 0, np.inf, 0, 0, -1, 0], [5, 5, 0, np.inf, 0, 0, -1, 0]], np.float32)
    prim, bary, t = e.trace_closest(rays)
    assert list(prim) == [1, 1, -1] and np.isinf(t[2])
    e.close()


@pytest.mark.parametrize("name,gen,kw,spp", [("cornell_64x64_4spp", "cornell", dict(res=(64, 64)), 4),
                                             ("cornell_128x128_16spp", "cornell", dict(res=(128, 128)), 16),
                                             ("heightfield64_96x96_8spp", "diffuse_heightfield", dict(n=64, res=(96, 96)), 8)])
def test_image_matches_reference_golden(E, tmp_path, name, gen, kw, spp):
    from pbrt_v4_b200 import scenes, imageio
    path = getattr(scenes, gen)(str(tmp_path), spp=spp, **kw)
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, spp)
    img = e.image()
    ref = imageio.read_pfm(os.path.join(ROOT, "tests", "golden", name + ".pfm"))
    rm = imageio.channel_rmse(img, ref)
    assert max(rm) < RMSE_TOL, rm
    assert int((bits(img) != bits(ref)).any(axis=-1).sum()) <= 1, "diffuse scenes are sample-exact up to exact-t ties"
    e.close()


def test_render_pbrt_one_call_equivalent_of_RenderWavefront(E, tmp_path):
    """engine.render_pbrt(path): load, render the file's own pixelsamples, resolve -- and the --spp override."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell(str(tmp_path), res=(64, 64), spp=4)
    img, st = E.render_pbrt(path)
    ref = imageio.read_pfm(os.path.join(ROOT, "tests", "golden", "cornell_64x64_4spp.pfm"))
    assert (bits(img) == bits(ref)).all() and st["camera_rays"] == 64 * 64 * 4
    img2, st2 = E.render_pbrt(path, spp=2)
    assert st2["camera_rays"] == 64 * 64 * 2


def _render(E, path, spp):
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, spp)
    img = e.image()
    e.close()
    return img


def _compare(img, ref):
    """(number of pixels that are not bit-identical, per-channel RMSE)"""
    from pbrt_v4_b200 import imageio
    ndiff = int((bits(img) != bits(ref)).any(axis=-1).sum())
    return ndiff, np.array(imageio.channel_rmse(img, ref))


# Golden images of the configuration classes that were only statistically comparable in round 1.  Since round 2 the device path is
# sample-exact for all of them (exact libm, Interaction::wo normalisation, the reference's RenderFromObject() inverse and its
# FMA-chain matrix product): the only admissible differences are exact-t ties on shared triangle edges (scripts/tie_audit.py), so the
# tests pin the NUMBER of differing pixels (measured: see profiles/r02*_parity_table.json) on top of the north-star RMSE bound.
@pytest.mark.parametrize("name,gen,kw,spp,max_diff_pixels", [
    ("mixed48_96x96_16spp", "mixed_envmap", dict(n=48, res=(96, 96), envres=64, nmesh=4), 16, 2),
    ("instanced50_96x96_16spp", "instanced", dict(n_inst=50, subdiv=2, res=(96, 96), material="diffuse"), 16, 2),
    ("coated50_64x64_64spp", "instanced", dict(n_inst=50, subdiv=2, res=(64, 64), material="coatedconductor"), 64, 2),
    ("smoke16_64x64_64spp", "smoke", dict(grid=16, res=(64, 64)), 64, 0),
    ("smokehomog_64x64_64spp", "smoke", dict(res=(64, 64), homogeneous=True), 64, 0)])
def test_all_configuration_classes_match_reference_golden_sample_exactly(E, tmp_path, name, gen, kw, spp, max_diff_pixels):
    """Config 3 ingredients (smooth dielectric, rough conductors, image infinite light), config 4 (two-level BVH, LayeredBxDF coated
    conductor), config 5 (grid / homogeneous media behind interface surfaces): per-channel RMSE < 1e-4 (relative to the image mean
    for the HDR-lit scene) AND at most `max_diff_pixels` pixels not bit-identical to the reference CPU VolPathIntegrator."""
    from pbrt_v4_b200 import scenes, imageio
    path = getattr(scenes, gen)(str(tmp_path), spp=spp, **kw)
    img = _render(E, path, spp)
    ref = imageio.read_pfm(os.path.join(ROOT, "tests", "golden", name + ".pfm"))
    ndiff, rm = _compare(img, ref)
    assert ndiff <= max_diff_pixels, (ndiff, rm)
    if ndiff == 0:
        assert (rm == 0).all()


LIVE_SCENES = {
    # name: (generator, kwargs, spp, max differing pixels) -- every BASELINE configuration class at a size the reference renders in seconds
    "config2_heightfield1m_256_16spp": ("diffuse_heightfield", dict(n=708, res=(256, 256)), 16, 8),
    "config3_mixed257k_256_16spp": ("mixed_envmap", dict(n=360, res=(256, 256), envres=512, nmesh=16), 16, 4),
    "config4_instanced1k_coated_256_16spp": ("instanced", dict(n_inst=1000, subdiv=2, res=(256, 256), material="coatedconductor"), 16, 4),
    "config4_instanced1k_diffuse_256_16spp": ("instanced", dict(n_inst=1000, subdiv=2, res=(256, 256), material="diffuse"), 16, 4),
    "config5_smoke64_128_64spp": ("smoke", dict(grid=64, res=(128, 128)), 64, 0),
    # a6: alpha-tested primitives (constant alpha), top-level and instanced -- closest-hit and shadow rays take the same test
    "alpha_tested_panels_96_16spp": ("cornell_alpha", dict(res=(96, 96)), 16, 0),
    # emissive media (cpu/integrators.cpp:989-1000; media.h:223-246, 296-318): "Le" on a homogeneous medium, on a grid medium, and
    # with a per-voxel "Lescale" array
    "emissive_homogeneous_64_32spp": ("smoke", dict(res=(64, 64), homogeneous=True, emission="le"), 32, 0),
    "emissive_grid16_64_32spp": ("smoke", dict(grid=16, res=(64, 64), emission="le"), 32, 0),
    "emissive_gridscale16_64_32spp": ("smoke", dict(grid=16, res=(64, 64), emission="grid"), 32, 0),
}


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("name", sorted(LIVE_SCENES))
def test_live_reference_parity_on_every_baseline_configuration_class(E, tmp_path, name):
    """The unmodified reference (oracle/_ref/bin/pbrt, CPU VolPathIntegrator, same seed) rendered live on this box against the
    engine: config-2 geometry at full triangle count, a 257 k-triangle dielectric / conductor / HDR-envmap scene, 1000 instances
    (coated conductor and diffuse), a 64^3 smoke.  Differences are admissible only as exact-t ties (a handful of pixels, pinned)."""
    from pbrt_v4_b200 import scenes, imageio
    gen, kw, spp, max_diff = LIVE_SCENES[name]
    path = getattr(scenes, gen)(str(tmp_path), spp=spp, **kw)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, spp)
    ndiff, rm = _compare(img, ref)
    assert ndiff <= max_diff, (name, ndiff, rm)
    # all the other pixels are bit-identical, so the RMSE is that of the tie pixels alone; without ties it is exactly zero
    if ndiff == 0:
        assert (rm == 0).all()


def test_ray_sort_changes_nothing(E, tmp_path, monkeypatch):
    """a18: with BPT_RAY_SORT=1 (radix sort of the extension-ray queue by direction / origin key) the film is bit-identical."""
    from pbrt_v4_b200 import scenes
    path = scenes.diffuse_heightfield(str(tmp_path), n=128, res=(128, 128), spp=8)
    base = _film(E, path, 8)
    monkeypatch.setenv("BPT_RAY_SORT", "1")
    assert (bits(_film(E, path, 8)) == bits(base)).all()


def _film(E, path, spp):
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, spp)
    f = e.film().copy()
    e.close()
    return f


def test_raw_any_hit_and_closest_hit_on_an_instanced_scene_match_the_oracle(E, tmp_path):
    """WavefrontAggregate::IntersectShadow / IntersectClosest on raw rays through a two-level BVH: occlusion flags and closest hits
    against the brute-force C oracle run on the instanced triangles transformed to render space on the host (double precision), with
    a tolerance on t because the reference tests instanced triangles in instance space (hit / miss and primitive must agree except
    within 1e-5 of a silhouette)."""
    from pbrt_v4_b200 import scenes
    path = scenes.instanced(str(tmp_path), n_inst=40, subdiv=1, res=(32, 32), spp=1, material="diffuse")
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    rng = np.random.default_rng(11)
    n = 20000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-4, 4, (n, 3)) + np.array([0, 3, 0])
    d = rng.normal(size=(n, 3)); d[:, 1] = -np.abs(d[:, 1])
    rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = np.inf
    prim, bary, t = e.trace_closest(rays)
    occ = e.trace_any(rays)
    e.close()
    assert ((prim >= 0) == (occ != 0)).all(), "any-hit disagrees with closest-hit about occlusion"
    assert (prim >= 0).mean() > 0.3 and (prim < 0).mean() > 0.01
    assert np.isfinite(t[prim >= 0]).all() and (t[prim >= 0] > 0).all()


@pytest.mark.parametrize("name,gen,kw,csys", [("gbuffer_cornell_64x64_4spp_camera", "cornell", dict(res=(64, 64)), "camera"),
                                              ("gbuffer_heightfield64_64x64_4spp_world", "diffuse_heightfield", dict(n=64, res=(64, 64)), "world")])
def test_gbuffer_film_matches_reference_golden(E, tmp_path, name, gen, kw, csys):
    """GBufferFilm (film.cpp:588-790): all 25 channels against the reference's own output.  Geometry channels (P, dzdx/dzdy from
    the camera ray differentials, N, Ns, u, v) must be bit-identical; colour, albedo and variance within 1e-4."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.as_gbuffer(getattr(scenes, gen)(str(tmp_path), spp=4, **kw), csys)
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, 4)
    g = e.gbuffer()
    e.close()
    names, ref = imageio.read_raw_channels(os.path.join(ROOT, "tests", "golden", name + ".raw"))
    assert names == E.Engine.GBUFFER_CHANNELS
    for c, nm in enumerate(names):
        if nm.split(".")[0] in ("P", "dzdx", "dzdy", "N", "Ns", "u", "v"):
            assert (bits(g[..., c]) == bits(ref[..., c])).all(), nm
        else:
            assert np.sqrt(((g[..., c] - ref[..., c]) ** 2).mean()) < RMSE_TOL, nm


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("flt", ['"box"', '"triangle"', '"mitchell"', '"sinc"', '"gaussian" "float sigma" 0.8 "float xradius" 2',
                                 '"mitchell" "float B" 0.2 "float C" 0.4 "float xradius" 1.5 "float yradius" 2.5'])
def test_every_pixel_filter_of_the_dispatch_surface_matches_the_live_reference(E, tmp_path, flt):
    """Filter {Box, Gaussian, Mitchell, LanczosSinc, Triangle} (base/filter.h:25): FilterSampler tables for the sampled ones (signed
    weights for the negative lobes), analytic sampling for box and triangle -- Cornell box, bit-identical to the reference."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell(str(tmp_path), res=(48, 48), spp=8)
    s = open(path).read().replace("WorldBegin", "PixelFilter " + flt + "\nWorldBegin")
    open(path, "w").write(s)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 8)
    assert (bits(img) == bits(ref)).all(), flt


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
def test_gbuffer_film_with_media_and_interface_surfaces_matches_the_live_reference(E, tmp_path):
    """GBufferFilm on the smoke scene: the first surface hit lies behind an "interface" surface (the medium's box), so the visible
    surface is recorded after SkipIntersection respawned the ray -- the ray differentials must still be the camera ray's
    (interaction.h SkipIntersection replaces only the Ray part).  Geometry channels bit-identical, colour within 1e-4."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.as_gbuffer(scenes.smoke(str(tmp_path), grid=16, res=(48, 48), spp=8), "world")
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.exr"), path])
    names, ref = imageio.read_raw_channels(str(tmp_path / "ref.exr"))
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, 8)
    g = e.gbuffer()
    e.close()
    assert names == E.Engine.GBUFFER_CHANNELS
    for c, nm in enumerate(names):
        if nm.split(".")[0] in ("P", "dzdx", "dzdy", "N", "Ns", "u", "v"):
            assert (bits(g[..., c]) == bits(ref[..., c])).all(), nm
        else:
            assert np.sqrt(((g[..., c] - ref[..., c]) ** 2).mean()) < RMSE_TOL, nm


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("smp", ['"paddedsobol" "integer pixelsamples" 8', '"paddedsobol" "integer pixelsamples" 16 "string randomization" "owen"',
                                 '"paddedsobol" "integer pixelsamples" 8 "string randomization" "permutedigits"',
                                 '"stratified" "integer xsamples" 3 "integer ysamples" 2', '"stratified" "bool jitter" false',
                                 '"independent" "integer pixelsamples" 6', '"zsobol" "integer pixelsamples" 8 "string randomization" "owen"',
                                 '"sobol" "integer pixelsamples" 8', '"sobol" "integer pixelsamples" 6 "string randomization" "none"',
                                 '"sobol" "integer pixelsamples" 4 "string randomization" "permutedigits" "integer seed" 3',
                                 '"halton" "integer pixelsamples" 8', '"halton" "integer pixelsamples" 5 "string randomization" "none"',
                                 '"halton" "integer pixelsamples" 4 "string randomization" "owen"',
                                 '"pmj02bn" "integer pixelsamples" 16', '"pmj02bn" "integer pixelsamples" 6 "integer seed" 2'])
def test_samplers_of_the_dispatch_surface_match_the_live_reference(E, tmp_path, smp):
    """Sampler {PMJ02BN, Independent, Stratified, Halton, PaddedSobol, Sobol, ZSobol} (base/sampler.h:40-42; samplers.h:367-625): every sampler is a
    stateless function of (pixel, sample index, dimension), so a whole Cornell render must be bit-identical to the reference's.
    Un-jittered stratified samples sit on an exact lattice, and in the axis-aligned box 0.8 % of those camera rays run exactly through
    edges shared by two walls (same hit point bit for bit, either primitive is a valid closest hit: profiles/r02o_stratified_diag.log);
    that case looks at the box from slightly off axis, so that what is compared is the sampler and not the tie order."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell(str(tmp_path), res=(48, 48), spp=8)
    txt = open(path).read()
    a = txt.index("Sampler "); b = txt.index("\n", a)
    txt = txt[:a] + "Sampler " + smp + txt[b:]
    if '"bool jitter" false' in smp or smp.startswith('"sobol"') or smp.startswith('"halton"'):   # (their pixel samples are dyadic / triadic rationals too)
        a = txt.index("LookAt "); b = txt.index("\n", a)
        f = [float(v) for v in txt[a + 7:b].split()]
        f[0] += 0.0137; f[1] += 0.0071
        txt = txt[:a] + "LookAt " + " ".join(repr(v) for v in f) + txt[b:]
    open(path, "w").write(txt)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, e.sampler_spp())
    img = e.image()
    e.close()
    assert (bits(img) == bits(ref)).all(), smp


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("mat", ['"coateddiffuse" "rgb reflectance" [0.2 0.5 0.7] "float roughness" 0.1',
                                 '"coateddiffuse" "float g" 0.3 "rgb albedo" [0.4 0.4 0.2] "integer maxdepth" 6 "integer nsamples" 2 "float roughness" 0.3',
                                 '"thindielectric" "float eta" 1.3', '"thindielectric" "spectrum eta" "glass-BK7"',
                                 '"diffusetransmission" "rgb reflectance" [0.3 0.2 0.1] "rgb transmittance" [0.5 0.6 0.7] "float scale" 1.2',
                                 '"diffusetransmission"'])
def test_more_materials_of_the_dispatch_surface_match_the_live_reference(E, tmp_path, mat):
    """f4: CoatedDiffuse (LayeredBxDF<Dielectric, Diffuse>), ThinDielectric, DiffuseTransmission (base/material.h:36-39) on the walls
    of the Cornell box, bit-identical to the reference."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell(str(tmp_path), res=(48, 48), spp=16)
    txt = open(path).read().replace('Material "diffuse" "rgb reflectance" [0.73 0.73 0.73]', "Material " + mat)
    open(path, "w").write(txt)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    assert (bits(img) == bits(ref)).all(), (mat, int((bits(img) != bits(ref)).any(axis=-1).sum()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("lights,area", [("point", False), ("spot", False), ("distant", False), ("point", True), ("spot", True),
                                         ("distant", True), ("all", True), ("all", False)])
def test_point_spot_and_distant_lights_match_the_live_reference(E, tmp_path, lights, area):
    """f4: PointLight / SpotLight / DistantLight (lights.h:189-297, 734-789): SampleLi, the delta-light estimator of SampleLd
    (cpu/integrators.cpp:1385-1388), their LightBounds in the light BVH (a lone bounded light is a one-leaf BVH, not a uniform sampler),
    DistantLight in the sampler's list of unbounded lights; with "power" / "illuminance" / blackbody / transforms in the "all" set."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell_lights(str(tmp_path), lights, area=area, res=(48, 48), spp=16)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    assert ref.max() > 0
    assert (bits(img) == bits(ref)).all(), (lights, area, int((bits(img) != bits(ref)).any(axis=-1).sum()), float(np.abs(img - ref).max()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
def test_delta_lights_through_media_match_the_live_reference(E, tmp_path):
    """Delta lights with participating media: shadow rays ratio-track through the grid medium and its interface surfaces with
    r_u = 0 (cpu/integrators.cpp:1331-1388); one point light sits inside the medium."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.smoke(str(tmp_path), grid=16, res=(48, 48), spp=16)
    txt = open(path).read().replace("WorldBegin\n", "WorldBegin\n" + 'LightSource "spot" "point3 from" [1.5 1.8 1.2] "point3 to" [0 0.5 0] "rgb I" [40 40 40]\n'
                                    'LightSource "point" "point3 from" [0.1 0.6 0.05] "rgb I" [0.5 0.4 0.3]\nLightSource "distant" "point3 from" [-1 2 0.5] "rgb L" [0.8 0.8 1]\n', 1)
    open(path, "w").write(txt)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    assert (bits(img) == bits(ref)).all(), (int((bits(img) != bits(ref)).any(axis=-1).sum()), float(np.abs(img - ref).max()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("which", ["albedo", "trilinear", "ewa", "point", "grey", "roughness", "bump", "bump_dielectric", "normalmap", "scale_mix", "checkerboard"])
def test_image_textures_and_bump_mapping_match_the_live_reference(E, tmp_path, which):
    """f1: Texture "imagemap" (float and spectrum; PNG 8-bit sRGB / linear, PFM float; point / bilinear / trilinear / EWA; repeat /
    clamp / black; scale, invert, uv scale and offset) on reflectance, transmittance and roughness parameters, and bump mapping
    through "displacement" -- with the screen-space differentials they are filtered with: camera ray differentials at the first hit,
    SpawnRay's specular differentials behind the mirror and the glass slab, Camera::Approximate_dp_dxy after any diffuse bounce
    (textures.cpp:359-387, util/mipmap.cpp:228-347, interaction.cpp:41-150, materials.h:107-138).  Bit-identical renders."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell_textured(str(tmp_path), which, res=(48, 48), spp=16)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    assert (bits(img) == bits(ref)).all(), (which, int((bits(img) != bits(ref)).any(axis=-1).sum()), float(np.abs(img - ref).max()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("cam", ['"orthographic" "float screenwindow" [-1.1 1.1 -1.1 1.1]',
                                 '"orthographic" "float lensradius" 0.05 "float focaldistance" 3.2 "float frameaspectratio" 1.2'])
def test_orthographic_camera_matches_the_live_reference(E, tmp_path, cam):
    """Camera "orthographic" (cameras.cpp:284-353), with and without a lens: RGB render of the textured, bump-mapped box (the camera's
    own differentials and its fixed "minimum differentials" feed the texture filter) and the GBuffer geometry channels."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell_textured(str(tmp_path), "bump", res=(48, 48), spp=8)
    txt = open(path).read()
    a = txt.index("Camera "); b = txt.index("\n", a)
    open(path, "w").write(txt[:a] + "Camera " + cam + txt[b:])
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 8)
    assert ref.max() > 0 and (bits(img) == bits(ref)).all(), (cam, int((bits(img) != bits(ref)).any(axis=-1).sum()), float(np.abs(img - ref).max()))
    gpath = scenes.as_gbuffer(path, "camera")
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.exr"), gpath])
    names, gref = imageio.read_raw_channels(str(tmp_path / "ref.exr"))
    e = E.Engine(0); e.load_pbrt(gpath, seed=0); e.render(0, 8); g = e.gbuffer(); e.close()
    for c, nm in enumerate(names):
        assert (bits(g[..., c]) == bits(gref[..., c])).all(), nm


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("emission", [False, True])
def test_rgb_grid_medium_matches_the_live_reference(E, tmp_path, emission):
    """Medium "rgbgrid" (RGBGridMedium, media.h:355-427): per-voxel RGB absorption / scattering coefficients (and emission) looked up
    trilinearly as spectra, majorants from the per-voxel spectral maxima; whole render bit-identical without emission, within one ulp of the pixel value with it."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.smoke_rgb(str(tmp_path), grid=12, res=(48, 48), spp=16, emission=emission)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    nd, rmse = _compare(img, ref)
    if emission:
        # r02u: with the "Le" grid 303 of 2304 pixels differ from the reference in the last bit of a channel (max |diff| 2^-15 on values of
        # 256..512), far inside the north-star tolerance; the RGB grid lookups themselves are pinned on the host.  Not yet explained:
        # pinned here as a residual, not as parity.
        assert (rmse < RMSE_TOL).all() and nd <= 400 and float(np.abs(img - ref).max()) <= 2.0 ** -14, (nd, rmse.tolist(), float(np.abs(img - ref).max()))
    else:
        assert nd == 0, (emission, nd, float(np.abs(img - ref).max()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
def test_textured_gbuffer_and_media_match_the_live_reference(E, tmp_path):
    """Textures with the other users of the ray differentials: GBufferFilm (albedo of a textured, bump-mapped surface; Ns after bump
    mapping) and a camera ray that crosses interface surfaces of a medium before it reaches the textured floor (SkipIntersection)."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.as_gbuffer(scenes.cornell_textured(str(tmp_path / "g"), "bump", res=(32, 32), spp=4), "world")
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.exr"), path])
    names, ref = imageio.read_raw_channels(str(tmp_path / "ref.exr"))
    e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, 4); g = e.gbuffer(); e.close()
    for c, nm in enumerate(names):
        assert (bits(g[..., c]) == bits(ref[..., c])).all(), nm
    path = scenes.smoke(str(tmp_path / "s"), grid=8, res=(48, 48), spp=8)
    scenes.texture_images(str(tmp_path / "s"))
    txt = open(path).read().replace("WorldBegin\n", 'WorldBegin\nTexture "alb" "spectrum" "imagemap" "string filename" "albedo.png" "string filter" "trilinear" "float uscale" 6 "float vscale" 6\n', 1)
    old = 'Material "diffuse" "rgb reflectance" [0.6 0.6 0.6]\nShape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [-3 0 -3  -3 0 3  3 0 3  3 0 -3]\n'
    assert old in txt
    txt = txt.replace(old, 'Material "diffuse" "texture reflectance" "alb"\nShape "trianglemesh" "integer indices" [0 1 2 0 2 3] "point3 P" [-3 0 -3  -3 0 3  3 0 3  3 0 -3] "point2 uv" [0 0  0 1  1 1  1 0]\n')
    open(path, "w").write(txt)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 8)
    assert (bits(img) == bits(ref)).all(), (int((bits(img) != bits(ref)).any(axis=-1).sum()), float(np.abs(img - ref).max()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("which", ["spheres", "disk_cylinder", "lights", "instanced", "media", "textured", "mappings"])
def test_spheres_disks_and_cylinders_match_the_live_reference(E, tmp_path, which):
    """f3: Sphere / Disk / Cylinder (shapes.h:107-790) as custom-primitive leaves of the BVH: BasicIntersect with the reference's interval
    arithmetic, InteractionFromIntersection (dpdu / dpdv / dndu / dndv, the doubly normalised wo), clipping by zmin / zmax / phimax,
    rotations, non-uniform and mirroring scales, ReverseOrientation; as DiffuseAreaLights (cone sampling of spheres from outside, area
    sampling from inside and for disks / cylinders, their LightBounds in the light BVH, "power"); inside object instances; as the
    boundary of a medium (interface material: transmittance rays cross them); with image textures and bump maps; "mappings": the
    spherical / cylindrical / planar texture mappings (textures.h:108-215) on a sphere, a cylinder and a patch.  Whole renders,
    bit-identical to the reference."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell_quadrics(str(tmp_path), which, res=(64, 64), spp=16)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    assert ref.max() > 0
    nd, rmse = _compare(img, ref)
    assert nd == 0, (which, nd, rmse.tolist(), float(np.abs(img - ref).max()))


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
@pytest.mark.parametrize("which", ["patches", "ply"])
def test_bilinear_patches_match_the_live_reference(E, tmp_path, which):
    """f3: BilinearPatch (shapes.h:1279-1545, shapes.cpp:1038-1357) as custom-primitive leaves: IntersectBilinearPatch, the interaction
    with (s,t) re-parameterisation and interpolated shading normals (RotateFromTo), rectangular patches as area lights sampled as
    spherical rectangles, twisted patches sampled by area with the bilinear warp, their NormalBounds in the light BVH, patches
    collapsed to triangles, mirroring transforms / ReverseOrientation; the quad faces of a PLY file (triangle mesh + patch mesh over
    the same vertices, util/mesh.cpp:296-306), top-level and inside an object instance.  Whole renders, bit-identical."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.cornell_bilinear(str(tmp_path), which, res=(64, 64), spp=16)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    ref = imageio.read_pfm(str(tmp_path / "ref.pfm"))
    img = _render(E, path, 16)
    assert ref.max() > 0
    nd, rmse = _compare(img, ref)
    assert nd == 0, (which, nd, rmse.tolist(), float(np.abs(img - ref).max()))


def test_raw_rays_against_quadrics_match_the_host_restatement(E, tmp_path):
    """The traversal's quadric leaves against a brute-force loop over the same shapes on the host is covered by the renders above;
    here: closest-hit t of raw rays is finite exactly where any-hit reports occlusion, on a scene with all three quadrics."""
    from pbrt_v4_b200 import scenes
    path = scenes.cornell_quadrics(str(tmp_path), "disk_cylinder", res=(16, 16), spp=1)
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    rng = np.random.default_rng(3)
    rays = np.zeros((8192, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-0.9, 0.9, (8192, 3)) + np.array([0, 1, 0])
    d = rng.normal(size=(8192, 3)); rays[:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    rays[:, 3] = np.inf
    prim, bary, t = e.trace_closest(rays)
    occ = e.trace_any(rays)
    e.close()
    assert ((prim >= 0) == (occ != 0)).all()
    assert (prim >= 12).any(), "no ray hit a quadric"


def test_film_write_exr_carries_the_resolved_channels(E, tmp_path):
    """bpt_film_write_exr: the 25 GBuffer channels (and R G B for Film "rgb") in the file == bpt_film_resolve[_gbuffer]."""
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.as_gbuffer(scenes.cornell(str(tmp_path), res=(64, 64), spp=4), "camera")
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, 4)
    g = e.gbuffer()
    e.write_exr(tmp_path / "g.exr")
    e.close()
    names, back, attrs = imageio.read_exr(tmp_path / "g.exr")
    assert sorted(names) == sorted(E.Engine.GBUFFER_CHANNELS) and "samplesPerPixel" in attrs and "pixelBounds" in attrs
    for k, nm in enumerate(names):
        assert (bits(back[..., k]) == bits(g[..., E.Engine.GBUFFER_CHANNELS.index(nm)])).all(), nm


@pytest.mark.skipif(not os.path.exists(REF_EXE), reason="oracle/_ref/bin/pbrt not on this box")
def test_image_matches_live_reference_on_32k_triangle_heightfield(E, tmp_path):
    from pbrt_v4_b200 import scenes, imageio
    path = scenes.diffuse_heightfield(str(tmp_path), n=128, res=(192, 192), spp=64)
    subprocess.check_call([REF_EXE, "--quiet", "--seed", "0", "--outfile", str(tmp_path / "ref.pfm"), path])
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    e.render(0, 64)
    rm = imageio.channel_rmse(e.image(), imageio.read_pfm(str(tmp_path / "ref.pfm")))
    assert max(rm) < RMSE_TOL, rm
    e.close()


def test_full_size_properties_1m_triangles(E, tmp_path):
    """BASELINE configs[1] geometry (999 698 triangles, 1024x1024) at 4 spp: determinism, tile additivity (rendering by
    bands == rendering at once, bit for bit) and sample-range additivity (film(0..1) + film(2..3) == film(0..3))."""
    from pbrt_v4_b200 import scenes
    from pbrt_v4_b200.parallel import partition_bands
    path = scenes.diffuse_heightfield(str(tmp_path), n=708, res=(1024, 1024), spp=4)
    e = E.Engine(0)
    e.load_pbrt(path, seed=0)
    st = e.stats()
    assert st["bvh_triangles"] == 999698 + 2
    e.render(0, 4)
    full = e.film().copy()
    assert np.isfinite(full).all() and (full[..., 3] > 0).all()
    e.film_clear(); e.render(0, 4)
    assert (e.film() == full).all(), "render is not deterministic"
    e.film_clear()
    for r in range(3):
        for a, b in partition_bands(1024, 3, r, band=16):
            e.render(0, 4, (0, 1024, a, b))
    assert (e.film() == full).all(), "band-partitioned render differs from the full render"
    e.film_clear()
    for r in range(3):
        e.render_bands(0, 4, 16, 3, r)   # the multi-GPU path: all bands of a part in one call
    assert (e.film() == full).all(), "interleaved-band render differs from the full render"
    e.film_clear(); e.render(0, 2); e.render(2, 2)
    assert (e.film() == full).all(), "sample-range split differs from the full render"
    e.close()
