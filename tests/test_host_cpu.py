"""CPU tests of the host-side logic: C-ABI library loads and exports every declared symbol, the engine refuses to run
without a GPU (no fallback), scene generators / PFM io, band partition, and the world_size=2 film reduce over gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(engine_module):
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    declared = set(re.findall(r"\b(bpt_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = engine_module.load_library()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", engine_module.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bpt_[a-z0-9_]+)", out))
    assert declared <= exported


def test_no_cpu_fallback(engine_module):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(engine_module.BptError) as ei:
        engine_module.Engine(0)
    assert "no CUDA device" in str(ei.value) or "-2" in str(ei.value)


def test_pfm_roundtrip(tmp_path):
    from pbrt_v4_b200 import imageio
    img = np.random.default_rng(0).random((5, 7, 3)).astype(np.float32)
    p = str(tmp_path / "a.pfm")
    imageio.write_pfm(p, img)
    assert (imageio.read_pfm(p) == img).all()
    assert imageio.rmse(img, img) == 0


def test_scene_generators(tmp_path):
    from pbrt_v4_b200 import scenes
    p = scenes.cornell(str(tmp_path / "c"), res=(32, 32), spp=4)
    txt = open(p).read()
    assert txt.count("Shape \"trianglemesh\"") == 6 and "AreaLightSource" in txt
    P, idx = scenes.heightfield_mesh(708)
    assert len(idx) == 999698 and P.shape == (708 * 708, 3)
    n = np.cross(P[idx[:, 1]] - P[idx[:, 0]], P[idx[:, 2]] - P[idx[:, 0]])
    assert (n[:, 1] > 0).all()          # geometric normals point up
    p = scenes.diffuse_heightfield(str(tmp_path / "h"), n=16, res=(32, 32), spp=4)
    hdr = open(os.path.join(os.path.dirname(p), "heightfield.ply"), "rb").read(200)
    assert b"element vertex 256" in hdr and b"element face 450" in hdr


def test_partition_bands_is_a_partition():
    from pbrt_v4_b200.parallel import partition_bands
    for yres, world, band in ((1024, 8, 16), (1080, 4, 16), (77, 3, 5), (16, 8, 16)):
        rows = np.zeros(yres, int)
        for r in range(world):
            for a, b in partition_bands(yres, world, r, band):
                rows[a:b] += 1
        assert (rows == 1).all()


GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from pbrt_v4_b200.parallel import partition_bands, reduce_film, gather_film
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
yres, xres = 40, 24
full = torch.from_numpy(np.random.default_rng(5).random((yres, xres, 4)))       # the film a single rank would produce
film = torch.zeros_like(full)
for a, b in partition_bands(yres, world, rank, band=4):
    film[a:b] = full[a:b]                                                          # this rank's tiles, zero elsewhere
tiles = film.clone()
reduce_film(film, dist)
gather_film(tiles, dist, band=4)                                                   # the tile exchange bench.py uses
# ragged case: 40 rows in bands of 16 over 2 ranks -> 32 + 8 rows
full2 = torch.from_numpy(np.random.default_rng(6).random((yres, xres, 4)))
t2 = torch.zeros_like(full2)
for a, b in partition_bands(yres, world, rank, band=16):
    t2[a:b] = full2[a:b]
gather_film(t2, dist, band=16)
if rank == 0:
    assert torch.equal(film, full), "reduced film differs from the single-rank film"
    assert torch.equal(tiles, full), "gathered film differs from the single-rank film"
    assert torch.equal(t2, full2), "gathered film differs from the single-rank film (ragged bands)"
    print("GLOO_REDUCE_OK")
dist.destroy_process_group()
'''


def test_world_size_2_film_reduce_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(w), ROOT], capture_output=True, text=True, env=env, timeout=300)
    assert "GLOO_REDUCE_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_bench_reference_arm_is_wired():
    # --impl reference must always print one JSON line (the reference binary exists in this container)
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    import bench
    path = bench.make_scene("heightfield_small", spp=4, res=(64, 64))
    ref = bench.ReferenceCPU(path, 4, threads=2, target_s=0.05)
    assert ref.available()
    r = ref.step()
    assert ref.spp == 4 and r["rays"] > 0 and r["mrays"] > 0 and r["seconds"] > 0
    # the whole arm, as the driver launches it: one JSON line on stdout
    import json, subprocess, sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--workload", "heightfield_small"], capture_output=True, text=True, timeout=600)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "reference" and line["e2e"]["h2d_bytes_per_step"] == 0


def test_interleaved_band_row_map_is_a_partition():
    """bpt_render_bands maps local row lr of part p to film row y0 + ((lr // band) * parts + p) * band + lr % band
    (MapRow in csrc/wavefront.cuh, row count in RenderImpl): every film row must be produced exactly once, by the part
    parallel.partition_bands assigns it to -- including a ragged last band."""
    from pbrt_v4_b200.parallel import partition_bands
    for (y0, y1, band, parts) in [(0, 1024, 16, 8), (0, 1080, 16, 8), (3, 77, 16, 3), (0, 10, 16, 4), (5, 6, 1, 2)]:
        seen = {}
        for p in range(parts):
            n_mine = sum(min(band, y1 - y) for k, y in enumerate(range(y0, y1, band)) if k % parts == p)
            rows = [y0 + ((lr // band) * parts + p) * band + lr % band for lr in range(n_mine)]
            expect = [y for (a, b) in partition_bands(y1, parts, p, band=band, y0=y0) for y in range(a, b)]
            assert rows == expect, (y0, y1, band, parts, p)
            for y in rows:
                assert y not in seen
                seen[y] = p
        assert sorted(seen) == list(range(y0, y1))


@pytest.fixture(scope="module")
def bvh_host_tool(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("bvhhost") / "bvh_build_host")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I/usr/local/cuda/include", "-I" + os.path.join(ROOT, "pbrt-v4_b200", "csrc"),
                           os.path.join(ROOT, "tests", "tools", "bvh_build_host.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("scene,size", [("heightfield", 96), ("soup", 6000), ("boxes", 3000), ("soup", 1), ("soup", 2), ("soup", 9)])
def test_bvh_builder_core_builds_valid_trees_and_treelets_lower_the_sah_cost(bvh_host_tool, scene, size):
    """The builder's per-element code (csrc/bvh_build_core.h -- the same functions the CUDA kernels call) run on the host:
    every primitive lands in exactly one leaf, every vertex lies inside its quantised leaf box, child frames lie inside their
    parent slots (the tool exits non-zero otherwise); treelet restructuring never raises the SAH objective it minimises, and
    cubic Morton cells beat per-axis normalisation on the flat bench geometry (the round-1 builder's weakness)."""
    import json

    def run(*args):
        r = subprocess.run([bvh_host_tool, scene, str(size), *args], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        return json.loads(r.stdout.strip().splitlines()[-1])
    plain, opt = run("0"), run("3")
    assert opt["sah_internal_area_sum"] <= plain["sah_internal_area_sum"] * 1.0001
    if scene == "heightfield":
        per_axis = run("0", "peraxis")
        assert plain["sah_internal_area_sum"] < 0.7 * per_axis["sah_internal_area_sum"]
        assert opt["camera"]["nodes_per_ray"] < per_axis["camera"]["nodes_per_ray"]


GBUFFER_NAMES = ["R", "G", "B", "Albedo.R", "Albedo.G", "Albedo.B", "P.X", "P.Y", "P.Z", "dzdx", "dzdy", "N.X", "N.Y", "N.Z", "Ns.X", "Ns.Y",
                 "Ns.Z", "u", "v", "Variance.R", "Variance.G", "Variance.B", "RelativeVariance.R", "RelativeVariance.G", "RelativeVariance.B"]


def test_exr_writer_is_read_back_by_openexr_and_by_an_independent_parser(tmp_path, engine_module):
    """csrc/exr_writer.h (f2: the multi-channel EXR a denoiser reads, film.cpp:688-803): an RGB file is read by OpenCV's bundled
    OpenEXR library bit for bit, and the 25-channel GBuffer layout by an independent parser of the file format (channels stored
    alphabetically, found by name)."""
    from pbrt_v4_b200 import imageio
    rng = np.random.default_rng(2)
    rgb = rng.normal(size=(9, 6, 3)).astype(np.float32)
    rgb[0, 0] = [np.inf, -0.0, 1e-42]    # specials survive (FLOAT channels, no compression)
    imageio.write_exr(tmp_path / "rgb.exr", rgb, ["R", "G", "B"])
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    try:
        import cv2
        a = cv2.imread(str(tmp_path / "rgb.exr"), cv2.IMREAD_UNCHANGED)
    except Exception:
        a = None
    if a is not None:
        assert a.dtype == np.float32 and (a[..., ::-1].copy().view(np.uint32) == rgb.view(np.uint32)).all()
    g = rng.normal(size=(5, 7, 25)).astype(np.float32)
    imageio.write_exr(tmp_path / "g.exr", g, GBUFFER_NAMES)
    names, back, attrs = imageio.read_exr(tmp_path / "g.exr")
    assert sorted(names) == sorted(GBUFFER_NAMES) and names == sorted(names)
    for k, nm in enumerate(names):
        assert (back[..., k].view(np.uint32) == g[..., GBUFFER_NAMES.index(nm)].view(np.uint32)).all(), nm
    for req in ("channels", "compression", "dataWindow", "displayWindow", "lineOrder", "pixelAspectRatio", "screenWindowCenter", "screenWindowWidth"):
        assert req in attrs
