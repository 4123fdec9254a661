"""Generates the committed golden images with the reference CPU renderer (oracle/_ref/bin/pbrt, built from the unmodified
reference by oracle/build_ref.py).  Run from the repo root in the build container: python tests/golden/make_golden_images.py"""
import os, subprocess, sys, shutil, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import scenes
exe = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
out = os.path.join(ROOT, "tests", "golden")
tmp = tempfile.mkdtemp()
jobs = {"cornell_64x64_4spp": scenes.cornell(os.path.join(tmp, "c1"), res=(64, 64), spp=4),
        "cornell_128x128_16spp": scenes.cornell(os.path.join(tmp, "c2"), res=(128, 128), spp=16),
        "heightfield64_96x96_8spp": scenes.diffuse_heightfield(os.path.join(tmp, "h1"), n=64, res=(96, 96), spp=8)}
# scenes whose float paths are chaotic in the last ulp (hash-seeded RNG streams: layered BxDFs, media): a second, independent
# reference render (seed 1) gives the Monte-Carlo noise level the statistical tests compare against
jobs.update({
    "mixed48_96x96_16spp": scenes.mixed_envmap(os.path.join(tmp, "m1"), n=48, res=(96, 96), spp=16, envres=64, nmesh=4),
    "instanced50_96x96_16spp": scenes.instanced(os.path.join(tmp, "i1"), n_inst=50, subdiv=2, res=(96, 96), spp=16, material="diffuse"),
    "coated50_64x64_64spp": scenes.instanced(os.path.join(tmp, "i2"), n_inst=50, subdiv=2, res=(64, 64), spp=64, material="coatedconductor"),
    "smoke16_64x64_64spp": scenes.smoke(os.path.join(tmp, "s1"), grid=16, res=(64, 64), spp=64),
    "smokehomog_64x64_64spp": scenes.smoke(os.path.join(tmp, "s2"), res=(64, 64), spp=64, homogeneous=True),
})
second_seed = {"coated50_64x64_64spp", "smoke16_64x64_64spp", "smokehomog_64x64_64spp"}
for name, path in jobs.items():
    subprocess.check_call([exe, "--quiet", "--seed", "0", "--outfile", os.path.join(out, name + ".pfm"), path])
    if name in second_seed:
        subprocess.check_call([exe, "--quiet", "--seed", "1", "--outfile", os.path.join(out, name + "_seed1.pfm"), path])
    print(name, "done")
# GBufferFilm: the oracle build writes "*.exr" outputs as a raw 25-channel dump (oracle/build_ref.py)
gb = {"gbuffer_cornell_64x64_4spp_camera": (scenes.cornell(os.path.join(tmp, "g1"), res=(64, 64), spp=4), "camera"),
      "gbuffer_heightfield64_64x64_4spp_world": (scenes.diffuse_heightfield(os.path.join(tmp, "g2"), n=64, res=(64, 64), spp=4), "world")}
for name, (path, csys) in gb.items():
    subprocess.check_call([exe, "--quiet", "--seed", "0", "--outfile", os.path.join(out, name + ".raw.exr"), scenes.as_gbuffer(path, csys)])
    os.replace(os.path.join(out, name + ".raw.exr"), os.path.join(out, name + ".raw"))
    print(name, "done")
shutil.rmtree(tmp)
