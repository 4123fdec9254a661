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
for name, path in jobs.items():
    subprocess.check_call([exe, "--quiet", "--seed", "0", "--outfile", os.path.join(out, name + ".pfm"), path])
    print(name, "done")
shutil.rmtree(tmp)
