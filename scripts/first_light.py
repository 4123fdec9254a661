"""First GPU smoke run: Cornell box through the engine vs the reference binary on the same box."""
import os, subprocess, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbrt_v4_b200 import engine as E, scenes, imageio

out = "gpurun_out/first_light"
os.makedirs(out, exist_ok=True)
res = int(os.environ.get("RES", "256")); spp = int(os.environ.get("SPP", "64"))
path = scenes.cornell(out, res=(res, res), spp=spp)
t0 = time.time()
subprocess.check_call(["oracle/_ref/bin/pbrt", "--quiet", "--seed", "0", "--outfile", f"{out}/cornell_ref.pfm", path])
print("ref seconds", time.time() - t0)
e = E.Engine(0)
e.load_pbrt(path, seed=0)
t0 = time.time()
e.render(0, spp)
print("gpu seconds", time.time() - t0, e.stats())
img = e.image()
imageio.write_pfm(f"{out}/cornell_gpu.pfm", img)
ref = imageio.read_pfm(f"{out}/cornell_ref.pfm")
print("mean ref", ref.mean(axis=(0, 1)), "mean gpu", img.mean(axis=(0, 1)))
print("rmse", imageio.channel_rmse(img, ref), "max abs", np.abs(img - ref).max(), "frac identical", float((img == ref).all(axis=-1).mean()))
d = np.abs(img - ref).max(axis=-1)
ys, xs = np.nonzero(d > 1e-3)
print("pixels > 1e-3:", len(ys), list(zip(xs[:10], ys[:10])))
