"""Render a generated scene with the engine and with the reference CPU binary on the same box; print RMSE."""
import os, subprocess, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbrt_v4_b200 import engine as E, scenes, imageio

which = sys.argv[1] if len(sys.argv) > 1 else "heightfield"
res = int(os.environ.get("RES", "256")); spp = int(os.environ.get("SPP", "16")); n = int(os.environ.get("N", "708"))
out = f"gpurun_out/parity_{which}"
os.makedirs(out, exist_ok=True)
if which == "heightfield":
    path = scenes.diffuse_heightfield(out, n=n, res=(res, res), spp=spp)
elif which == "cornell":
    path = scenes.cornell(out, res=(res, res), spp=spp)
elif which == "instanced":
    path = scenes.instanced(out, n_inst=int(os.environ.get("NINST", "200")), subdiv=int(os.environ.get("SUBDIV", "2")), res=(res, res), spp=spp, material=os.environ.get("MAT", "diffuse"))
elif which == "mixed":
    path = scenes.mixed_envmap(out, n=n, res=(res, res), spp=spp, envres=int(os.environ.get("ENVRES", "128")), nmesh=int(os.environ.get("NMESH", "8")))
elif which == "smoke":
    path = scenes.smoke(out, grid=int(os.environ.get("GRID", "32")), res=(res, res), spp=spp, homogeneous=bool(int(os.environ.get("HOMOG", "0"))))
else:
    path = getattr(scenes, which)(out, res=(res, res), spp=spp)
t0 = time.time()
p = subprocess.run(["oracle/_ref/bin/pbrt", "--quiet", "--stats", "--seed", "0", "--outfile", f"{out}/ref.pfm", path], capture_output=True, text=True)
tref = time.time() - t0
stats = {k: v for k, v in [l.strip().rsplit("  ", 1) for l in (p.stdout + p.stderr).splitlines() if "ray intersection tests" in l or "Camera rays" in l]} if p.returncode == 0 else p.stderr[-2000:]
print("ref seconds", tref, stats)
e = E.Engine(0)
t0 = time.time(); e.load_pbrt(path, seed=0); tload = time.time() - t0
t0 = time.time(); e.render(0, spp); tr = time.time() - t0
st = e.stats()
print("gpu load s", tload, "render s", tr, json.dumps(st))
img = e.image()
ref = imageio.read_pfm(f"{out}/ref.pfm")
imageio.write_pfm(f"{out}/gpu.pfm", img)
d = np.abs(img - ref)
print("mean ref", ref.mean(axis=(0, 1)), "mean gpu", img.mean(axis=(0, 1)))
print("RMSE per channel", imageio.channel_rmse(img, ref), "max abs", float(d.max()), "frac identical", float((img == ref).all(axis=-1).mean()))
ys, xs = np.nonzero(d.max(axis=-1) > 1e-3)
print("pixels with |diff| > 1e-3:", len(ys), [(int(x), int(y), img[y, x].tolist(), ref[y, x].tolist()) for x, y in list(zip(xs, ys))[:5]])
import glob
for f in glob.glob(os.path.join(out, "*.ply")): os.remove(f)
