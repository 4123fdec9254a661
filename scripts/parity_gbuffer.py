"""GBufferFilm parity: the reference CPU binary (raw dump of its 25 channels) against Engine.gbuffer()."""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbrt_v4_b200 import engine as E, scenes, imageio

which = sys.argv[1] if len(sys.argv) > 1 else "cornell"
res = int(os.environ.get("RES", "64")); spp = int(os.environ.get("SPP", "4")); csys = os.environ.get("CSYS", "camera")
out = f"gpurun_out/gbuffer_{which}"
os.makedirs(out, exist_ok=True)
path = scenes.cornell(out, res=(res, res), spp=spp) if which == "cornell" else scenes.diffuse_heightfield(out, n=64, res=(res, res), spp=spp)
s = open(path).read().replace('Film "rgb"', f'Film "gbuffer" "string coordinatesystem" "{csys}"').replace('"out.pfm"', '"out.exr"')
path = os.path.join(out, "gb.pbrt"); open(path, "w").write(s)
subprocess.check_call(["oracle/_ref/bin/pbrt", "--quiet", "--seed", "0", "--outfile", f"{out}/ref.exr", path])
names, ref = imageio.read_raw_channels(f"{out}/ref.exr")
e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, spp)
g = e.gbuffer()
assert names == E.Engine.GBUFFER_CHANNELS, names
for c, nm in enumerate(names):
    d = np.abs(g[..., c] - ref[..., c])
    print(f"{nm:22s} ref mean {ref[..., c].mean():12.6g}  gpu mean {g[..., c].mean():12.6g}  rmse {np.sqrt((d ** 2).mean()):10.3g}  max {d.max():10.3g}  identical {float((g[..., c] == ref[..., c]).mean()):.3f}")
