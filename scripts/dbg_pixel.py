import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbrt_v4_b200 import engine as E, scenes, imageio
out = "gpurun_out/dbg"; os.makedirs(out, exist_ok=True)
path = scenes.mixed_envmap(out, n=64, res=(128, 128), spp=16, envres=128, nmesh=8)
e = E.Engine(0); e.load_pbrt(path, seed=0)
for pb in ((0,1,0,1),(0,2,0,1),(0,3,0,1),(0,32,0,1),(0,33,0,1),(0,128,0,1),(0,128,0,2),(0,128,0,8),(0,128,0,64),(0,128,0,128)):
    for ns in (1, 16):
        e.film_clear(); e.render(0, ns, pb); print(pb, ns, e.film()[0, 0][:2], e.stats()["kernel_launches"])
