"""Small deterministic target for ncu: one scene load + a few wavefront passes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
import bench
from pbrt_v4_b200 import engine as E
spp = int(os.environ.get("SPP", "4"))
path = bench.make_scene(os.environ.get("BPT_WORKLOAD", "heightfield1m"), spp=int(os.environ.get("SCENE_SPP", "256")))
e = E.Engine(0)
e.load_pbrt(path, seed=0)
e.render(0, spp)
print(e.stats())
