"""The rendered film must not depend on the BVH: closest hits are decided by the watertight triangle test alone, the hierarchy only
culls.  Renders the same scene with different builder settings (treelet rounds, leaf size) and compares the films bit for bit.
A difference means a non-conservative box test (a missed hit) or an exact-t tie between two triangles."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes

out = {}
d = os.path.join(ROOT, "gpurun_out", "bvh_invariance")
cases = {"heightfield1m": lambda: (scenes.diffuse_heightfield(os.path.join(d, "hf"), n=708, res=(256, 256), spp=16), 16),
         "mixed257k": lambda: (scenes.mixed_envmap(os.path.join(d, "mx"), n=360, res=(256, 256), spp=16, envres=512, nmesh=16), 16),
         "instanced1k": lambda: (scenes.instanced(os.path.join(d, "in"), n_inst=1000, subdiv=2, res=(256, 256), spp=16, material="diffuse"), 16)}
for name, mk in cases.items():
    path, spp = mk()
    films = {}
    for rounds, leaf in ((0, 3), (2, 3), (2, 1), (3, 2)):
        os.environ["BPT_TREELET_ROUNDS"] = str(rounds); os.environ["BPT_LEAF_MAX"] = str(leaf)
        e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, spp)
        st = e.stats()
        films[(rounds, leaf)] = (e.film().copy(), st["bvh_nodes"], st["seconds_build"])
        e.close()
    base = films[(0, 3)][0]
    row = {}
    for k, (f, nodes, tb) in films.items():
        diff = np.argwhere((f != base).any(axis=-1))
        row["rounds%d_leaf%d" % k] = {"bvh_nodes": nodes, "build_s": tb, "pixels_differing_from_rounds0_leaf3": int(len(diff)),
                                      "where": [[int(x), int(y)] for y, x in diff[:8]]}
    out[name] = row
    print(name, json.dumps(row), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "bvh_invariance.json"), "w"), indent=1)
