"""Are film differences between two BVHs (and against the reference's BVH) exact-t ties, or missed hits?
Traces the same rays (camera-like rays and rays leaving the surface) through two engines built with different builder
settings and classifies every ray whose hit record differs: |t1 - t2| <= 4 ulp and both triangles valid = tie between neighbouring
triangles (the accepted triangle depends on the test order, as in the reference); anything else would be a non-conservative box
test.  python scripts/tie_audit.py [out.json]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes

def rays_for(n, seed):
    rng = np.random.default_rng(seed)
    r = np.zeros((n, 8), np.float32)
    h = n // 2
    # camera-like: from above towards the field
    r[:h, 0:3] = np.array([0, 1.6, 2.6], np.float32) + rng.normal(scale=1e-3, size=(h, 3))
    tgt = np.stack([rng.uniform(-1, 1, h), np.zeros(h), rng.uniform(-1, 1, h)], 1)
    d = tgt - r[:h, 0:3]
    r[:h, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    # grazing rays starting just above the surface
    r[h:, 0:3] = np.stack([rng.uniform(-1, 1, n - h), rng.uniform(0.0, 0.08, n - h), rng.uniform(-1, 1, n - h)], 1)
    d = rng.normal(size=(n - h, 3)); d[:, 1] = np.abs(d[:, 1]) * 0.2 - 0.05
    r[h:, 4:7] = d / np.linalg.norm(d, axis=1, keepdims=True)
    r[:, 3] = np.inf
    return r

def main():
    out = {}
    d = os.path.join(ROOT, "gpurun_out", "tie_audit")
    path = scenes.diffuse_heightfield(os.path.join(d, "hf"), n=708, res=(64, 64), spp=1)
    rays = rays_for(4_000_000, 5)
    res = {}
    for name, env in (("lbvh_leaf3", dict(BPT_TREELET_ROUNDS="0", BPT_LEAF_MAX="3")), ("treelet2_leaf1", dict(BPT_TREELET_ROUNDS="2", BPT_LEAF_MAX="1")),
                      ("treelet3_leaf2", dict(BPT_TREELET_ROUNDS="3", BPT_LEAF_MAX="2"))):
        os.environ.update(env)
        e = E.Engine(0); e.load_pbrt(path, seed=0)
        prim, bary, t = e.trace_closest(rays)
        res[name] = (prim.copy(), bary.copy(), t.copy())
        e.close()
    base = res["lbvh_leaf3"]
    for name in ("treelet2_leaf1", "treelet3_leaf2"):
        p, b, t = res[name]
        diff = np.nonzero((p != base[0]) | (t.view(np.uint32) != base[2].view(np.uint32)))[0]
        both_hit = (p[diff] >= 0) & (base[0][diff] >= 0)
        ulp = np.abs(t[diff].view(np.int32).astype(np.int64) - base[2][diff].view(np.int32).astype(np.int64))
        out[name] = {"rays": int(len(rays)), "differ": int(len(diff)), "both_hit": int(both_hit.sum()),
                     "max_ulp_t_when_both_hit": int(ulp[both_hit].max()) if both_hit.any() else 0,
                     "hit_vs_miss": int((~both_hit).sum()),
                     "neighbour_triangles": int((np.abs(p[diff] - base[0][diff]) <= 2 * 707 * 707).sum()),
                     "examples": [[int(base[0][i]), int(p[i]), float(base[2][i]), float(t[i])] for i in diff[:6]]}
        print(name, json.dumps(out[name]), flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "tie_audit.json"), "w"), indent=1)

main()
