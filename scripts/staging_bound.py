"""Upper bound on what ANY staging of node packets / triangle vertices (TMA, cp.async.bulk, L1 prefetch) could gain for the
closest-hit kernel: the same kernel on the same ray distribution with BVHs from 60 MB (L2-resident) down to 0.5 MB (L1-resident).
If the cost per traversal step (node visit or triangle test) does not fall as the whole hierarchy moves into L1, memory staging is
not what limits the kernel.   python scripts/staging_bound.py out.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes
out = {}
for n in (708, 354, 177, 89, 45):
    d = os.path.join(ROOT, "gpurun_out", "staging_bound", f"hf{n}")
    path = scenes.diffuse_heightfield(d, n=n, res=(1024, 1024), spp=32)
    e = E.Engine(0); e.load_pbrt(path, seed=0)
    st0 = e.stats()
    e.render(0, 32)
    e.stats_reset(); e.render(0, 32, profile=True)
    s = e.stats()
    e.stats_reset(); e.render(0, 4, count_stats=True)
    c = e.stats()
    rays = s["camera_rays"] + s["indirect_rays"]
    N, T = c["nodes_visited"] / max(c["counted_rays"], 1), c["tris_tested"] / max(c["counted_rays"], 1)
    row = {"triangles": st0["bvh_triangles"], "bvh_MB": round((st0["bytes_bvh_nodes"] + st0["bytes_bvh_tris"]) / 1e6, 2),
           "closest_ns_per_ray": round(s["ms_trace_closest"] * 1e6 / rays, 4), "nodes_per_ray": round(N, 2), "tris_per_ray": round(T, 2),
           "ps_per_step": round(s["ms_trace_closest"] * 1e9 / rays / (N + T), 2)}
    out[f"heightfield_{n}"] = row
    print(n, row, flush=True)
    e.close()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "staging_bound.json"), "w"), indent=1)
