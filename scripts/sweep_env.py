"""Times the in-tree library on a bench workload for a list of environment settings, one subprocess each.
usage: SWEEP="A=1,B=2;A=2" BPT_WORKLOAD=heightfield1m SPP=32 python scripts/sweep_env.py"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import bench
from pbrt_v4_b200 import engine as E
res = tuple(int(v) for v in os.environ["SCENE_RES"].split("x")) if "SCENE_RES" in os.environ else None
path = bench.make_scene(os.environ.get("BPT_WORKLOAD", "heightfield1m"), spp=int(os.environ.get("SCENE_SPP", "256")), res=res)
e = E.Engine(0); e.load_pbrt(path, seed=0)
b0 = e.stats()["seconds_build"]
e.load_pbrt(path, seed=0)
b1 = e.stats()["seconds_build"]
spp = int(os.environ.get("SPP", "32"))
e.render(0, spp)
e.stats_reset(); e.film_clear(); e.render(0, spp)
plain = e.stats()["seconds_render"]
e.stats_reset(); e.render(0, spp, profile=True)
s = e.stats()
out = {"render_s": round(plain, 4)}
out.update({k: round(s[k], 3) for k in ("ms_trace_closest", "ms_trace_shadow", "ms_shade", "ms_other")})
out["mrays"] = round((s["camera_rays"] + s["indirect_rays"] + s["shadow_rays"]) / plain / 1e6, 1)
e.stats_reset(); e.render(0, min(spp, 4), count_stats=True)
c = e.stats()
out["N"] = round(c["nodes_visited"] / max(c["counted_rays"], 1), 2); out["T"] = round(c["tris_tested"] / max(c["counted_rays"], 1), 2)
out["Ns"] = round(c["nodes_visited_shadow"] / max(c["counted_rays_shadow"], 1), 2); out["Ts"] = round(c["tris_tested_shadow"] / max(c["counted_rays_shadow"], 1), 2)
out["build_first_s"] = round(b0, 4); out["build_again_s"] = round(b1, 4); out["bvh_nodes"] = c["bvh_nodes"]
print(json.dumps(out))
''' % ROOT
for spec in os.environ.get("SWEEP", "").split(";"):
    env = dict(os.environ)
    for kv in [v for v in spec.split(",") if v]:
        k, _, v = kv.partition("=")
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(spec or "(default)", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-600:], flush=True)
