"""Times every library variant in pbrt-v4_b200/variants/ on the bench scene (one subprocess per variant)."""
import glob, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import bench
from pbrt_v4_b200 import engine as E
path = bench.make_scene(os.environ.get("BPT_WORKLOAD", "heightfield1m"), spp=int(os.environ.get("SCENE_SPP", "256")), res=tuple(int(v) for v in os.environ["SCENE_RES"].split("x")) if "SCENE_RES" in os.environ else None)
e = E.Engine(0); e.load_pbrt(path, seed=0)
spp = int(os.environ.get("SPP", "32"))
e.render(0, spp)
e.stats_reset(); e.render(0, spp, profile=True)
s = e.stats()
out = {k: round(s[k], 3) for k in ("seconds_render", "ms_trace_closest", "ms_trace_shadow", "ms_shade")}
e.stats_reset(); e.render(0, 2, count_stats=True)
c = e.stats()
out["nodes_per_ray"] = round(c["nodes_visited"] / max(c["counted_rays"], 1), 3)
out["tris_per_ray"] = round(c["tris_tested"] / max(c["counted_rays"], 1), 3)
print(json.dumps(out))
''' % ROOT
for so in sorted(glob.glob(os.path.join(ROOT, "pbrt-v4_b200", "libb200pt_*.so"))):
    env = dict(os.environ, BPT_LIB=so)
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(os.path.basename(so), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
# the in-tree library itself, plain and with each runtime knob in SWEEP_ENVS ("BPT_RAY_SORT=1,OTHER=2")
for spec in [""] + [v for v in os.environ.get("SWEEP_ENVS", "").split(",") if v]:
    env = dict(os.environ)
    env.pop("BPT_LIB", None)
    if spec:
        k, _, v = spec.partition("=")
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print("libb200pt.so", spec or "(default)", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:], flush=True)
