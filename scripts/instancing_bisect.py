"""Which part of an instance transform breaks sample-exactness?  Renders a 200-instance diffuse scene whose instances use only a
translation / translation + scale / translation + rotation / all three, with the engine and with the live reference, and counts the
pixels that are not bit-identical.  Also a single-instance identity case.  python scripts/instancing_bisect.py out.json"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes, imageio
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
out = {}
for parts in ("", "t", "s", "ts", "r", "tr", "trs"):
    for mat in ("diffuse",):
        d = os.path.join(ROOT, "gpurun_out", "inst_bisect", parts or "identity")
        path = scenes.instanced(d, n_inst=200, subdiv=2, res=(128, 128), spp=16, material=mat, parts=parts)
        subprocess.check_call([REF, "--quiet", "--seed", "0", "--outfile", os.path.join(d, "ref.pfm"), path])
        ref = imageio.read_pfm(os.path.join(d, "ref.pfm"))
        e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, 16); img = e.image(); e.close()
        diff = np.abs(img - ref).max(axis=-1)
        out[(parts or "identity") + "_" + mat] = {"identical": float((diff == 0).mean()), "gt_1e-5": int((diff > 1e-5).sum()), "max": float(diff.max()),
                                                  "rmse": [float(v) for v in imageio.channel_rmse(img, ref)]}
        print(parts or "identity", mat, out[(parts or "identity") + "_" + mat], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "inst_bisect.json"), "w"), indent=1)

# ---- the same question for the config-3 ingredients: which material / light breaks bit-exactness on the 257 k-triangle mixed scene?
import re
MATS = {"diffuse": 'Material "diffuse" "rgb reflectance" [0.5 0.5 0.5]', "dielectric": 'Material "dielectric" "float eta" 1.5 "float roughness" 0',
        "conductor005": 'Material "conductor" "spectrum eta" "metal-Cu-eta" "spectrum k" "metal-Cu-k" "float roughness" 0.05',
        "conductor03": 'Material "conductor" "spectrum eta" "metal-Cu-eta" "spectrum k" "metal-Cu-k" "float roughness" 0.3',
        "conductor0": 'Material "conductor" "spectrum eta" "metal-Cu-eta" "spectrum k" "metal-Cu-k" "float roughness" 0'}
out2 = {}
d = os.path.join(ROOT, "gpurun_out", "inst_bisect", "mixed")
base = scenes.mixed_envmap(d, n=128, res=(128, 128), spp=16, envres=256, nmesh=4)
txt = open(base).read()
for name, line in MATS.items():
    for light in ("envmap", "uniform"):
        t = re.sub(r'Material "(dielectric|conductor)"[^\n]*', line, txt)
        if light == "uniform":
            t = t.replace('LightSource "infinite" "string filename" "env.pfm"', 'LightSource "infinite" "rgb L" [0.8 0.9 1.0]')
        path = os.path.join(d, f"{name}_{light}.pbrt")
        open(path, "w").write(t)
        subprocess.check_call([REF, "--quiet", "--seed", "0", "--outfile", os.path.join(d, "ref.pfm"), path])
        ref = imageio.read_pfm(os.path.join(d, "ref.pfm"))
        e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, 16); img = e.image(); e.close()
        diff = np.abs(img - ref).max(axis=-1)
        out2[name + "_" + light] = {"identical": float((diff == 0).mean()), "gt_1e-5": int((diff > 1e-5).sum()), "max": float(diff.max()), "mean": float(ref.mean())}
        print(name, light, out2[name + "_" + light], flush=True)
json.dump({"instancing": out, "mixed": out2}, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "inst_bisect.json"), "w"), indent=1)
