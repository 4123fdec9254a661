"""First-hit geometry of instanced scenes against the reference: GBufferFilm channels (P, N, Ns, uv, dzdx, dzdy are functions of the
primary hit only) for instance transforms made of scale / translation+scale / rotation / all three.  A channel that is not
bit-identical localises the difference to the instance transform of the interaction; identical geometry channels with a differing
RGB image localise it to what follows (light sampling context, shadow rays, the next bounce).
python scripts/instancing_gbuffer.py out.json"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes, imageio
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
out = {}
for parts in ("t", "s", "ts", "r", "trs"):
    for spp, depth in ((1, 0), (1, 1), (4, 5)):
        d = os.path.join(ROOT, "gpurun_out", "inst_gbuffer", f"{parts}_{spp}_{depth}")
        path = scenes.as_gbuffer(scenes.instanced(d, n_inst=200, subdiv=2, res=(128, 128), spp=spp, maxdepth=depth, material="diffuse", parts=parts), "world")
        subprocess.check_call([REF, "--quiet", "--seed", "0", "--outfile", os.path.join(d, "ref.exr"), path])
        names, ref = imageio.read_raw_channels(os.path.join(d, "ref.exr"))
        e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, spp); g = e.gbuffer(); e.close()
        row = {}
        for c, nm in enumerate(names):
            row[nm] = float((g[..., c].view(np.uint32) == ref[..., c].view(np.uint32)).mean())
        key = f"{parts}_spp{spp}_maxdepth{depth}"
        out[key] = row
        print(key, {k: round(v, 4) for k, v in row.items() if v < 1.0} or "all channels bit-identical", flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "inst_gbuffer.json"), "w"), indent=1)
