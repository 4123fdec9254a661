"""Where does the un-jittered stratified sampler (sample values on exact k/16 + 1/32 lattices) leave the reference?  Cornell box with
maxdepth 0 / 1 / 5, RGB and GBuffer, engine vs live reference: fraction of bit-identical pixels per channel."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes, imageio
REF = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
out = {}
for smp in ('"stratified" "bool jitter" false', '"stratified" "bool jitter" false "integer xsamples" 3 "integer ysamples" 3'):
    for depth in (0, 1, 5):
        for gb in (False, True):
            d = os.path.join(ROOT, "gpurun_out", "strat_diag", f"{len(out)}")
            path = scenes.cornell(d, res=(48, 48), spp=16, maxdepth=depth)
            txt = open(path).read()
            a = txt.index("Sampler "); b = txt.index("\n", a)
            open(path, "w").write(txt[:a] + "Sampler " + smp + txt[b:])
            if gb:
                path = scenes.as_gbuffer(path, "world")
            outp = os.path.join(d, "ref.exr" if gb else "ref.pfm")
            subprocess.check_call([REF, "--quiet", "--seed", "0", "--outfile", outp, path])
            e = E.Engine(0); e.load_pbrt(path, seed=0); e.render(0, e.sampler_spp())
            if gb:
                names, ref = imageio.read_raw_channels(outp); img = e.gbuffer()
                row = {nm: float((img[..., c].view(np.uint32) == ref[..., c].view(np.uint32)).mean()) for c, nm in enumerate(names)}
                row = {k: v for k, v in row.items() if v < 1}
            else:
                ref = imageio.read_pfm(outp); img = e.image()
                row = {"identical": float((img.view(np.uint32) == ref.view(np.uint32)).all(axis=-1).mean()), "max": float(np.abs(img - ref).max())}
            e.close()
            key = f"{smp[-12:]} depth{depth} {'gbuffer' if gb else 'rgb'}"
            out[key] = row
            print(key, row, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "strat_diag.json"), "w"), indent=1)
