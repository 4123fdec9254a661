"""Per-scene parity table: renders every BASELINE configuration class with the engine (through the C ABI) and with the reference
CPU binary (oracle/_ref/bin/pbrt, live on this box) at the same seed and prints / stores per-channel RMSE, the number of
differing pixels and both ray counts.   python scripts/parity_table.py [out.json] [scene ...]"""
import json
import os
import re
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pbrt_v4_b200 import engine as E, scenes, imageio   # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
SCENES = {
    "cornell_256_64spp": lambda d: (scenes.cornell(d, res=(256, 256), spp=64), 64),
    "heightfield1m_256_16spp": lambda d: (scenes.diffuse_heightfield(d, n=708, res=(256, 256), spp=16), 16),
    "heightfield1m_256_64spp": lambda d: (scenes.diffuse_heightfield(d, n=708, res=(256, 256), spp=64), 64),
    "mixed257k_256_16spp": lambda d: (scenes.mixed_envmap(d, n=360, res=(256, 256), spp=16, envres=512, nmesh=16), 16),
    "instanced1k_diffuse_256_16spp": lambda d: (scenes.instanced(d, n_inst=1000, subdiv=2, res=(256, 256), spp=16, material="diffuse"), 16),
    "instanced1k_coated_256_16spp": lambda d: (scenes.instanced(d, n_inst=1000, subdiv=2, res=(256, 256), spp=16, material="coatedconductor"), 16),
    "coated50_64_64spp": lambda d: (scenes.instanced(d, n_inst=50, subdiv=2, res=(64, 64), spp=64, material="coatedconductor"), 64),
    "smoke16_64_64spp": lambda d: (scenes.smoke(d, grid=16, res=(64, 64), spp=64), 64),
    "smokehomog_64_64spp": lambda d: (scenes.smoke(d, res=(64, 64), spp=64, homogeneous=True), 64),
    "smoke64_128_64spp": lambda d: (scenes.smoke(d, grid=64, res=(128, 128), spp=64), 64),
}


def main():
    out_json = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_table.json")
    names = sys.argv[2:] or list(SCENES)
    base = os.path.join(ROOT, "gpurun_out", "parity_table")
    rows = {}
    for name in names:
        d = os.path.join(base, name)
        os.makedirs(d, exist_ok=True)
        path, spp = SCENES[name](d)
        t0 = time.time()
        p = subprocess.run([REF, "--quiet", "--stats", "--seed", "0", "--outfile", os.path.join(d, "ref.pfm"), path], capture_output=True, text=True)
        tref = time.time() - t0
        if p.returncode != 0:
            rows[name] = {"error": p.stderr[-500:]}
            continue
        m = re.search(r"Regular ray intersection tests\s+(\d+)", p.stdout + p.stderr)
        ref_rays = int(m.group(1)) if m else None
        e = E.Engine(0)
        e.load_pbrt(path, seed=0)
        e.render(0, spp)
        st = e.stats()
        img = e.image()
        e.close()
        ref = imageio.read_pfm(os.path.join(d, "ref.pfm"))
        diff = np.abs(img - ref)
        row = {"rmse": [float(v) for v in imageio.channel_rmse(img, ref)], "mean_ref": [float(v) for v in ref.mean(axis=(0, 1))],
               "mean_gpu": [float(v) for v in img.mean(axis=(0, 1))], "max_abs": float(diff.max()),
               "frac_identical_pixels": float((img == ref).all(axis=-1).mean()),
               "pixels_diff_gt_1e-3": int((diff.max(axis=-1) > 1e-3).sum()), "pixels_diff_gt_1e-5": int((diff.max(axis=-1) > 1e-5).sum()),
               "n_pixels": int(img.shape[0] * img.shape[1]),
               "rays_gpu": int(st["camera_rays"] + st["indirect_rays"] + st["shadow_rays"]), "rays_ref": ref_rays, "ref_seconds": round(tref, 2)}
        rows[name] = row
        print(name, json.dumps(row), flush=True)
        for f in os.listdir(d):
            if f.endswith(".ply") or f.endswith(".pfm") or f.endswith(".pbrt"):
                os.remove(os.path.join(d, f))
    json.dump(rows, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
