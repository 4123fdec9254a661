import os, subprocess, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbrt_v4_b200 import engine as E
os.makedirs("gpurun_out/pm", exist_ok=True)
subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "scripts/libm_ref.c", "-o", "gpurun_out/pm/libm_ref", "-lm"])
e = E.Engine(0)
lib = e.lib
lib.bpt_probe_math.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
rng = np.random.default_rng(1)
n = 1 << 20
cases = {  # libm fn id -> (domain sampler, [device candidate ids])
    0: (lambda: rng.uniform(-7, 7, n), [0, 10, 30]), 1: (lambda: rng.uniform(-7, 7, n), [1, 11, 31]),
    2: (lambda: rng.uniform(-1, 1, n), [2, 12, 33]), 3: (lambda: rng.uniform(-1, 1, n), [3, 13, 34]),
    4: (lambda: rng.uniform(-2, 2, n), [4, 14, 35]), 5: (lambda: rng.uniform(-0.98, 0.86, n), [5, 15, 36]),
    6: (lambda: rng.uniform(-1.3, 2.2, n), [6, 16, 37]), 7: (lambda: rng.uniform(-20, 3, n), [7, 17, 38]),
    8: (lambda: rng.uniform(1e-6, 10, n), [8, 18, 32]), 9: (lambda: rng.uniform(-1.5, 1.5, n), [9, 23]),
    19: (lambda: rng.uniform(0.01, 3, n), [19, 20]), 21: (lambda: rng.uniform(0, 100, n), [21]), 22: (lambda: rng.uniform(-5, 5, n), [22]),
}
names = {0: "sinf", 1: "cosf", 2: "asinf", 3: "acosf", 4: "atan2f", 5: "atanhf", 6: "coshf", 7: "expf", 8: "logf", 9: "tanf", 19: "powf", 21: "sqrtf", 22: "div"}
for fn, (gen, cands) in cases.items():
    x = gen().astype(np.float32); y = rng.uniform(0.1, 3, n).astype(np.float32)
    if fn == 4: y = rng.uniform(-2, 2, n).astype(np.float32)
    np.stack([x, y], 1).astype(np.float32).tofile("gpurun_out/pm/in.bin")
    subprocess.check_call(["gpurun_out/pm/libm_ref", str(fn), "gpurun_out/pm/in.bin", "gpurun_out/pm/out.bin"])
    ref = np.fromfile("gpurun_out/pm/out.bin", dtype=np.float32)
    for c in cands:
        out = np.empty(n, np.float32)
        rc = lib.bpt_probe_math(e.h, c, x.ctypes.data, y.ctypes.data, n, out.ctypes.data)
        mism = (out.view(np.uint32) != ref.view(np.uint32)) & ~(np.isnan(out) & np.isnan(ref))
        print(f"{names[fn]:7s} device candidate {c:2d}: mismatch rate {mism.mean():.6f}  max ulp {np.abs(out.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64))[~np.isnan(ref)].max()}")
