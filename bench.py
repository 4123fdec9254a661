#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native wavefront path tracer (BASELINE.json metric: Mrays/s).

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path (oracle/_ref/bin/pbrt)

A "step" is one full render of the workload (BASELINE configs[1]: procedural ~1M-triangle diffuse height field, one
quad area light, 1024x1024, 256 spp, maxdepth 5, zsobol) -- all wavefront depths, shadow rays and the film update.
`value` = rays traced (camera + indirect + shadow, counted on the device) per second over all GPUs, scene resident in
HBM, timed with CUDA events inside the engine; `e2e` = the same metric through the public scene-file API with host
inputs: parse + host->device upload + BVH build + render + film read-back every step.
Multi-GPU: scanline bands of the film are split over the ranks (weak in rays per band is not fixed: total work is
fixed, so scaling is "strong"); one NCCL reduce of the per-rank film accumulators per step (frame end).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the headline metric is quoted on at N = 1
    "heightfield1m": dict(gen="diffuse_heightfield", kw=dict(n=708), res=(1024, 1024), spp=256,
                          desc="procedural 999698-triangle diffuse height field, 1 quad area light, 1024x1024, 256 spp, maxdepth 5, zsobol"),
    "heightfield_small": dict(gen="diffuse_heightfield", kw=dict(n=128), res=(256, 256), spp=16,
                              desc="procedural 32258-triangle diffuse height field, 256x256, 16 spp (smoke size)"),
    # the other BASELINE.json configs (parity-test cases; runnable here for measured numbers)
    "mixed10m": dict(gen="mixed_envmap", kw=dict(n=2236, envres=2048, nmesh=64), res=(1920, 1080), spp=1024,
                     desc="configs[2]: 10.0M-triangle height field in 64 meshes, dielectric / rough conductor, 2048^2 HDR image light, 1920x1080, 1024 spp"),
    "instanced20k": dict(gen="instanced", kw=dict(n_inst=20000, subdiv=3, material="coatedconductor"), res=(512, 512), spp=64, desc="20k instances (tuning size)"),
    "instanced100k": dict(gen="instanced", kw=dict(n_inst=100000, subdiv=3, material="coatedconductor"), res=(2048, 2048), spp=512,
                          desc="configs[3]: 100k instances of a 1280-triangle icosphere (two-level BVH), coated conductor, 2048x2048, 512 spp"),
    "smoke64": dict(gen="smoke", kw=dict(grid=64), res=(512, 512), spp=256, desc="64^3 smoke (tuning size)"),
    "smoke256": dict(gen="smoke", kw=dict(grid=256), res=(1024, 1024), spp=2048,
                     desc="configs[4]: 256^3 grid-density smoke in a unit interface box, null-scattering VolPath, 1024x1024, 2048 spp"),
}


def scene_dir():
    d = os.environ.get("BPT_BENCH_DIR", os.path.join(tempfile.gettempdir(), "bpt_bench"))
    os.makedirs(d, exist_ok=True)
    return d


def make_scene(name, spp=None, res=None):
    import glob
    from pbrt_v4_b200 import scenes
    w = WORKLOADS[name]
    spp = spp or w["spp"]
    res = res or w["res"]
    d = os.path.join(scene_dir(), f"{name}_{res[0]}x{res[1]}_{spp}")
    found = glob.glob(os.path.join(d, "*.pbrt"))
    if found:
        return found[0]
    return getattr(scenes, w["gen"])(d, res=res, spp=spp, **w["kw"])


def scene_bytes(path):
    d = os.path.dirname(path)
    return sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d))


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for k, nm in enumerate(names):
                if len(s) > 2 + k and s[2 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(self.samples)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_once(scene_path, spp, threads):
    """One run of the reference's CPU VolPathIntegrator (oracle/_ref/bin/pbrt) on the bench scene file, film as in the file.
    Rays = its own 'Regular ray intersection tests' statistic (VolPath uses Intersect for shadow rays too);
    render seconds = the elapsed time its ProgressReporter prints for the render loop (util/progressreporter.cpp:140-146), which
    excludes parsing and the BVH build; wall seconds = the whole process."""
    exe = os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt")
    out_pfm = os.path.join(scene_dir(), f"ref_{os.getpid()}.pfm")
    t0 = time.time()
    p = subprocess.run([exe, "--stats", "--seed", "0", "--nthreads", str(threads), "--spp", str(spp), "--outfile", out_pfm, scene_path],
                       capture_output=True, text=True)
    wall = time.time() - t0
    txt = p.stdout.replace("\r", "\n")
    m = re.search(r"Regular ray intersection tests\s+(\d+)", txt + p.stderr)
    rays = int(m.group(1)) if m else 0
    done = re.findall(r"Rendering: \[\++\]\s+\(([0-9.]+)s\)", txt)
    render_s = float(done[-1]) if done else 0.0
    try:
        os.remove(out_pfm)
    except OSError:
        pass
    if p.returncode != 0 or rays == 0:
        raise RuntimeError("reference pbrt failed: " + (p.stderr[-400:] or p.stdout[-400:]))
    return {"spp": spp, "rays": rays, "render_s": render_s, "wall_s": wall}


class ReferenceCPU:
    """The reference's own CPU path on the box's host cores, on the bench workload's scene and FILM (same config as the GPU arm),
    at a bounded sample count: the first call calibrates spp so that one render loop takes about `target_s` seconds (>= 5 s, so the
    0.1 s resolution of the reference's own timer is <= 2 %)."""

    def __init__(self, scene_path, full_spp, threads=None, target_s=7.0):
        self.path, self.full_spp, self.target_s = scene_path, full_spp, target_s
        self.threads = threads or os.cpu_count()
        self.spp = None
        self.calib = None

    def available(self):
        return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bin", "pbrt"))

    def calibrate(self):
        c = run_reference_once(self.path, 16 if self.full_spp >= 16 else self.full_spp, self.threads)
        rate = c["rays"] / max(c["render_s"], 0.05)            # rays / s, coarse
        rays_per_spp = c["rays"] / c["spp"]
        want = self.target_s * rate / rays_per_spp
        spp = 4
        while spp * 2 <= want and spp * 2 <= self.full_spp:
            spp *= 2
        self.spp, self.calib = spp, c
        return spp

    def step(self):
        if self.spp is None:
            self.calibrate()
        r = run_reference_once(self.path, self.spp, self.threads)
        # render-loop time from the reference's own timer; when it is too short to resolve, fall back to the wall clock of the
        # process minus the calibration run's non-render part (parse + BVH build)
        t = r["render_s"]
        if t < 1.0:
            t = max(r["wall_s"] - max(self.calib["wall_s"] - self.calib["render_s"], 0.0), 1e-3)
        r["mrays"] = r["rays"] / t / 1e6
        r["seconds"] = t
        return r

    def describe(self, runs):
        ts = ", ".join("%.1f" % r["seconds"] for r in runs)
        return (f"oracle/_ref/bin/pbrt (unmodified reference, CPU VolPathIntegrator), same scene and film as the GPU arm, --spp {self.spp} of "
                f"{self.full_spp} (calibrated for ~{self.target_s:.0f} s per render loop), {len(runs)} run(s), render-loop seconds [{ts}] from the "
                f"reference's own progress timer (parse and BVH build excluded), {runs[-1]['rays']} rays per run, {_cpu_model()}")


def cpu_baseline(scene_path, full_spp, runs=3):
    ref = ReferenceCPU(scene_path, full_spp)
    if not ref.available():
        return None
    rs = [ref.step() for _ in range(runs)]
    vals = sorted(r["mrays"] for r in rs)
    return {"value": vals[len(vals) // 2], "unit": "Mrays/s", "cores": ref.threads, "kind": "reference", "sample": ref.describe(rs),
            "runs_mrays": [round(r["mrays"], 3) for r in rs], "spp": ref.spp}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    path = make_scene(args.workload, spp=args.spp or w["spp"])
    ref = ReferenceCPU(path, args.spp or w["spp"])
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/pbrt is not built (python oracle/build_ref.py)"}))
        return
    ref.calibrate()
    runs = []
    for i in range(args.warmup + args.steps):
        r = ref.step()
        if i >= args.warmup:
            runs.append(r)
    vals = [r["mrays"] for r in runs]
    v = sum(vals) / len(vals)
    secs = sum(r["seconds"] for r in runs) / len(runs)
    base = {"value": v, "unit": "Mrays/s", "cores": ref.threads, "kind": "reference", "sample": ref.describe(runs),
            "runs_mrays": [round(x, 3) for x in vals], "spp": ref.spp}
    line = {"metric": "Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": w["desc"], "spp": args.spp or w["spp"],
                       "sample": f"same scene and {w['res'][0]}x{w['res'][1]} film, --spp {ref.spp} per step (bounded sample of the {args.spp or w['spp']} spp workload)"},
            "cpu_baseline": base, "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default=os.environ.get("BPT_WORKLOAD", "heightfield1m"))
    ap.add_argument("--spp", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)

    import numpy as np
    import torch
    from pbrt_v4_b200 import engine as E
    from pbrt_v4_b200.parallel import gather_film

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL announces its version on stdout when the first communicator is created; the contract is ONE JSON line on
        # stdout, so route fd 1 to stderr until the communicator exists
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    w = WORKLOADS[args.workload]
    spp = args.spp or w["spp"]
    path = make_scene(args.workload, spp=spp) if rank == 0 else None
    if world > 1:
        dist.barrier()
        path = make_scene(args.workload, spp=spp)
    xres, yres = w["res"]

    eng = E.Engine(local_rank)
    # builder effort by the number of samples that are going to be rendered: two treelet rounds cost ~6.5 ms per million triangles
    # and save ~4 % of the trace time, which pays from ~100 M pixel samples per GPU (profiles/r02c sweep)
    # (decided on the whole frame, not per GPU: exact-t ties on shared triangle edges resolve by traversal order, so the same
    # builder setting on every N is what keeps the film -- and `film_checksum` -- identical for every N)
    treelet_rounds = 2 if xres * yres * spp >= 100e6 else 0
    eng.set_option("bvh_treelet_rounds", treelet_rounds)
    eng.load_pbrt(path, seed=0)
    st0 = eng.stats()
    # scanline bands, interleaved in blocks of 16 rows for load balance
    band = 16
    my_bands = [(y, min(y + band, yres)) for i, y in enumerate(range(0, yres, band)) if i % world == rank]
    film_t = torch.zeros((yres, xres, 4), dtype=torch.float64, device="cuda")

    def step():
        eng.film_clear()
        if world == 1:
            eng.render(0, spp)
        else:
            eng.render_bands(0, spp, band, world, rank)
            # frame end: one NCCL exchange of the per-rank film tiles (each rank sends the 1/N of the film it rendered)
            eng.film_to_device_ptr(film_t.data_ptr())
            gather_film(film_t, dist, band)
            if rank == 0:
                eng.film_from_device_ptr(film_t.data_ptr())
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    eng.stats_reset()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    clocks.stop_flag = True
    s_after = eng.stats()
    rays_rank_total = sum(s_after[k] for k in ("camera_rays", "indirect_rays", "shadow_rays"))   # over all K steps
    dev_seconds = s_after["seconds_render"]                                                     # engine CUDA events, K steps
    tvals = torch.tensor([wall, float(rays_rank_total), dev_seconds], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = tvals.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = tvals.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        wall_max = float(tmax[0]); total_rays = float(tsum[1])
    else:
        wall_max = wall; total_rays = float(rays_rank_total)
    # N = 1: device time of the K steps from the engine's CUDA events on its launch stream;
    # N > 1: max over ranks of the barrier-bracketed region (includes the NCCL film reduce)
    step_seconds = (dev_seconds if world == 1 else wall_max) / args.steps
    value = total_rays / args.steps / step_seconds / 1e6
    msamples = xres * yres * spp / step_seconds / 1e6

    # ---- e2e (all ranks): scene file -> parse -> H2D upload -> BVH build -> render own bands -> NCCL film reduce -> rank 0
    # resolves and reads the image back to the host; every step, through the public API, wall clock, max over ranks.
    # The Engine object (CUDA context, streams, data tables) is created once, like the reference's process start-up.
    mesh_bytes = scene_bytes(path)
    e2e_steps = max(1, min(args.steps, 3))
    img_bytes = 0

    def e2e_step():
        nonlocal img_bytes
        eng.load_pbrt(path, seed=0)
        step()
        if rank == 0:
            img_bytes = eng.image().nbytes

    e2e_step()   # warm-up (first re-commit)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    te = (time.perf_counter() - te) / e2e_steps
    tev = torch.tensor([te], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tev, op=dist.ReduceOp.MAX)
    te = float(tev[0])
    st2 = eng.stats()
    e2e = {"value": total_rays / args.steps / te / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(mesh_bytes) * world,
           "d2h_bytes_per_step": int(img_bytes), "seconds_per_step": te, "seconds_bvh_build": st2["seconds_build"], "n_gpus": world,
           "note": "scene replicated: every rank parses and uploads it"}

    # checksum of the final film of the last step (rank 0 holds the gathered film): must be the same for every N
    film_checksum = None
    if rank == 0:
        import hashlib
        film_checksum = hashlib.sha256(eng.film().tobytes()).hexdigest()[:16]
    line = None
    if rank == 0:
        # ---- roofline: counting pass (nodes / triangles per ray) + per-kernel CUDA-event timing of a full step
        # (N > 1: over this rank's own bands -- 1/N of the frame, interleaved, so representative of the whole image)
        def render_mine(nspp, **kw):
            if world == 1:
                eng.render(0, nspp, None, **kw)
            else:
                eng.render_bands(0, nspp, band, world, rank, **kw)
        eng.stats_reset(); eng.film_clear()
        render_mine(min(spp, 16), count_stats=True)
        sc = eng.stats()
        nc, tc = sc["nodes_visited"] / max(1, sc["counted_rays"]), sc["tris_tested"] / max(1, sc["counted_rays"])
        ns, ts = sc["nodes_visited_shadow"] / max(1, sc["counted_rays_shadow"]), sc["tris_tested_shadow"] / max(1, sc["counted_rays_shadow"])
        eng.stats_reset(); eng.film_clear()
        render_mine(spp, profile=True)
        sp = eng.stats()
        peak, peak_src = measured_peak_gbs()
        rays_c = sp["camera_rays"] + sp["indirect_rays"]
        rays_s = sp["shadow_rays"]
        # algorithmic bytes per ray of each kernel: queue index 4 + ray 32 + result (hit record 16 + queue push 4 | L RMW 32)
        bytes_c = rays_c * (4 + 32 + 20 + nc * 80.0 + tc * 48.0)
        bytes_s = rays_s * (48 + 16 + ns * 80.0 + ts * 48.0)
        kern = {"k_trace_closest": {"ms_total": sp["ms_trace_closest"], "launches": sp["launches_trace_closest"], "rays": rays_c,
                                    "nodes_per_ray": nc, "tris_per_ray": tc, "gbs": bytes_c / max(1e-9, sp["ms_trace_closest"] * 1e-3) / 1e9},
                "k_trace_shadow": {"ms_total": sp["ms_trace_shadow"], "launches": sp["launches_trace_shadow"], "rays": rays_s,
                                   "nodes_per_ray": ns, "tris_per_ray": ts, "gbs": bytes_s / max(1e-9, sp["ms_trace_shadow"] * 1e-3) / 1e9},
                "k_shade": {"ms_total": sp["ms_shade"]}, "step_ms_profiled": sp["seconds_render"] * 1e3}
        dom = "k_trace_closest" if sp["ms_trace_closest"] >= sp["ms_trace_shadow"] else "k_trace_shadow"
        achieved = kern[dom]["gbs"]
        all_rays = rays_c + rays_s
        n_bar = (sc["nodes_visited"] + sc["nodes_visited_shadow"]) / max(1, sc["counted_rays"] + sc["counted_rays_shadow"])
        t_bar = (sc["tris_tested"] + sc["tris_tested_shadow"]) / max(1, sc["counted_rays"] + sc["counted_rays_shadow"])
        b_ray = 300.0 + n_bar * 80.0 + t_bar * 48.0
        # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/, first-bounce launch of the same
        # workload): dram read + write per ray there, times the rays of an average launch here
        traffic, traffic_file = None, None
        # newest capture first: (summary file, camera rays of its first-bounce launch = spp of the capture x 1024^2)
        for fn, ncu_rays in (("r02v_hf1m.summary.txt", 8 * 1048576.0), ("r02u_hf1m.summary.txt", 8 * 1048576.0), ("r02a_hf1m.summary.txt", 8 * 1048576.0), ("r01h_closest_shade_camera.summary.txt", 32 * 1048576.0)):
            try:
                txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", fn)).read()
                blk = txt[txt.index("k_trace_closest"):]
                rd = float(re.search(r"dram__bytes_read.sum = ([0-9.]+)\s+\[Gbyte\]", blk).group(1))
                wr = float(re.search(r"dram__bytes_write.sum = ([0-9.]+)\s+\[Gbyte\]", blk).group(1))
                traffic = (rd + wr) * 1e9 / ncu_rays * (rays_c / max(1, sp["launches_trace_closest"]))
                traffic_file = fn
                break
            except Exception:
                continue
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic if (dom == "k_trace_closest" and args.workload == "heightfield1m") else None,
                    "traffic_source": "ncu --set full capture profiles/%s (dram read+write per ray of the first-bounce launch) x rays of an average launch; "
                                      "a few percent of the algorithmic bytes because the 60 MB BVH stays in the 126 MB L2" % traffic_file,
                    "kernel": dom, "peak_source": peak_src, "kernels": kern,
                    "launch_avg_ms": kern[dom]["ms_total"] / max(1, kern[dom]["launches"]),
                    "bytes_per_launch_avg": (bytes_c if dom == "k_trace_closest" else bytes_s) / max(1, kern[dom]["launches"]),
                    # whole-step figure of SURVEY 8(d): B_ray = 300 + N*80 + T*48 over all rays of the step
                    "step": {"bytes_per_ray": b_ray, "nodes_per_ray": n_bar, "tris_per_ray": t_bar,
                             "achieved_gbs": value * 1e6 / world * b_ray / 1e9, "frac": value * 1e6 / world * b_ray / 1e9 / peak,
                             "roofline_mrays_per_gpu": peak * 1e9 / b_ray / 1e6}}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline(path, spp)
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": step_seconds * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "msamples_per_s": msamples,
                "config": {"workload": w["desc"], "spp": spp, "rays_per_step": total_rays / args.steps, "tiling": f"{band}-row bands round-robin over ranks, tile gather to rank 0 at frame end",
                           "bvh_treelet_rounds": treelet_rounds,
                           "l2": "inputs larger than L2: BVH+triangles %.0f MB + path state %.0f MB per pass" % ((st0["bytes_bvh_nodes"] + st0["bytes_bvh_tris"]) / 1e6, min(32 << 20, xres * yres * spp) * 240 / 1e6),
                           "bvh": {"nodes": st0["bvh_nodes"], "depth": st0["bvh_depth"], "build_seconds": st0["seconds_build"]}},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(s_after["kernel_launches"]),
                "film_checksum": film_checksum,
                "clocks": clocks.summary(), "wall_seconds_timed_region": wall_max}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
