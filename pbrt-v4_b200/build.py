"""Builds libb200pt.so (the C-ABI engine: CUDA kernels for sm_100a + host C++) in-tree with nvcc.

Flags that matter for parity: -fmad=false (device) and -ffp-contract=off (host) so that only the explicit
fmaf()/FMA calls fuse, as in the reference CPU build; no --use_fast_math.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200pt.so")
SOURCES = ["shade_tex_part1.cu", "shade_tex_part2.cu", "shade_tex_part3.cu", "shade_tex.cu", "wavefront_shapes.cu", "wavefront.cu", "engine.cu", "bvh_build.cu", "pbrt_frontend.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17", "-DBPT_FRONTEND_IN_LIBRARY=1",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=default", "-diag-suppress", "20096"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for root, _, files in os.walk(CSRC):
        for f in files:
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    inc = os.path.join(HERE, "..", "include", "b200pt.h")
    return os.path.getmtime(inc) > t


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(objdir, s + ".o")
        objs.append(o)
        cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, s), "-o", o]
        if s.endswith(".cpp"):
            cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, cmd))
    # at most BPT_BUILD_JOBS compilers at a time (ptxas on the shading kernels takes several GB each; nine at once were OOM-killed)
    jobs = int(os.environ.get("BPT_BUILD_JOBS", "5"))
    running, failed = [], None
    pending = list(procs)
    while pending or running:
        while pending and len(running) < jobs:
            s, cmd = pending.pop(0)
            running.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        s, p = running.pop(0)
        out, _ = p.communicate()
        if p.returncode != 0 and failed is None:
            sys.stderr.write(out.decode())
            failed = s
    if failed:
        raise RuntimeError(f"nvcc failed on {failed}")
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ldl", "-lz"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
