"""Builds tuning variants of libb200pt.so (different -D knobs for the traversal kernels) into pbrt-v4_b200/variants/.

Usage: python scripts/build_variants.py name1:-DA=1,-DB=2 name2:...   then on the GPU box: python scripts/sweep.py
"""
import os, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pbrt-v4_b200"))
import build as B

def main():
    B.build()
    objdir = os.path.join(B.HERE, "build")
    vdir = B.HERE   # next to data/ (the library finds its tables relative to itself)
    os.makedirs(vdir, exist_ok=True)
    procs = []
    for spec in sys.argv[1:]:
        name, _, defs = spec.partition(":")
        name, _, src = name.partition("@")          # name@source.cu:-D... (default source: wavefront.cu)
        src = src or "wavefront.cu"
        defs = [d for d in defs.split(",") if d]
        o = os.path.join(objdir, f"variant_{name}.o")
        cmd = [B.NVCC] + B.FLAGS + defs + ["-c", os.path.join(B.CSRC, src), "-o", o]
        procs.append((name, o, src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for name, o, src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out.decode()); raise SystemExit(1)
        objs = [o] + [os.path.join(objdir, s + ".o") for s in B.SOURCES if s != src]
        so = os.path.join(vdir, f"libb200pt_{name}.so")
        subprocess.check_call([B.NVCC, "-shared", "-o", so] + objs + ["-ldl"])
        print(so)

if __name__ == "__main__":
    main()
