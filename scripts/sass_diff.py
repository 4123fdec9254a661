"""Compares the SASS of the hot kernels of two builds of libb200pt.so (e.g. to show that an off-by-default experiment
leaves the default build's machine code unchanged).  usage: python scripts/sass_diff.py old.so new.so"""
import os, re, subprocess, sys, tempfile


def kernels(so):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(d, "wavefront.sm_100a.cubin")], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur = m.group(1); out[cur] = []; continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur and m:
            out[cur].append(re.sub(r"\.L_x_\d+", ".L", re.sub(r"\s+", " ", m.group(1))))   # label numbers are module-wide counters
    return out


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    for k in sorted(b):
        tag = "SAME" if a.get(k) == b[k] else ("NEW " if k not in a else "DIFF")
        print(tag, k[:90], len(a.get(k, [])), len(b[k]))


if __name__ == "__main__":
    main()
