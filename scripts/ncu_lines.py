"""Joins an ncu report's per-SASS-instruction counters with nvdisasm's line info (inline chains) of the in-tree library and
prints executed warp-instructions / samples aggregated by source location.

usage: python scripts/ncu_lines.py report.ncu-rep kernel_regex [launch_index] [--top N] [--by inner|outer|func] [--mangled RE] [--lib path/to/the/libb200pt.so/the/report/was/taken/with]
The library must be the build the report was taken with (instruction counts are checked).
"""
import collections, csv, os, re, subprocess, sys, tempfile

def disasm(kernel_re, so=None):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = so or os.path.join(root, "pbrt-v4_b200", "libb200pt.so")
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=d, capture_output=True)
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, "wavefront.sm_100a.cubin")], capture_output=True, text=True).stdout
    funcs = {}
    cur = None; chain = []
    pending = []
    for line in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
        if m:
            cur = m.group(1); funcs[cur] = []; pending = []; continue
        if cur is None: continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            pending.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            if pending:
                chain = pending; pending = []
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip(), chain))
    for name, ins in funcs.items():
        if re.search(kernel_re, name) and ins:
            return name, ins
    raise SystemExit("kernel not found in cubin")

def main():
    rep, kre = sys.argv[1], sys.argv[2]
    args = sys.argv[3:]
    launch = int(args[0]) if args and not args[0].startswith("--") else 0
    top = int(args[args.index("--top") + 1]) if "--top" in args else 40
    by = args[args.index("--by") + 1] if "--by" in args else "inner"
    name, ins = disasm(args[args.index("--mangled") + 1] if "--mangled" in args else kre,
                       os.path.abspath(args[args.index("--lib") + 1]) if "--lib" in args else None)
    if rep.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
    else:   # exported on the GPU box: <prefix>.source.csv.gz holds every profiled launch
        import gzip
        out = gzip.open(rep + ".source.csv.gz", "rt").read()
    rows = list(csv.reader(out.splitlines()))
    secs = []; cur = None; hdr = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            keep = rep.endswith(".ncu-rep") or re.search(kre, r[1] if len(r) > 1 else "")
            cur = [] if keep else None
            if keep: secs.append(cur)
            continue
        if r and r[0] == "Address": hdr = r; continue
        if cur is not None and len(r) > 5: cur.append(r)
    sec = secs[launch]
    iex, ismp, ith = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    if len(sec) != len(ins):
        print(f"warning: instruction count differs (report {len(sec)}, library {len(ins)}) -- stale library?")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r, (_, txt, chain) in zip(sec, ins):
        if by == "inner": key = chain[0] if chain else ("?", 0)
        elif by == "outer": key = chain[-1] if chain else ("?", 0)
        else: key = tuple(chain[:3])
        v = (int(r[iex]), int(r[ismp]), int(r[ith]))
        for i in range(3): agg[key][i] += v[i]; tot[i] += v[i]
    print(f"{name}: {tot[0]} warp-inst, {tot[1]} samples, avg lanes {tot[2] / max(tot[0], 1):.1f}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  samples {v[1] / tot[1]:6.3f}  inst {v[0] / tot[0]:6.3f}  lanes {v[2] / max(v[0], 1):5.1f}  {key}")

if __name__ == "__main__":
    main()
