#!/bin/bash
# round 2, GPU call W (1 GPU): N = 1 bench lines of the other BASELINE configurations on the shipped binary (configs 3, 4 and the
# tuning-size smoke of config 5), to check that the stateless-sampler / two-kernel-set changes cost them nothing
O=gpurun_out/r02w; mkdir -p $O
timeout 400 python bench.py --workload instanced100k --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_instanced100k_n1.json 2> $O/bench_instanced100k_n1.err; echo "inst rc=$?"
timeout 400 python bench.py --workload mixed10m --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_mixed10m_n1.json 2> $O/bench_mixed10m_n1.err; echo "mixed rc=$?"
timeout 300 python bench.py --workload smoke64 --steps 3 --warmup 1 --no-cpu-baseline > $O/bench_smoke64_n1.json 2> $O/bench_smoke64_n1.err; echo "smoke rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02w/bench_*.json")):
    try:
        d = json.load(open(f)); k = d["roofline"]["kernels"]
        print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d.get("film_checksum"), {a: round(b["ms_total"], 1) for a, b in k.items() if isinstance(b, dict)})
    except Exception as ex:
        print(f, "unreadable", ex)
PY
