#!/bin/bash
# round 2, GPU call P (1 GPU): full GPU suite with the point / spot / distant light tests, then the N=1 lines of the 8-GPU
# configurations (for the scaling table next to profiles/r02m_*)
O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -n 4 $O/pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m_n1.json 2> $O/bench_hf1m_n1.err; echo "hf1m rc=$?"
timeout 400 python bench.py --workload instanced100k --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_instanced100k_n1.json 2> $O/bench_instanced100k_n1.err; echo "inst rc=$?"
timeout 400 python bench.py --workload mixed10m --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_mixed10m_n1.json 2> $O/bench_mixed10m_n1.err; echo "mixed rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02p/bench_*.json")):
    try:
        d = json.load(open(f)); print(f, round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d.get("film_checksum"))
    except Exception as ex: print(f, "unreadable", ex)
PY
