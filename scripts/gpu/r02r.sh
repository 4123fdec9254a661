#!/bin/bash
# round 2, GPU call R (1 GPU): normal maps, textures behind media (flag fix), whole GPU suite, bench with the outlined alpha hash
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -n 30 $O/pytest.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_hf1m.json 2> $O/bench_hf1m.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02r/bench_hf1m.json")); k = d["roofline"]["kernels"]
print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), {a: round(b["ms_total"], 1) for a, b in k.items() if isinstance(b, dict)})
PY
