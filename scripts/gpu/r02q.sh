#!/bin/bash
# round 2, GPU call Q (1 GPU): image textures / bump mapping / delta lights against the live reference, then the whole GPU suite and
# a bench line (did splitting k_shade into two translation units or the outlined delta-light sampling move the headline number?)
O=gpurun_out/r02q; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -k "texture or lights or delta or binding" > $O/pytest_new.log 2>&1; echo "new tests rc=$?" >> $O/pytest_new.log
tail -n 25 $O/pytest_new.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -n 4 $O/pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_hf1m.json 2> $O/bench_hf1m.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02q/bench_hf1m.json")); k = d["roofline"]["kernels"]
print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), {a: round(b["ms_total"], 1) for a, b in k.items() if isinstance(b, dict)})
PY
