#!/bin/bash
# round 2, GPU call V (1 GPU): the whole GPU suite after the stateless-sampler change (texture mappings new), bench of the shipped binary,
# then its ncu launch list and one --set full capture of the four top kernels (profiles/r02v_*: what bench.py's roofline.traffic reads)
O=gpurun_out/r02v; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -n 25 $O/pytest.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02v/bench_hf1m.json")); k = d["roofline"]["kernels"]
print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d["cpu_baseline"] and round(d["cpu_baseline"]["value"], 2), {a: round(b["ms_total"], 1) for a, b in k.items() if isinstance(b, dict)})
PY
SCENE_SPP=8 SPP=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02v_launches_hf1m.csv python scripts/ncu_target.py > $O/ncu_launch.log 2>&1
SCENE_SPP=8 SPP=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_generate_camera|k_trace_shadow" -c 8 -o $O/r02v_hf1m python scripts/ncu_target.py > $O/ncu_hf1m.log 2>&1
bash scripts/ncu_export.sh $O/r02v_hf1m
ls -la $O | tail -12
