mkdir -p gpurun_out
SECONDS=0
for w in instanced100k mixed10m smoke256; do
  timeout 700 python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; tail -2 gpurun_out/bench_$w.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$w.json"))
    print({k:d[k] for k in ("value","ms_per_step","msamples_per_s")}, d["e2e"]["seconds_per_step"], d["roofline"]["frac"], d["roofline"]["step"]["frac"], {k:(v.get("ms_total"),v.get("nodes_per_ray"),v.get("tris_per_ray")) for k,v in d["roofline"]["kernels"].items() if isinstance(v,dict)}, d["config"]["bvh"])
except Exception as e: print("no json", e)
PY
  echo "elapsed $SECONDS s"
done
