#!/bin/bash
# round 2, GPU call Y (8 GPUs): the multi-GPU path on the shipped binary -- config 2 over eight ranks, film checksum against N = 1
O=gpurun_out/r02y; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_hf1m_n8.json 2> $O/bench_hf1m_n8.err; echo "n2 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02y/bench_hf1m_n8.json"))
print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d.get("film_checksum"), d["n_gpus"])
PY
tail -3 $O/bench_hf1m_n8.err
