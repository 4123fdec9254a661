#!/bin/bash
# round 2, GPU call X (2 GPUs): the multi-GPU path on the shipped binary -- config 2 over two ranks, film checksum against N = 1
O=gpurun_out/r02x; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_hf1m_n2.json 2> $O/bench_hf1m_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02x/bench_hf1m_n2.json"))
print(round(d["value"], 1), round(d["ms_per_step"], 1), round(d["e2e"]["value"], 1), d.get("film_checksum"), d["n_gpus"])
PY
tail -3 $O/bench_hf1m_n2.err
