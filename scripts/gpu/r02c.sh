#!/bin/bash
# round 2, GPU call C: level-synchronous treelet builder + device triangle assembly (validation, timing), parity bisection, bench
O=gpurun_out/r02c; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/bvh_invariance.py r02c/bvh_invariance.json > $O/bvh_invariance.log 2>&1
timeout 400 python scripts/instancing_bisect.py r02c/inst_bisect.json > $O/inst_bisect.log 2>&1
BPT_TIMING=1 timeout 200 python - > $O/commit_timing.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import bench
from pbrt_v4_b200 import engine as E
p = bench.make_scene('heightfield1m')
e = E.Engine(0)
for i in range(3):
    t = time.perf_counter(); e.load_pbrt(p, seed=0); print('load_pbrt wall ms', (time.perf_counter() - t) * 1e3, file=sys.stderr)
t = time.perf_counter(); img = e.image(); print('image() ms', (time.perf_counter() - t) * 1e3, file=sys.stderr)
PY
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2;BPT_TREELET_ROUNDS=3;BPT_TREELET_ROUNDS=4;BPT_TREELET_ROUNDS=6" SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_hf1m.log 2>&1
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=3" BPT_WORKLOAD=mixed10m SCENE_SPP=16 SPP=16 timeout 600 python scripts/sweep_env.py > $O/sweep_mixed10m.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err
ls -la $O
