#!/bin/bash
# round 2, GPU call K: warp-cooperative treelets, moved mesh arrays + mapped PLY (re-commit cost), bench
O=gpurun_out/r02k; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
BPT_TIMING=1 timeout 200 python - > $O/commit_timing.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, '.')
import bench
from pbrt_v4_b200 import engine as E
for wl in ('heightfield1m', 'instanced20k'):
    p = bench.make_scene(wl, spp=64)
    e = E.Engine(0)
    for i in range(3):
        t = time.perf_counter(); e.load_pbrt(p, seed=0); print(wl, 'load_pbrt wall ms', (time.perf_counter() - t) * 1e3, file=sys.stderr)
    e.close()
PY
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2;BPT_TREELET_ROUNDS=3;BPT_TREELET_ROUNDS=4" SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_hf1m.log 2>&1
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2" BPT_WORKLOAD=instanced20k SCENE_SPP=64 SPP=64 timeout 600 python scripts/sweep_env.py > $O/sweep_instanced20k.log 2>&1
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2" BPT_WORKLOAD=mixed10m SCENE_SPP=16 SPP=16 timeout 600 python scripts/sweep_env.py > $O/sweep_mixed10m.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err
ls -la $O
