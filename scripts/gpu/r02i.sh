#!/bin/bash
# round 2, GPU call I: cooperative treelet kernel (shared-memory DP), tight instance bounds, bench
O=gpurun_out/r02i; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/bvh_invariance.py r02i/bvh_invariance.json > $O/bvh_invariance.log 2>&1
SWEEP="X=0;BPT_LOOSE_INSTANCE_BOUNDS=1" BPT_WORKLOAD=instanced20k SCENE_SPP=64 SPP=64 timeout 600 python scripts/sweep_env.py > $O/sweep_instanced20k.log 2>&1
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2;BPT_TREELET_ROUNDS=3" SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_hf1m.log 2>&1
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2" BPT_WORKLOAD=mixed10m SCENE_SPP=16 SPP=16 timeout 600 python scripts/sweep_env.py > $O/sweep_mixed10m.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err
ls -la $O
