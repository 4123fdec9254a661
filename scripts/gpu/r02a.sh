#!/bin/bash
# round 2, GPU call A: parity on the shipped binary, BVH builder check, bench, ncu of the shipped binary (config 2 and mixed10m)
O=gpurun_out/r02a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python scripts/parity_table.py $O/parity_table.json > $O/parity.log 2>&1
timeout 300 python scripts/bvh_invariance.py r02a/bvh_invariance.json > $O/bvh_invariance.log 2>&1
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2;BPT_TREELET_ROUNDS=3;BPT_TREELET_ROUNDS=2,BPT_LEAF_MAX=2;BPT_TREELET_ROUNDS=2,BPT_LEAF_MAX=1" SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_hf1m.log 2>&1
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err
# ncu: launch list + full capture, config 2 (8 Mi paths per pass so that ncu's save/restore stays small)
SCENE_SPP=8 SPP=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_hf1m.csv python scripts/ncu_target.py > $O/ncu_launch.log 2>&1
SCENE_SPP=8 SPP=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_generate_camera|k_trace_shadow" -c 8 -o $O/hf1m python scripts/ncu_target.py > $O/ncu_hf1m.log 2>&1
bash scripts/ncu_export.sh $O/hf1m
# mixed10m
SWEEP="BPT_TREELET_ROUNDS=0;BPT_TREELET_ROUNDS=2" BPT_WORKLOAD=mixed10m SCENE_SPP=16 SPP=16 timeout 600 python scripts/sweep_env.py > $O/sweep_mixed10m.log 2>&1
BPT_WORKLOAD=mixed10m SCENE_SPP=4 SPP=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest" -c 3 -o $O/mixed10m python scripts/ncu_target.py > $O/ncu_mixed10m.log 2>&1
bash scripts/ncu_export.sh $O/mixed10m
ls -la $O
