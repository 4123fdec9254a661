#!/bin/bash
# round 2, GPU call G: where configs 4 and 5 stand with the round-2 binary (tuning sizes), ncu of the two-level traversal + coated shade, builder launch list
O=gpurun_out/r02g; mkdir -p $O
SWEEP="X=0" BPT_WORKLOAD=instanced20k SCENE_SPP=64 SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_instanced20k.log 2>&1
SWEEP="X=0" BPT_WORKLOAD=smoke64 SCENE_SPP=256 SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_smoke64.log 2>&1
BPT_WORKLOAD=instanced20k SCENE_SPP=16 SPP=16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_instanced20k.csv python scripts/ncu_target.py > $O/ncu_launch_inst.log 2>&1
BPT_WORKLOAD=smoke64 SCENE_SPP=16 SPP=16 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_smoke64.csv python scripts/ncu_target.py > $O/ncu_launch_smoke.log 2>&1
SCENE_SPP=1 SPP=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_build_hf1m.csv python scripts/ncu_target.py > $O/ncu_launch_build.log 2>&1
BPT_WORKLOAD=instanced20k SCENE_SPP=16 SPP=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_closest|k_shade|k_trace_shadow" -c 6 -o $O/inst20k python scripts/ncu_target.py > $O/ncu_inst.log 2>&1
bash scripts/ncu_export.sh $O/inst20k
BPT_WORKLOAD=smoke64 SCENE_SPP=16 SPP=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_trace_shadow_tr|k_sample_medium|k_medium_scatter" -c 3 -o $O/smoke64 python scripts/ncu_target.py > $O/ncu_smoke.log 2>&1
bash scripts/ncu_export.sh $O/smoke64
ls -la $O
