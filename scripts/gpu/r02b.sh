#!/bin/bash
# round 2, GPU call B: outlined libm / sampler (code-size) variants, tie audit, commit-time breakdown
O=gpurun_out/r02b; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/parity_table.py $O/parity_table.json cornell_256_64spp heightfield1m_256_16spp smoke16_64_64spp mixed257k_256_16spp > $O/parity.log 2>&1
SPP=64 timeout 600 python scripts/sweep.py > $O/sweep_variants_hf1m.log 2>&1
timeout 300 python scripts/tie_audit.py r02b/tie_audit.json > $O/tie_audit.log 2>&1
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
SCENE_SPP=8 SPP=8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_shade|k_generate_camera" -c 3 -o $O/hf1m_outlined python scripts/ncu_target.py > $O/ncu.log 2>&1
bash scripts/ncu_export.sh $O/hf1m_outlined
ls -la $O
