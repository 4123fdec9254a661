#!/bin/bash
# round 2, GPU call H: outlined BxDFs, instance entry/exit gate variants, emissive media / filters / EXR tests, staging bound
O=gpurun_out/r02h; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
BPT_WORKLOAD=instanced20k SCENE_SPP=64 SPP=64 timeout 900 python scripts/sweep.py > $O/sweep_instanced20k_variants.log 2>&1
# correctness of the gate variants: the instanced parity scenes against the live reference with each variant library
for v in ig8 ig16 ig24; do
  BPT_LIB=$PWD/pbrt-v4_b200/libb200pt_$v.so timeout 300 python scripts/parity_table.py $O/parity_$v.json instanced1k_diffuse_256_16spp instanced1k_coated_256_16spp coated50_64_64spp > $O/parity_$v.log 2>&1
done
SWEEP="X=0" BPT_WORKLOAD=mixed10m SCENE_SPP=16 SPP=16 timeout 600 python scripts/sweep_env.py > $O/sweep_mixed10m.log 2>&1
BPT_LIB=$PWD/pbrt-v4_b200/libb200pt_ob0.so SWEEP="X=0" BPT_WORKLOAD=mixed10m SCENE_SPP=16 SPP=16 timeout 600 python scripts/sweep_env.py > $O/sweep_mixed10m_ob0.log 2>&1
SWEEP="X=0" BPT_WORKLOAD=smoke64 SCENE_SPP=256 SPP=64 timeout 400 python scripts/sweep_env.py > $O/sweep_smoke64.log 2>&1
timeout 600 python scripts/staging_bound.py r02h/staging_bound.json > $O/staging_bound.log 2>&1
ls -la $O
