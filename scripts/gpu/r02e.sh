#!/bin/bash
# round 2, GPU call E: numeric RenderFromObject inverse (instancing parity), GBuffer maxdepth 0, shade occupancy 3/4/5
O=gpurun_out/r02e; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 400 python scripts/instancing_bisect.py r02e/inst_bisect.json > $O/inst_bisect.log 2>&1
timeout 400 python scripts/instancing_gbuffer.py r02e/inst_gbuffer.json > $O/inst_gbuffer.log 2>&1
timeout 600 python scripts/parity_table.py $O/parity_table.json > $O/parity.log 2>&1
SPP=64 timeout 600 python scripts/sweep.py > $O/sweep_shade_ctas.log 2>&1
ls -la $O
