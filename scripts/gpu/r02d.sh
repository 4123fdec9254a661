#!/bin/bash
# round 2, GPU call D: wo normalisation fix (parity), instanced first-hit geometry, shade occupancy variants
O=gpurun_out/r02d; mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 400 python scripts/instancing_bisect.py r02d/inst_bisect.json > $O/inst_bisect.log 2>&1
timeout 400 python scripts/instancing_gbuffer.py r02d/inst_gbuffer.json > $O/inst_gbuffer.log 2>&1
timeout 600 python scripts/parity_table.py $O/parity_table.json > $O/parity.log 2>&1
SPP=64 timeout 600 python scripts/sweep.py > $O/sweep_shade_ctas.log 2>&1
ls -la $O
