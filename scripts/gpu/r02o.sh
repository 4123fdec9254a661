#!/bin/bash
# round 2, GPU call O: three more materials, stratified no-jitter diagnosis, full GPU suite
O=gpurun_out/r02o; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python scripts/stratified_diag.py r02o/strat_diag.json > $O/strat_diag.log 2>&1
tail -12 $O/pytest.log
