#!/bin/bash
# round 2, GPU call N: new samplers, GBuffer differentials behind interface surfaces; full GPU suite
O=gpurun_out/r02n; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
SWEEP="X=0" SPP=64 timeout 300 python scripts/sweep_env.py > $O/sweep_hf1m.log 2>&1
tail -5 $O/pytest.log
