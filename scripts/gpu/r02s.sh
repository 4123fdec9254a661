#!/bin/bash
# round 2, GPU call S (1 GPU): orthographic camera, textures behind media (differentials survive a medium scattering event), GPU suite
O=gpurun_out/r02s; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -n 30 $O/pytest.log | cut -c1-300
