set -x
mkdir -p gpurun_out
RES=128 SPP=16 python scripts/parity_scene.py cornell 2>&1 | tail -4
HOMOG=1 RES=128 SPP=16 timeout 300 python scripts/parity_scene.py smoke 2>&1 | tail -5
GRID=32 RES=128 SPP=16 timeout 300 python scripts/parity_scene.py smoke 2>&1 | tail -5
