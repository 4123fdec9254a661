set -x
mkdir -p gpurun_out
BPT_TIMING=1 python - <<'PY' 2>&1 | tail -24
import sys, time
sys.path.insert(0, '.')
import bench
from pbrt_v4_b200 import engine as E
path = bench.make_scene("heightfield1m", spp=256)
e = E.Engine(0)
for i in range(3):
    t = time.time(); e.load_pbrt(path, seed=0); print("load_pbrt wall", time.time() - t, flush=True)
t = time.time(); e.render(0, 32); print("render 32spp", time.time() - t)
t = time.time(); img = e.image(); print("image()", time.time() - t)
PY
SPP=32 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01h_launches.csv python scripts/ncu_target.py > gpurun_out/ncu_l.log 2>&1; tail -2 gpurun_out/ncu_l.log
SPP=32 ncu --set full --import-source on --clock-control none -k regex:"k_trace_closest|k_shade|k_generate_camera" -c 3 -o gpurun_out/r01h_closest_shade_camera -f python scripts/ncu_target.py > gpurun_out/ncu_f.log 2>&1; tail -2 gpurun_out/ncu_f.log
cp pbrt-v4_b200/libb200pt.so gpurun_out/r01h_libb200pt.so
