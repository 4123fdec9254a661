set -x
mkdir -p gpurun_out
N=708 RES=256 SPP=16 python scripts/parity_scene.py heightfield > gpurun_out/par_hf.log 2>&1; tail -5 gpurun_out/par_hf.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01c.json 2> gpurun_out/bench_r01c.err; cat gpurun_out/bench_r01c.json
SPP=4 ncu --set full --import-source on --clock-control none -k regex:"k_generate_camera_rays|k_shade" -s 2 -c 4 -o gpurun_out/r01c_camera_shade -f python scripts/ncu_target.py > gpurun_out/ncu_cs.log 2>&1; tail -3 gpurun_out/ncu_cs.log
SPP=4 ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 2 -c 2 -o gpurun_out/r01c_closest -f python scripts/ncu_target.py > gpurun_out/ncu_cl.log 2>&1; tail -3 gpurun_out/ncu_cl.log
