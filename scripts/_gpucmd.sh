mkdir -p gpurun_out
cp pbrt-v4_b200/libb200pt.so gpurun_out/r01o_libb200pt.so
BPT_WORKLOAD=instanced20k SPP=16 ncu --set full --import-source on --clock-control none -k regex:"k_trace_closest|k_shade" -c 2 -o gpurun_out/r01o_instanced_closest_shade -f python scripts/ncu_target.py > gpurun_out/ncu_i.log 2>&1; tail -2 gpurun_out/ncu_i.log
