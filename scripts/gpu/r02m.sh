#!/bin/bash
# round 2, GPU call M (8 GPUs, charged 8x: bounded to ~7 min of box time): scaling of the headline workload, then the 8-GPU
# BASELINE configurations for as long as the time box allows
O=gpurun_out/r02m; mkdir -p $O
T0=$(date +%s)
left() { echo $(( ${BUDGET_S:-400} - ( $(date +%s) - T0 ) )); }
run() {  # name nproc workload steps warmup extra
  L=$(left); if [ $L -lt 60 ]; then echo "$1 skipped (time box)" >> $O/status.log; return; fi
  timeout $L python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 295$2$2 bench.py --gpus $2 --workload $3 --steps $4 --warmup $5 $6 > $O/$1.json 2> $O/$1.err
  echo "$1 rc=$? t=$(( $(date +%s) - T0 ))s" >> $O/status.log
}
run bench_hf1m_n8 8 heightfield1m 10 3 "--no-cpu-baseline"
run bench_hf1m_n4 4 heightfield1m 10 3 "--no-cpu-baseline"
run bench_instanced100k_n8 8 instanced100k 2 1 "--no-cpu-baseline"
run bench_mixed10m_n8 8 mixed10m 2 1 "--no-cpu-baseline"
cat $O/status.log; cat $O/bench_hf1m_n8.json
