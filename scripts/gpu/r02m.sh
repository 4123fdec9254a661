#!/bin/bash
# round 2, GPU call M (8 GPUs): scaling of the headline workload and the 8-GPU BASELINE configurations
O=gpurun_out/r02m; mkdir -p $O
run() {  # name nproc workload steps warmup extra
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 295$2$2 bench.py --gpus $2 --workload $3 --steps $4 --warmup $5 $6 > $O/$1.json 2> $O/$1.err
  echo "$1 rc=$?" >> $O/status.log
}
run bench_hf1m_n8 8 heightfield1m 10 3 ""
run bench_hf1m_n4 4 heightfield1m 10 3 "--no-cpu-baseline"
run bench_mixed10m_n8 8 mixed10m 2 1 "--no-cpu-baseline"
run bench_instanced100k_n8 8 instanced100k 2 1 "--no-cpu-baseline"
cat $O/status.log; tail -n 3 $O/bench_mixed10m_n8.err
