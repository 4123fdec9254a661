#!/bin/bash
# round 2, GPU call J (2 GPUs): multi-GPU path with the tile gather; film checksum must equal the 1-GPU one
O=gpurun_out/r02j; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_hf1m_n2.json 2> $O/bench_hf1m_n2.err
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_hf1m_n1.json 2> $O/bench_hf1m_n1.err
tail -3 $O/*.err
