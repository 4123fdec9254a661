#!/bin/bash
# round 2, GPU call F: FMA-chain matrix product (instancing parity), tightened parity tests, binding vs front end on the GPU, bench
O=gpurun_out/r02f; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 400 python scripts/instancing_bisect.py r02f/inst_bisect.json > $O/inst_bisect.log 2>&1
timeout 600 python scripts/parity_table.py $O/parity_table.json > $O/parity.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err
ls -la $O
