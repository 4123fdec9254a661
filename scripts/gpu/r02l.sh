#!/bin/bash
# round 2, GPU call L: alpha-tested primitives, GBufferFilm with media, threaded PLY decode; smoke(); bench
O=gpurun_out/r02l; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_hf1m.json 2> $O/bench_hf1m.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
ls -la $O
