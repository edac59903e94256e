#!/bin/bash
OUT=gpurun_out/r02s; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee $OUT/pytest.txt
AB_N=125000 AB_OVERLAPS=0,6 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_125k.txt
timeout 300 python bench.py --no-cpu --no-e2e 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-300
