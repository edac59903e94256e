#!/bin/bash
TAG=${1:-r02e}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== bench"; timeout 1200 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json | cut -c1-3000
tail -5 $OUT/bench.err
echo "== elbo tail"; timeout 300 python scripts/profile_elbo_tail.py 2>&1 | tail -4 | tee $OUT/elbo_tail.txt
