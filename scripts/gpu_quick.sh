#!/bin/bash
# One short gpurun call: GPU parity tests (all failures listed), smoke, default bench.  No profiler passes.
# usage: gpurun --timeout 700 -- 'bash scripts/gpu_quick.sh [tag]'
TAG=${1:-quick}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -40 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 300 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
