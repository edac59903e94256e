#!/bin/bash
# gpu_quick.sh plus the ncu launch list of the bench command.  usage: gpurun --timeout 900 -- 'bash scripts/gpu_final.sh [tag]'
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 300 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json
echo "== ncu launch list"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/bench_under_ncu.log 2>&1
wc -l $OUT/launches.csv
