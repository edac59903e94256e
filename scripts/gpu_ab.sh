#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest models"; timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -3
python scripts/ab_gram.py 2>&1 | grep -v Warning | tee $OUT/ab.txt
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/bench.json | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], l['roofline']['frac'], l['clocks']); s=l['extra']['sgpr_elbo']; print(s['value'], s['ms_per_eval'], s['ms_stats_phase'], s['ms_tail_and_collective'])"
