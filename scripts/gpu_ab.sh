#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
{
python scripts/ab_gram.py
AB_N=65536 python scripts/ab_gram.py
} 2>&1 | grep -v Warning | tee $OUT/ab.txt
