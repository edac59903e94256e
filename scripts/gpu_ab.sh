#!/bin/bash
# gpurun -- 'bash scripts/gpu_ab.sh tag'   (A/B of Gram-kernel build/launch variants)
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 | tee $OUT/pytest_gpu.txt
{
python scripts/ab_gram.py
OAK_GRAM_NOTMA=1 python scripts/ab_gram.py
OAK_GRAM_VARIANT=1 python scripts/ab_gram.py
} 2>&1 | grep -v Warning | tee $OUT/ab.txt
