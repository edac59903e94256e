#!/bin/bash
# A/B of backward-tile build variants on one box.  usage: gpurun --timeout 900 -- 'bash scripts/gpu_ab_bw.sh tag v1 v2 ...'
TAG=${1:-abbw}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 250 python scripts/ab_backward.py 2>&1 | tail -1 | tee $OUT/ab.txt
for v in "$@"; do
  OAK_B200_LIB=$PWD/scripts/ubench/liboak_$v.so timeout 250 python scripts/ab_backward.py 2>&1 | tail -1 | tee -a $OUT/ab.txt
done
timeout 250 python scripts/ab_backward.py 2>&1 | tail -1 | tee -a $OUT/ab.txt
timeout 200 python -m pytest tests/test_gpu_training.py -q -x 2>&1 | tail -2 | tee -a $OUT/ab.txt
for v in "$@"; do
  OAK_B200_LIB=$PWD/scripts/ubench/liboak_$v.so timeout 200 python -m pytest tests/test_gpu_training.py -q -x 2>&1 | tail -2 | tee -a $OUT/ab.txt
done
