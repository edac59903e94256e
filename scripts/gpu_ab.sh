#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
{
scripts/ubench/dmma_dfma
for m in 1 0 9 20 21 22; do OAK_SYRK_MODE=$m python scripts/quick_sgpr.py; done
} 2>&1 | grep -v Warning | tee $OUT/sgpr.txt
