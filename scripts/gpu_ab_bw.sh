#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash scripts/gpu_ab_bw.sh [tag]'
TAG=${1:-abbw}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 250 python scripts/ab_backward.py 2>&1 | tail -3 | tee $OUT/ab.txt
OAK_B200_LIB=$PWD/scripts/ubench/liboak_bwprev.so timeout 250 python scripts/ab_backward.py 2>&1 | tail -3 | tee -a $OUT/ab.txt
