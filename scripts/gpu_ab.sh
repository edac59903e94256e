#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
U=$PWD/scripts/ubench
{
OAK_SYRK_MODE=30 python scripts/quick_sgpr.py
for v in s16x3 s16x6 s32x3; do echo $v; OAK_B200_LIB=$U/liboak_$v.so OAK_SYRK_MODE=30 python scripts/quick_sgpr.py; done
} 2>&1 | grep -v Warning | tee $OUT/sgpr.txt
