#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
U=$PWD/scripts/ubench
{
python scripts/quick_sgpr.py
for v in m4 m4hf; do echo $v; OAK_B200_LIB=$U/liboak_$v.so python scripts/quick_sgpr.py; done
python scripts/quick_sgpr.py
} 2>&1 | grep -v Warning | tee $OUT/sgpr.txt
