#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
U=$PWD/scripts/ubench
for v in exp10r8 exp9r8; do
OAK_B200_LIB=$U/liboak_$v.so python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
done
{
python scripts/ab_gram.py
OAK_B200_LIB=$U/liboak_exp10r8.so python scripts/ab_gram.py
OAK_B200_LIB=$U/liboak_exp9r8.so python scripts/ab_gram.py
} 2>&1 | grep -v Warning | tee $OUT/ab.txt
OAK_B200_LIB=$U/liboak_exp10r8.so timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
