#!/bin/bash
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
U=$PWD/scripts/ubench
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
OAK_B200_LIB=$U/liboak_exp8.so python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -1
{
python scripts/ab_gram.py
OAK_B200_LIB=$U/liboak_exp8.so python scripts/ab_gram.py
} 2>&1 | grep -v Warning | tee $OUT/ab.txt
