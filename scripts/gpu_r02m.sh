#!/bin/bash
TAG=${1:-r02m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pgemm A/B (BM=64 default)"; timeout 300 python scripts/ab_pgemm.py 2>&1 | tail -2 | tee $OUT/ab_pgemm_bm64.txt
echo "== pgemm A/B (BM=128)"; OAK_B200_LIB=$PWD/scripts/ubench/liboak_pg128.so timeout 300 python scripts/ab_pgemm.py 2>&1 | tail -2 | tee $OUT/ab_pgemm_bm128.txt
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_sgpr_routes.py -q -x 2>&1 | tail -3 | tee $OUT/pytest.txt
echo "== training step"; timeout 600 python scripts/quick_train.py 2>&1 | tail -5 | tee $OUT/quick_train.txt
