#!/bin/bash
# Cholesky phase timing + the tests that exercise it.
TAG=${1:-r02b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== chol timing"; OAK_B200_LIB=$PWD/scripts/ubench/liboak_cholt.so timeout 300 python scripts/chol_timing.py 2>&1 | tee $OUT/chol_timing.txt
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_training.py tests/test_gpu_models.py -q 2>&1 | tail -8 | tee $OUT/pytest_new.txt
echo "== elbo tail"; timeout 300 python scripts/profile_elbo_tail.py 2>&1 | tail -4 | tee $OUT/elbo_tail.txt
