#!/bin/bash
# four-pivot potf2: parity, phase timing, effect on the ELBO tail
OUT=gpurun_out/r02ag; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest.txt
OAK_B200_LIB=$PWD/scripts/ubench/liboak_cholt.so timeout 300 python scripts/chol_timing.py 2>&1 | tee $OUT/chol_timing.txt
AB_N=125000 AB_OVERLAPS=0,4,6,8,12 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_125k.txt
