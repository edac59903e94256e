#!/bin/bash
OUT=gpurun_out/r02z; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_models.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
AB_N=125000 AB_OVERLAPS=0,4,8,12 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_125k.txt
AB_N=1000000 AB_OVERLAPS=0,4,8 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_1m.txt
