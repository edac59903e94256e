#!/bin/bash
# overlapped factorisation: parity + timing at the per-rank sizes of 8 and 1 GPUs
OUT=gpurun_out/r02p; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_models.py -m gpu -x -q 2>&1 | tail -5 | tee $OUT/pytest.txt
AB_N=125000 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_125k.txt
AB_N=250000 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_250k.txt
AB_N=1000000 AB_OVERLAPS=0,8 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_1m.txt
