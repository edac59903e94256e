#!/bin/bash
OUT=gpurun_out/r02ad; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/pytest.txt
AB_N=1000000 AB_OVERLAPS=4 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_1m.txt
AB_N=125000 AB_OVERLAPS=8 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tee $OUT/elbo_tail_125k.txt
