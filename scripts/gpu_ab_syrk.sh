#!/bin/bash
# A/B of the SGPR contraction kernel (development aid): correctness of the SGPR tests, then the statistics phase
# at N = 10^6 and at 125 000 (the per-rank size of the 8-GPU run).  OAK_SYRK_PIPE=1 selects the mbarrier variant.
OUT=gpurun_out/${1:-absyrk}
mkdir -p $OUT
for PIPE in ${AB_PIPES:-0}; do
  [ -n "$AB_SKIP_TESTS" ] || {
  echo "== OAK_SYRK_PIPE=$PIPE tests" | tee -a $OUT/ab.txt
  OAK_SYRK_PIPE=$PIPE timeout 300 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3 | tee -a $OUT/ab.txt
  }
  for N in 1000000 125000; do
    echo "== OAK_SYRK_PIPE=$PIPE N=$N" | tee -a $OUT/ab.txt
    OAK_SYRK_PIPE=$PIPE AB_N=$N AB_OVERLAPS=8 timeout 200 python scripts/profile_elbo_tail.py 2>&1 | tail -4 | cut -c1-330 | tee -a $OUT/ab.txt
  done
done
