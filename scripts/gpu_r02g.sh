#!/bin/bash
TAG=${1:-r02g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_gram.py -q -x 2>&1 | tail -3 | tee $OUT/pytest.txt
echo "== 1M (Kuf y in the Gram epilogue)"; AB_N=1000000 timeout 300 python scripts/profile_elbo_tail.py 2>&1 | tail -4 | tee $OUT/elbo_tail_1m_fold.txt
echo "== 1M (cuBLAS gemv)"; OAK_NO_FOLD_KY=1 AB_N=1000000 timeout 300 python scripts/profile_elbo_tail.py 2>&1 | tail -4 | tee $OUT/elbo_tail_1m_gemv.txt
echo "== 125k"; timeout 300 python scripts/profile_elbo_tail.py 2>&1 | tail -4 | tee $OUT/elbo_tail_fold.txt
