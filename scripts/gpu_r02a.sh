#!/bin/bash
# Round 2, first GPU call: new factor-first SGPR path (bordered Cholesky, panel GEMM, routes), whole GPU suite,
# bench, ncu launch list of the ELBO part.  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_r02a.sh'
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.txt 2>&1
echo "== new tests"; timeout 600 python -m pytest tests/test_gpu_sgpr_routes.py -q 2>&1 | tail -40 | tee $OUT/pytest_routes.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 900 python bench.py --no-cpu 2> $OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
echo "== ncu launch list (ELBO part)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_elbo.csv \
    python scripts/profile_elbo_tail.py > $OUT/elbo_under_ncu.log 2>&1
tail -3 $OUT/elbo_under_ncu.log
ls -la $OUT
