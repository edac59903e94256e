#!/bin/bash
# compute-sanitizer on small invocations of every kernel family (scripts/sanitize_small.py)
TAG=${1:-san}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== memcheck"; timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > $OUT/memcheck.log 2>&1; tail -4 $OUT/memcheck.log
echo "== racecheck (default build: 3-stage cp.async pipelines)"; SAN_WHICH=gram,strips,chol,pgemm,sgpr timeout 1500 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
echo "== racecheck (2-stage build: every cp.async wait is wait_group 0)"; OAK_B200_LIB=$PWD/scripts/ubench/liboak_st2.so SAN_WHICH=gram,strips,chol,pgemm,sgpr timeout 1500 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > $OUT/racecheck_2stage.log 2>&1; tail -3 $OUT/racecheck_2stage.log
echo "== racecheck backward"; SAN_WHICH=backward timeout 1500 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > $OUT/racecheck_backward.log 2>&1; tail -3 $OUT/racecheck_backward.log
echo "== synccheck"; timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_small.py > $OUT/synccheck.log 2>&1; tail -3 $OUT/synccheck.log
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_sgpr_routes.py tests/test_gpu_models.py -q -x 2>&1 | tail -3
