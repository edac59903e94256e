#!/bin/bash
# compute-sanitizer on the kernel families added late in round 2 (overlapped SGPR front, k-means, unique, row-tiled prologue)
TAG=${1:-san3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
export SAN_WHICH=overlap,kmeans,unique
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > $OUT/memcheck.log 2>&1; tail -4 $OUT/memcheck.log
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python scripts/sanitize_small.py > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_small.py > $OUT/synccheck.log 2>&1; tail -3 $OUT/synccheck.log
