#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list of the bench command and one
# `--set full` capture of the dominant kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand).
# usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/nvidia_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench"; timeout 900 python bench.py 2> $OUT/bench.err | tee $OUT/bench.json
tail -5 $OUT/bench.err
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> $OUT/bench.err | tee $OUT/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $OUT/bench_under_ncu.log 2>&1
echo "== ncu dram bytes at the bench size (one pass)"
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg \
    --clock-control none -k regex:gram_kernel -s 2 -c 1 --csv --log-file $OUT/gram_dram_n65536.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-elbo --no-sweep > /dev/null 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 2 -c 1 -o $OUT/gram_full_n32768 -f \
    python bench.py --n 32768 --steps 1 --warmup 1 --no-cpu --no-e2e --no-elbo --no-sweep > $OUT/ncu_full.log 2>&1
tail -3 $OUT/ncu_full.log
ls -la $OUT
