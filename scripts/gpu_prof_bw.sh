#!/bin/bash
# usage: gpurun --timeout 600 -- 'bash scripts/gpu_prof_bw.sh [tag]'
TAG=${1:-profbw}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 250 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:gram_backward_kernel -s 2 -c 2 --csv --log-file $OUT/backward_metrics.csv \
    python scripts/prof_backward.py > $OUT/log1.txt 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:gram_backward_kernel -s 2 -c 1 -o $OUT/backward_full -f \
    python scripts/prof_backward.py > $OUT/log2.txt 2>&1
tail -2 $OUT/log2.txt
cat $OUT/backward_metrics.csv | tail -20
