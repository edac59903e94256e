#!/bin/bash
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for c in D A B; do AB_CFG=$c timeout 300 python scripts/prof_gram_configs.py 2>&1 | tail -1 | tee -a $OUT/gram_configs.txt; done
M=gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,dram__bytes_write.sum,dram__bytes_read.sum
for c in D A B; do
  AB_CFG=$c timeout 600 ncu --metrics $M --clock-control none -k regex:gram_kernel -s 2 -c 1 --csv --log-file $OUT/gram_${c}_metrics.csv python scripts/prof_gram_configs.py > /dev/null 2>&1
  echo "--- $c"; cut -d, -f13- $OUT/gram_${c}_metrics.csv | tail -17
done
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_gram.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -q -x 2>&1 | tail -3 | tee $OUT/pytest.txt
