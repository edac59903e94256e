#!/bin/bash
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== ncu full syrk2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk2_lower_dmma_kernel -s 1 -c 1 -o $OUT/syrk2_full -f \
    python scripts/prof_stats.py > $OUT/ncu_syrk2.log 2>&1
tail -2 $OUT/ncu_syrk2.log
echo "== ncu full syrk gen1"
OAK_SYRK_GEN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:syrk_lower_dmma_kernel -s 1 -c 1 -o $OUT/syrk1_full -f \
    python scripts/prof_stats.py > $OUT/ncu_syrk1.log 2>&1
tail -2 $OUT/ncu_syrk1.log
ls -la $OUT
