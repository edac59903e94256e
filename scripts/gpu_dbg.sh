#!/bin/bash
OUT=gpurun_out/dbg; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_sgpr_routes.py -q -x -k "both_routes" 2>&1 | tail -40 | tee $OUT/pytest.txt
