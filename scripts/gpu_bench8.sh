#!/bin/bash
# 8-GPU bench (both the headline arm and a per-rank ELBO breakdown)
OUT=gpurun_out/${1:-r02t}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/nvidia_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu ${BENCH8_EXTRA} 2> $OUT/bench8.err | tee $OUT/bench_n8.json | cut -c1-400
tail -3 $OUT/bench8.err
