#!/bin/bash
# 2-GPU validation of the bench arms (mirror strips, sharded ELBO with all-reduce, symmetric host-buffer strips)
TAG=${1:-r02f}; OUT=gpurun_out/$TAG; mkdir -p $OUT
N=${2:-2}
echo "== bench --gpus $N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 2> $OUT/bench_n$N.err | tee $OUT/bench_n$N.json | cut -c1-1500
tail -5 $OUT/bench_n$N.err
echo "== distributed check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    scripts/check_distributed.py 2>&1 | tail -12 | tee $OUT/check_distributed.txt
