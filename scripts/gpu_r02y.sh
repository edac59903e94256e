#!/bin/bash
OUT=gpurun_out/r02y; mkdir -p $OUT
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 scripts/ubench/d2h_bw.py 2>/dev/null | tee $OUT/d2h_bw.txt
NO_BIND=1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 scripts/ubench/d2h_bw.py 2>/dev/null | tee -a $OUT/d2h_bw.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --no-e2e 2> $OUT/bench8.err | tail -1 > $OUT/bench_n8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y/bench_n8.json').read())
e=d['sgpr_elbo']; print({k:e[k] for k in ['value','ms_per_eval','ms_stats_phase','ms_factor_and_stats_overlapped','ms_finish_tail','ms_allreduce']})
PY
