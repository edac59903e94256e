#!/bin/bash
# bench at N GPUs under torchrun, as the driver launches it (both arms).  usage: gpurun --gpus N -- 'bash scripts/gpu_bench_n.sh N tag'
N=${1:-2}; OUT=gpurun_out/${2:-bench_n$N}; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 5 --warmup 3 2> $OUT/bench.err | tail -1 > $OUT/bench_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 2>> $OUT/bench.err | tail -1 > $OUT/bench_reference_n$N.json
python - <<PY
import json
d = json.loads(open("$OUT/bench_n$N.json").read()); e = d["sgpr_elbo"]
print("gram", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print({k: e[k] for k in ("value", "ms_per_eval", "ms_stats_phase", "ms_factor_and_stats_overlapped", "ms_finish_tail", "ms_allreduce", "ms_per_eval_by_rank")})
print(e.get("training_step"))
r = json.loads(open("$OUT/bench_reference_n$N.json").read()); print("reference arm", r.get("value"), r.get("cpu_baseline", {}).get("cores"))
PY
