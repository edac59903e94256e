"""Host-link check for the 8-GPU host-buffer leg (development aid): device-to-host bandwidth of every rank alone and
of all ranks together, contiguous 1 GiB copies into pinned memory first-touched on the GPU's NUMA node.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/ubench/d2h_bw.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import bench
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
bound = bench.bind_to_gpu_numa_node(local) if os.environ.get("NO_BIND") is None else 0
n = 1 << 27  # 1 GiB of doubles
dev = torch.empty(n, dtype=torch.float64, device="cuda").normal_()
host = torch.empty(n, dtype=torch.float64).pin_memory(); host.zero_()
def bw(reps=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize(); return reps * n * 8 / (time.perf_counter() - t0) / 1e9
bw(1)
solo = None
for r in range(world):
    if world > 1: dist.barrier()
    if r == rank: solo = bw()
if world > 1: dist.barrier()
together = bw()
out = torch.tensor([solo, together], device="cuda")
if world > 1:
    allv = [torch.zeros(2, device="cuda") for _ in range(world)]; dist.all_gather(allv, out)
else:
    allv = [out]
if rank == 0:
    print(f"cpus bound {bound}; D2H GB/s per rank alone:   ", " ".join(f"{v[0].item():5.1f}" for v in allv))
    print(f"                D2H GB/s per rank together:", " ".join(f"{v[1].item():5.1f}" for v in allv), f"  sum {sum(v[1].item() for v in allv):.1f}")
if world > 1: dist.destroy_process_group()
