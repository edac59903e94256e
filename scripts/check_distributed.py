"""torchrun check (N >= 2 GPUs): the N-sharded SGPR ELBO and its gradient equal the single-GPU ones.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_distributed.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oak_b200 import parallel
from oak_b200.models import SGPR
from oak_b200.training import freeze_unsupported, sgpr_elbo_and_grad
from oak_b200.workloads import build_kernel, config_C

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = config_C(300_000, 20, 512, 3)
b, e = parallel.partition_rows(300_000, world)[rank]
m = SGPR((cfg["X"][b:e], cfg["y"][b:e]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=32768, distributed=True)
m.likelihood.variance.assign(cfg["noise"])
freeze_unsupported(m)  # nothing to freeze: the inducing points are differentiated too
elbo_d = m.elbo()
val_d, gl_d, gv_d, gn_d = sgpr_elbo_and_grad(m)
gz_d = m._inducing_grad
ok = True
if rank == 0:
    s = SGPR((cfg["X"], cfg["y"]), kernel=build_kernel(cfg), inducing_variable=cfg["Z"], chunk=32768, distributed=False)
    s.likelihood.variance.assign(cfg["noise"])
    freeze_unsupported(s)
    elbo_s = s.elbo()
    val_s, gl_s, gv_s, gn_s = sgpr_elbo_and_grad(s)
    rel = lambda a, c: float(np.max(np.abs(np.asarray(a) - np.asarray(c))) / np.max(np.abs(np.asarray(c))))
    print(f"world {world}: elbo {elbo_d:.6f} vs {elbo_s:.6f} (rel {abs(elbo_d-elbo_s)/abs(elbo_s):.2e}); "
          f"grad ls rel {rel(gl_d, gl_s):.2e}, var rel {rel(gv_d, gv_s):.2e}, noise rel {abs(gn_d-gn_s)/abs(gn_s):.2e}, "
          f"Z rel {rel(gz_d, s._inducing_grad):.2e}")
    ok = (abs(elbo_d - elbo_s) < 1e-10 * abs(elbo_s) and rel(gl_d, gl_s) < 1e-9 and rel(gv_d, gv_s) < 1e-9
          and rel(gz_d, s._inducing_grad) < 1e-9)
    print("DISTRIBUTED CHECK", "OK" if ok else "FAILED")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
