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
# ---- the SVGP classification objective sharded the same way --------------------------------
from oak_b200._gpflow_shim import Bernoulli, inv_logit
from oak_b200.models import SVGP
from oak_b200.training import svgp_elbo_and_grad

rng = np.random.default_rng(11)
nS, dS, mS = 60_000, 6, 96
Xc = rng.standard_normal((nS, dS))
yc = (rng.random((nS, 1)) < 1.0 / (1.0 + np.exp(-np.sin(Xc[:, :1]) - Xc[:, 1:2]))).astype(np.float64)
cfgS = {"dims": [{"type": "rbf", "lengthscale": 1.0 + 0.2 * i, "variance": 1.0, "measure": ("gaussian", 0.0, 1.0)}
                 for i in range(dS)], "depth": 3, "variances": [1.0, 0.8, 0.5, 0.3], "share_var": True}
qm, qs = rng.standard_normal((mS, 1)) * 0.3, rng.uniform(0.5, 1.2, (mS, 1))


def svgp(distributed):
    m_ = SVGP(kernel=build_kernel(cfgS), likelihood=Bernoulli(invlink=inv_logit), inducing_variable=Xc[:mS].copy(),
              whiten=True, q_diag=True, q_mu=qm, q_sqrt=qs, chunk=8192, distributed=distributed)
    return m_


b, e = parallel.partition_rows(nS, world)[rank]
sd = svgp(True)
out_d = svgp_elbo_and_grad(sd, (Xc[b:e], yc[b:e]))
gz_d, gv_d = sd._inducing_grad, sd._variational_grads
if rank == 0:
    ss = svgp(False)
    out_s = svgp_elbo_and_grad(ss, (Xc, yc))
    rel = lambda a, c: float(np.max(np.abs(np.asarray(a) - np.asarray(c))) / np.max(np.abs(np.asarray(c))))
    errs = [abs(out_d[0] - out_s[0]) / abs(out_s[0]), rel(out_d[1], out_s[1]), rel(out_d[2], out_s[2]),
            rel(gz_d, ss._inducing_grad), rel(gv_d[id(sd.q_mu)], ss._variational_grads[id(ss.q_mu)]),
            rel(gv_d[id(sd.q_sqrt)], ss._variational_grads[id(ss.q_sqrt)])]
    print("SVGP world", world, "rel errors (elbo, d/dl, d/dsigma2, d/dZ, d/dq_mu, d/dq_sqrt):", " ".join(f"{x:.1e}" for x in errs))
    ok_s = max(errs) < 1e-9
    print("DISTRIBUTED SVGP CHECK", "OK" if ok_s else "FAILED")
    ok = ok and ok_s
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
