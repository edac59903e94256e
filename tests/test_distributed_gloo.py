"""World-size-2 gloo test of the N > 1 host path (CPU): rows are partitioned, each rank builds the
packed SGPR statistics of its shard (oracle arithmetic stands in for the CUDA tiles here), one
all-reduce combines them, and the bound computed from the reduced statistics equals the unsharded
one.  Exercises oak_b200.parallel exactly as models.SGPR / bench.py use it."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stats(kern, X, y, Z):
    """Phi | Kuf y | sum K_diag | y^T y -- the layout of oak_sgpr_stats_f64."""
    kuf = kern.K(Z, X)
    return np.concatenate([(kuf @ kuf.T).reshape(-1), (kuf @ y).reshape(-1), [kern.K_diag(X).sum()], [float(y.T @ y)]])


def _elbo_from_stats(kern, stats, Z, n_total, noise, jitter=1e-6):
    import scipy.linalg as sla

    m = Z.shape[0]
    phi, kufy, skd, yty = stats[: m * m].reshape(m, m), stats[m * m: m * m + m].reshape(-1, 1), stats[-2], stats[-1]
    L = np.linalg.cholesky(kern.K(Z) + jitter * np.eye(m))
    AAT = sla.solve_triangular(L, sla.solve_triangular(L, phi, lower=True).T, lower=True).T / noise
    LB = np.linalg.cholesky(AAT + np.eye(m))
    c = sla.solve_triangular(LB, sla.solve_triangular(L, kufy, lower=True) / np.sqrt(noise), lower=True) / np.sqrt(noise)
    return (-0.5 * n_total * np.log(2 * np.pi) - np.sum(np.log(np.diag(LB))) - 0.5 * n_total * np.log(noise)
            - 0.5 * yty / noise + 0.5 * float(c.T @ c) - 0.5 * skd / noise + 0.5 * np.trace(AAT))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import build_oracle, mixed_config
    from oak_b200 import parallel
    from oracle import oak_oracle as oo

    cfg = mixed_config(n=300, seed=5, depth=2)
    kern = build_oracle(cfg)
    assert parallel.is_distributed() and parallel.rank_world() == (rank, world)
    b, e = parallel.partition_rows(300, world)[rank]
    stats = torch.as_tensor(_stats(kern, cfg["X"][b:e], cfg["y"][b:e], cfg["Z"]))
    parallel.allreduce_sum_(stats)
    n_total = parallel.allreduce_int(e - b)
    elbo = _elbo_from_stats(kern, stats.numpy(), cfg["Z"], n_total, cfg["noise"])
    ref = oo.sgpr_elbo(kern, cfg["X"], cfg["y"], cfg["Z"], cfg["noise"])
    full = _stats(kern, cfg["X"], cfg["y"], cfg["Z"])
    out[rank] = (n_total, float(elbo), float(ref), float(np.max(np.abs(stats.numpy() - full)) / np.max(np.abs(full))))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_statistics_allreduce_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert sorted(res) == [0, 1]
    for r in range(world):
        n_total, elbo, ref, err = res[r]
        assert n_total == 300
        assert err < 1e-13
        assert abs(elbo - ref) < 1e-9 * abs(ref)
    assert res[0][1] == res[1][1]  # every rank computes the identical bound from the reduced statistics
