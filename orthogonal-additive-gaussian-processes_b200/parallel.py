"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in the CPU tests).

The path shards without a data-path collective for the Gram (row blocks are independent) and with
exactly one exchange for the SGPR bound: the packed statistics ``Phi | Kuf y | sum K_diag | y^T y``
(M^2 + M + 2 doubles) are summed across ranks.  The reference has no distributed code at all
(SURVEY.md section 5); this is the B200-side design of SURVEY.md section 8(e).
"""
from __future__ import annotations

from typing import List, Tuple

TILE = 64  # row ranges handed to oak_gram_f64 must start on a tile boundary


def is_distributed() -> bool:
    try:
        import torch.distributed as dist

        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    except Exception:  # pragma: no cover
        return False


def rank_world() -> Tuple[int, int]:
    if is_distributed():
        import torch.distributed as dist

        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_rows(n: int, world: int, tile: int = TILE) -> List[Tuple[int, int]]:
    """Contiguous, tile-aligned row ranges [begin, end) covering [0, n): whole tiles are dealt
    out as evenly as possible, the remainder tiles going to the lowest ranks."""
    if n < 0 or world < 1:
        raise ValueError("partition_rows: need n >= 0 and world >= 1")
    tiles = (n + tile - 1) // tile
    base, extra = divmod(tiles, world)
    out, start = [], 0
    for r in range(world):
        cnt = base + (1 if r < extra else 0)
        b = min(start * tile, n)
        e = min((start + cnt) * tile, n)
        out.append((b, e))
        start += cnt
    return out


def balanced_symmetric_rows(n: int, world: int, tile: int = TILE) -> List[List[Tuple[int, int]]]:
    """Folded assignment for the symmetric Gram: the row blocks are cut into 2*world strips and
    rank r owns strips r and 2*world-1-r, which balances the work of the lower triangle
    (strip s costs ~ (s + 1/2) strip-areas)."""
    strips = partition_rows(n, 2 * world, tile)
    return [[strips[r], strips[2 * world - 1 - r]] for r in range(world)]


def allreduce_sum_(t):
    """In-place sum over ranks of the packed statistics vector (CPU/gloo or CUDA/NCCL tensor)."""
    import torch.distributed as dist

    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_int(v: int) -> int:
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(v)], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def global_count(model, n_local: int) -> int:
    """Sum of the ranks' local row counts, exchanged ONCE per model and kept on it: the data of a model are fixed at
    construction (so every rank hits or misses this cache together, which a collective needs), and the ``.item()`` of
    an integer all-reduce is a host synchronisation in the middle of every objective evaluation (measured at 8 GPUs:
    0.7 ms of an 8.7 ms ELBO, the statistics phase drained before the tail could be enqueued)."""
    cached = getattr(model, "_global_count", None)
    if cached is None or cached[0] != int(n_local):
        cached = (int(n_local), allreduce_int(n_local))
        model._global_count = cached
    return cached[1]
