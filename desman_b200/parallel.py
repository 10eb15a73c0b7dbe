"""Multi-GPU plumbing: one process per GPU, variant positions sharded in contiguous blocks, one small all-reduce
of the S*G+16 integer statistics (+ ll, nchange) per sweep inside the engine.

torch.distributed (or any other launcher) is only used here to agree on the NCCL unique id; the data path never
touches it.  gamma, eta and the RNG key are replicated: every rank draws the identical gamma/eta from the reduced
statistics, so no broadcast is needed.
"""
import os

import numpy as np


def shard_bounds(V, rank, world):
    """Contiguous block [lo, hi) of the V sites owned by `rank` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(V, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def exchange_unique_id(dist, make_id, nbytes=128):
    """Rank 0 creates the id (make_id() -> bytes), everybody receives it through `dist` (torch.distributed, any
    backend; CPU tensors for gloo, CUDA tensors for nccl)."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == 0:
        t = torch.tensor(list(make_id()), dtype=torch.uint8, device=dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def sharded_sampler(snps_full, G, randomState, dist=None, **kw):
    """HaploSNP_Sampler over this rank's block of `snps_full` (every rank passes the same full tensor, or a tensor
    whose rows [lo,hi) are valid).  Returns (sampler, (lo, hi))."""
    from .engine import Engine
    from .HaploSNP_Sampler import HaploSNP_Sampler
    rank, world, local = env_rank_world()
    V = snps_full.shape[0]
    lo, hi = shard_bounds(V, rank, world)
    comm = None
    if world > 1:
        uid = exchange_unique_id(dist, Engine.comm_unique_id)
        comm = (uid, rank, world)
    hs = HaploSNP_Sampler(np.ascontiguousarray(snps_full[lo:hi]), G, randomState, device=kw.pop("device", local),
                          shard=(lo, V), comm=comm, **kw)
    return hs, (lo, hi)
