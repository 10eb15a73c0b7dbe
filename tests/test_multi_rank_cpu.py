"""world_size-2 `gloo` tests on CPU of the N>1 host logic: shard bounds, unique-id exchange, and the sharding
contract itself (counters keyed by the GLOBAL site index, integer statistics summed over ranks) exercised with the
oracle standing in for the per-rank device work."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds_cover_and_balance():
    from desman_b200.parallel import shard_bounds
    for V in (1, 7, 100, 100003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(V, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == V
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from conftest import onehot, synth_problem
    from desman_b200.parallel import exchange_unique_id, shard_bounds
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = exchange_unique_id(dist, lambda: bytes(range(128)))          # stand-in for Engine.comm_unique_id
    assert uid == bytes(range(128))
    p = synth_problem(101, 12, 4, depth=20.0, seed=5)                  # same problem on every rank
    lo, hi = shard_bounds(101, rank, world)
    tau = onehot(p["tau0"][lo:hi])
    gamma, eta = p["gamma0"].copy(), p["eta0"].copy()
    seed = 77
    for sweep in range(3):
        sm, es = oracle.mu_stats(tau, gamma, eta, p["counts"][lo:hi], seed, sweep, v0=lo)
        t = torch.from_numpy(np.concatenate([sm.ravel(), es.ravel()]))
        dist.all_reduce(t)                                             # the one exchange step of a sweep
        sm = t[:sm.size].numpy().reshape(sm.shape)
        es = t[sm.size:].numpy().reshape(4, 4)
        gamma = oracle.draw_gamma(sm, 0.1, 1e-6, seed, sweep)          # replicated: same statistics, same key
        oracle.sample_tau_philox(tau, gamma, eta, p["counts"][lo:hi], seed, sweep, v0=lo)
        eta = oracle.draw_eta(es, 0.1, seed, sweep)
    g_all = [torch.zeros(gamma.size, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(g_all, torch.from_numpy(gamma.ravel().copy()))
    assert all(torch.equal(g_all[0], g) for g in g_all)                # identical gamma on every rank, no broadcast
    np.save(os.path.join(out, "tau_%d.npy" % rank), np.argmax(tau, 2))
    if rank == 0:
        np.save(os.path.join(out, "gamma.npy"), gamma)
    dist.destroy_process_group()


def test_two_rank_sharded_chain_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from conftest import onehot, synth_problem
    from oracle import oracle
    oracle.build()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p = synth_problem(101, 12, 4, depth=20.0, seed=5)
    want = oracle.update(onehot(p["tau0"]), p["gamma0"], p["eta0"], p["counts"], 3, seed=77)
    tau = np.concatenate([np.load(tmp_path / "tau_0.npy"), np.load(tmp_path / "tau_1.npy")])
    assert np.array_equal(tau, np.argmax(want["tau"], 2))              # integer tau independent of the rank count
    assert np.array_equal(np.load(tmp_path / "gamma.npy"), want["gamma"])
