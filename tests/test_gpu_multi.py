"""2-GPU test of the sharded chain (NCCL all-reduce inside the engine): runs only where >= 2 devices are visible."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["DESMAN_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DESMAN_ROOT"], "tests"))
import numpy as np, torch, torch.distributed as dist
from conftest import onehot, synth_problem
from desman_b200 import engine
from desman_b200.parallel import exchange_unique_id, shard_bounds
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl")
grouped = os.environ.get("DESMAN_CASE") == "grouped"
if grouped or os.environ.get("DESMAN_CASE") == "agg":      # converged start, few flips: the screening pass of the tau update is active on every rank
    from test_gpu_group import mild_problem
    p = mild_problem(4001, 64, 8, 30.0, 9)
    tau0, gamma0 = p["tau_true"], p["gamma_true"]
else:
    p = synth_problem(4001, 64, 8, depth=30.0, seed=9, ambiguous=True)
    tau0, gamma0 = p["tau0"], p["gamma0"]
lo, hi = shard_bounds(4001, rank, world)
agg = os.environ.get("DESMAN_CASE") == "agg"
e = engine.Engine(rank, seed=4242)
# per-read contract: statistics independent of how the sites are sharded; "agg": what bench.py runs at N > 1 -- the
# pattern-aggregated contract with shard-keyed streams (mu_mode auto -> 1 at this shape)
e.set_option("mu_mode", 1 if agg else 0)
e.set_option("tau_group", 1 if (grouped or agg) else 0)
e.set_counts(p["counts"][lo:hi], v0=lo, V_total=4001)
e.comm_init(exchange_unique_id(dist, engine.Engine.comm_unique_id), rank, world)
e.set_state(onehot(tau0[lo:hi]), gamma0, p["eta0"])
out = e.update(8)
assert (e.get_timing()["kernel_launches"]["tau_group"] == 8) == (grouped or agg)
open(os.path.join(os.environ["DESMAN_OUT"], "kind%d.txt" % rank), "w").write(e.comm_kind())
np.savez(os.path.join(os.environ["DESMAN_OUT"], "r%d.npz" % rank), tau=e.get_tau_index(), gamma=out["gamma_store"],
         eta=out["eta_store"], ll=out["ll_store"], nchange=out["nchange"], star=e.get_star()["iter"])
e.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("case", ["plain", "grouped"])
def test_two_gpu_sharded_update_equals_one_gpu(tmp_path, case):
    from conftest import onehot, synth_problem
    from desman_b200 import _lib, engine
    if _lib.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, DESMAN_ROOT=ROOT, DESMAN_OUT=str(tmp_path), DESMAN_CASE=case)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    if case == "grouped":
        from test_gpu_group import mild_problem
        p = mild_problem(4001, 64, 8, 30.0, 9)
        tau0, gamma0 = p["tau_true"], p["gamma_true"]
    else:
        p = synth_problem(4001, 64, 8, depth=30.0, seed=9, ambiguous=True)
        tau0, gamma0 = p["tau0"], p["gamma0"]
    e = engine.Engine(0, seed=4242)
    e.set_option("mu_mode", 0)
    e.set_option("tau_group", 0)      # one GPU, per-site kernel only: the sharded grouped run must reproduce it
    e.set_counts(p["counts"])
    e.set_state(onehot(tau0), gamma0, p["eta0"])
    one = e.update(8)
    tau1 = e.get_tau_index()
    e.close()
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(np.concatenate([r0["tau"], r1["tau"]]), tau1)         # integer tau: independent of GPU count
    assert np.array_equal(r0["gamma"], r1["gamma"]) and np.array_equal(r0["gamma"], one["gamma_store"])
    assert np.array_equal(r0["eta"], one["eta_store"])
    assert np.array_equal(r0["nchange"], one["nchange"]) and np.array_equal(r1["nchange"], one["nchange"])
    assert np.allclose(r0["ll"], one["ll_store"], rtol=1e-12, atol=0) and np.array_equal(r0["ll"], r1["ll"])


def _oracle_sharded_chain(oracle_mod, p, tau0, gamma0, eta0, bounds, n_iter, seed):
    """The sharded chain restated with the oracle's primitives: aggregated statistics per shard (streams keyed by the shard's
    first site) summed over the shards, gamma / eta drawn once from the sums, tau per shard under the global site index."""
    tau = [onehot_(tau0[lo:hi]) for lo, hi in bounds]
    gamma, eta = gamma0.copy(), eta0.copy()
    gs, es_, nch = [], [], []
    for k in range(n_iter):
        sm, es = 0, 0
        for (lo, hi), t in zip(bounds, tau):
            a, b = oracle_mod.mu_stats(t, gamma, eta, p["counts"][lo:hi], seed, k, v0=lo, mode=1)
            sm, es = sm + a, es + b
        gamma = oracle_mod.draw_gamma(sm, 0.1, 1e-6, seed, k)
        n = 0
        for (lo, hi), t in zip(bounds, tau):
            n += oracle_mod.sample_tau_philox(t, gamma, eta, p["counts"][lo:hi], seed, k, v0=lo)
        eta = oracle_mod.draw_eta(es, 0.1, seed, k)
        gs.append(gamma.copy()); es_.append(eta.copy()); nch.append(n)
    return np.concatenate([np.argmax(t, 2) for t in tau]), np.array(gs), np.array(es_), np.array(nch)


def onehot_(idx):
    from conftest import onehot
    return onehot(idx)


@pytest.mark.parametrize("world,p2p", [(2, "0"), (2, "1"), (4, "1"), (4, "0")])
def test_sharded_aggregated_chain_vs_oracle(tmp_path, oracle_mod, world, p2p):
    """The configuration bench.py runs at N > 1 (aggregated statistics, shard-keyed streams, screening pass on every rank)
    against the oracle's restatement of the sharded chain -- with the NCCL all-reduce (p2p 0) and with the one-shot
    peer-memory exchange kernel (p2p 1: exchange_kernel.cuh, the default from 4 ranks) as the data plane."""
    from desman_b200 import _lib
    from desman_b200.parallel import shard_bounds
    if _lib.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, DESMAN_ROOT=ROOT, DESMAN_OUT=str(tmp_path), DESMAN_CASE="agg", DESMAN_B200_P2P=p2p)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    from test_gpu_group import mild_problem
    p = mild_problem(4001, 64, 8, 30.0, 9)
    bounds = [shard_bounds(4001, k, world) for k in range(world)]
    tau, gs, es, nch = _oracle_sharded_chain(oracle_mod, p, p["tau_true"], p["gamma_true"], p["eta0"], bounds, 8, 4242)
    rs = [np.load(tmp_path / ("r%d.npz" % k)) for k in range(world)]
    assert np.array_equal(np.concatenate([r["tau"] for r in rs]), tau)
    for r in rs:
        assert np.array_equal(r["nchange"], nch)
        assert np.array_equal(r["gamma"], rs[0]["gamma"]) and np.array_equal(r["eta"], rs[0]["eta"])     # bit-identical on every rank
    assert np.allclose(rs[0]["gamma"], gs, rtol=1e-10, atol=0) and np.allclose(rs[0]["eta"], es, rtol=1e-10, atol=0)
    kinds = {open(tmp_path / ("kind%d.txt" % k)).read() for k in range(world)}
    assert kinds == {"p2p-mailbox" if p2p == "1" else "nccl-allreduce"}, kinds
