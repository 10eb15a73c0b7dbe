"""GPU parity at the BASELINE sizes against an INDEPENDENT implementation (the CPU oracle, and the reference's own
compiled c_sample_tau.c where oracle/_ref travelled with the snapshot) -- not against another GPU mode.

What only shows at scale is covered here: the 2.5 V-slot pattern table, regroup / counting sort over 1e4..1e5 sites, the
work list, the lazy occupancy counters, table rebuilds during burn-in, and the per-read / ungrouped paths at G >= 16.

  (a) C2  V=10000 S=64 G=8, 200 sweeps of update() from the benchmark's random start   vs oracle.update
  (b) C2  10 c_sample_tau calls (MT19937 stream, fresh gamma per call)                  vs oracle/_ref (reference C)
  (c) C3  V=100000 S=64 G=8, 12 sweeps                                                  vs oracle.update
  (d) C4  V=100000 S=256 G=16 and the C5 shard V=125000 S=128 G=20: the engine's own gamma/eta trace replayed through the
      oracle's tau update on 2000-site slices (the Philox draws are keyed by the global site index, so a slice's tau must
      equal the engine's), and the device ll against the oracle's on the slice-independent whole only where it is cheap.
Integers bit-exact; gamma / eta 1e-10, ll / lp 1e-9 relative (north star: 1e-6).
"""
import ctypes as C

import numpy as np
import pytest

from conftest import onehot

pytestmark = pytest.mark.gpu

SEED = 23724839
RTOL_TIGHT = 1e-10
RTOL_LL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


@pytest.fixture(scope="module")
def eng_mod():
    from desman_b200 import _lib, engine
    assert _lib.device_count() >= 1
    return engine


def _chain_vs_oracle(eng_mod, oracle_mod, V, S, G, n_iter):
    from desman_b200.synth import synth_counts
    p = synth_counts(V, S, G)
    mode = eng_mod.auto_mu_mode(V, G)
    want = oracle_mod.update(onehot(p["tau0"]), p["gamma0"], p["eta0"], p["counts"], n_iter, SEED, mu_mode=mode)
    e = eng_mod.Engine(0, seed=SEED)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
    got = e.update(n_iter)
    tau_idx = e.get_tau_index()
    _, gamma, eta = e.get_state(want_tau=False)
    star = e.get_star(want_tau=False)
    star_idx = e.get_star_index()
    tau_sum = e.get_tau_sum(compact=True)
    stats = e.get_group_stats()
    e.close()
    assert np.array_equal(got["nchange"], want["nchange"])
    assert np.array_equal(tau_idx, np.argmax(want["tau"], 2).astype(np.uint8))
    assert np.array_equal(tau_sum.astype(np.int64), want["tau_sum"])
    assert star["iter"] == want["iter_star"]
    assert np.array_equal(star_idx, np.argmax(want["tau_star"], 2).astype(np.uint8))
    assert rel(got["gamma_store"], want["gamma_store"]) < RTOL_TIGHT
    assert rel(got["eta_store"], want["eta_store"]) < RTOL_TIGHT
    assert rel(gamma, want["gamma"]) < RTOL_TIGHT and rel(eta, want["eta"]) < RTOL_TIGHT
    assert rel(got["ll_store"], want["ll_store"]) < RTOL_LL and rel(got["lp_store"], want["lp_store"]) < RTOL_LL
    assert abs(star["lp"] - want["lp_star"]) <= RTOL_LL * abs(want["lp_star"])
    assert rel(star["gamma"], want["gamma_star"]) < RTOL_TIGHT and rel(star["eta"], want["eta_star"]) < RTOL_TIGHT
    assert want["nchange"][0] > V and want["nchange"][-1] < V // 50         # a burn-in, then a calm chain
    return stats


def test_C2_200_sweeps_vs_oracle(eng_mod, oracle_mod):
    """BASELINE config 2 as stated: V=10000 S=64 G=8, 200 iterations, fixed seed, vs the CPU chain."""
    stats = _chain_vs_oracle(eng_mod, oracle_mod, 10000, 64, 8, 200)
    assert stats["have"] == 1 and stats["items"] > 0                         # the grouped steady-state path carried the chain


def test_C3_12_sweeps_vs_oracle(eng_mod, oracle_mod):
    """BASELINE config 3 shape (V=100000 S=64 G=8): burn-in, table rebuilds, regroup, then the screening pass."""
    stats = _chain_vs_oracle(eng_mod, oracle_mod, 100000, 64, 8, 12)
    assert stats["have"] == 1 and stats["items"] > 0


def test_C2_mt_stream_vs_reference_C(oracle_mod):
    """10 sample_tau calls at C2 size through the drop-in ABI against the reference's own compiled C (oracle/_ref) -- or, where
    that library did not travel, the oracle's restatement driven by the same MT19937 stream.  gamma is redrawn per call."""
    from desman_b200 import sampletau
    from desman_b200.synth import synth_counts
    V, S, G = 10000, 64, 8
    p = synth_counts(V, S, G)
    rng = np.random.default_rng(5)
    tau_g, tau_r = onehot(p["tau0"]), onehot(p["tau0"])
    use_ref = oracle_mod.have_ref()
    if use_ref:
        R = oracle_mod.RefSampleTau(SEED)
    else:
        st = oracle_mod.MT19937()
        oracle_mod.lib().oracle_mt_seed(C.byref(st), SEED)
    sampletau.initRNG()
    sampletau.setRNG(SEED)
    flips = []
    for k in range(10):
        if k < 3:
            gamma = rng.dirichlet(np.ones(G), size=S)
        else:                                              # near the truth: the calm regime the screening pass serves
            gamma = p["gamma_true"] * rng.uniform(0.9, 1.1, size=(S, G))
        gamma[gamma < 1e-6] = 1e-6
        gamma /= gamma.sum(1)[:, None]
        eta = p["eta0"] if k < 5 else 0.997 * np.identity(4) + 0.001 * (1 - np.identity(4))
        n_gpu = sampletau.sample_tau(tau_g, gamma, eta, p["counts"])
        if use_ref:
            n_ref = R.sample_tau(tau_r, gamma, eta, p["counts"])
        else:
            n_ref = oracle_mod.lib().oracle_sample_tau_mt(
                tau_r.ctypes.data_as(oracle_mod._p64), oracle_mod._f64(gamma)[1], oracle_mod._f64(eta)[1],
                oracle_mod._i64(p["counts"])[1], V, G, S, C.byref(st))
        assert n_gpu == n_ref, k
        assert np.array_equal(tau_g, tau_r), k
        flips.append(n_ref)
    sampletau.freeRNG()
    if use_ref:
        R.close()
    assert flips[0] > V and flips[-1] < V // 10


@pytest.mark.parametrize("name,V,S,G", [("C4", 100000, 256, 16), ("C5_shard", 125000, 128, 20)])
def test_C4_C5_tau_slices_vs_oracle_replay(eng_mod, oracle_mod, name, V, S, G):
    """The engine's chain at the C4 / C5-shard shapes, one sweep per update() call; after every sweep the tau of three
    2000-site slices must equal the oracle's tau update replayed on the slice with the engine's own gamma (this sweep's) and
    eta (the previous sweep's: tau is drawn before eta is committed, HaploSNP_Sampler.py:345-347)."""
    from desman_b200.synth import synth_counts
    p = synth_counts(V, S, G)
    n_iter = 5
    slices = [(0, 2000), (V // 2 - 1000, V // 2 + 1000), (V - 2000, V)]
    e = eng_mod.Engine(0, seed=SEED)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
    tau_o = [onehot(p["tau0"][lo:hi]) for lo, hi in slices]
    eta_prev = p["eta0"]
    total = 0
    for it in range(n_iter):
        out = e.update(1)
        tau_idx = e.get_tau_index()
        gamma = out["gamma_store"][0]
        ll_whole = out["ll_store"][0]
        n_o = 0
        for (lo, hi), t in zip(slices, tau_o):
            n_o += oracle_mod.sample_tau_philox(t, gamma, eta_prev, p["counts"][lo:hi], SEED, it, v0=lo)
            assert np.array_equal(np.argmax(t, 2).astype(np.uint8), tau_idx[lo:hi]), (name, it, lo)
        eta_prev = out["eta_store"][0]
        total += n_o
        assert np.isfinite(ll_whole)
    e.close()
    assert total > 0
