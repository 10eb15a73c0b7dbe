"""GPU parity tests of the pattern-grouped tau update (tau_group_kernel.cuh + the work-list mode of tau_sample_kernel):
the screening pass may only ever decide "stay" where the per-site kernel would, so whole chains must be bit-identical
with the grouping on and off, and both equal to the oracle."""
import numpy as np
import pytest

from conftest import onehot, synth_problem

pytestmark = pytest.mark.gpu


def mild_problem(V, S, G, depth, seed, weak=0.3):
    """Biallelic sites with one weak strain (abundance x `weak`) and a few sites of near-zero depth: most (v,g) steps are
    decided by a wide margin, some are not, and a handful of sites flip every sweep (they become orphans of their
    group) -- the regime the screening pass is built for.  conftest.synth_problem(ambiguous=True) flips ~1 strain per
    site and sweep and keeps the screening switched off."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, size=V)
    alt = (anc + rng.integers(1, 4, size=V)) % 4
    carry = rng.random((V, G)) < 0.3
    bad = carry.all(1) | ~carry.any(1)
    carry[bad, 0] = ~carry[bad, 0]
    tau_true = np.where(carry, alt[:, None], anc[:, None]).astype(np.uint8)
    gamma_true = rng.dirichlet(np.ones(G), size=S)
    gamma_true[:, 0] *= weak
    gamma_true /= gamma_true.sum(1)[:, None]
    eta_true = 0.997 * np.identity(4) + 0.001 * (1 - np.identity(4))
    p = np.einsum("sg,vga->vsa", gamma_true, eta_true[tau_true])
    p /= p.sum(-1, keepdims=True)
    N = rng.poisson(depth, size=(V, S))
    amb = rng.choice(V, size=max(4, V // (25 * G)), replace=False)
    N[amb] = rng.poisson(depth * 0.02, size=(len(amb), S))
    counts = rng.multinomial(N, p).astype(np.int64)
    tau0 = rng.integers(0, 4, size=(V, G)).astype(np.uint8)
    gamma0 = rng.dirichlet(np.ones(G), size=S)
    gamma0[gamma0 < 1e-6] = 1e-6
    gamma0 /= gamma0.sum(1)[:, None]
    eta0 = 0.96 * np.identity(4) + 0.01 * np.ones((4, 4))
    return dict(counts=counts, tau_true=tau_true, gamma_true=gamma_true, tau0=tau0, gamma0=gamma0, eta0=eta0)


@pytest.fixture(scope="module")
def eng_mod():
    from desman_b200 import _lib, engine
    assert _lib.device_count() >= 1
    return engine


def run_chain(eng_mod, p, G, group, n_iter, tau0, gamma0, eta0, seed=4242, mu_mode=1, chunks=1, mma=1):
    e = eng_mod.Engine(0, seed=seed)
    e.set_option("tau_group", group)
    e.set_option("tau_group_mma", mma)
    e.set_option("mu_mode", mu_mode)
    e.set_counts(p["counts"])
    e.set_state(onehot(tau0), gamma0, eta0)
    e.get_tier_counts()
    outs = [e.update(n_iter // chunks) for _ in range(chunks)]
    out = {k: np.concatenate([o[k] for o in outs]) for k in outs[0]}
    res = dict(tau=e.get_tau_index(), nchange=out["nchange"], ll=out["ll_store"], lp=out["lp_store"],
               gamma=out["gamma_store"], tau_sum=e.get_tau_sum(), star=e.get_star_index(),
               tiers=e.get_tier_counts(), stats=e.get_group_stats(), launches=e.get_timing()["kernel_launches"])
    e.close()
    return res


CASES = [  # V, S, G, depth: strain blocks of 8 (G = 5, 8, 6), of 4 (G = 3, 4, 9, 12), ragged S, S < 8, deep and shallow counts
    (3000, 64, 5, 20.0), (2500, 64, 8, 30.0), (1500, 7, 3, 8.0), (1200, 130, 12, 10.0), (1000, 40, 4, 5.0),
    (900, 64, 9, 25.0), (700, 33, 6, 300.0), (400, 3, 2, 50.0),
    (300, 5, 1, 30.0),        # one strain: patterns are the four bases
    (600, 16, 4, 4000.0),     # counts >= 2048 are not exact in TF32: the FFMA form takes over whatever tau_group_mma says
]


@pytest.mark.parametrize("V,S,G,depth", CASES)
def test_grouped_chain_identical_to_per_site_chain_from_converged_state(eng_mod, oracle_mod, V, S, G, depth):
    """Start at the true haplotypes (few patterns, many sites each): the screening pass is active from the first sweep."""
    p = mild_problem(V, S, G, depth, 31 * V + G)
    tau0 = p["tau_true"]
    a = run_chain(eng_mod, p, G, 1, 8, tau0, p["gamma_true"], p["eta0"])              # tensor-core screening pass
    f = run_chain(eng_mod, p, G, 1, 8, tau0, p["gamma_true"], p["eta0"], mma=0)       # FFMA screening pass
    b = run_chain(eng_mod, p, G, 0, 8, tau0, p["gamma_true"], p["eta0"])              # per-site kernel only
    for k in ("tau", "nchange", "ll", "lp", "gamma", "tau_sum", "star"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(f[k], b[k]), k
    assert a["tiers"].sum() == 8 * V * G and b["tiers"].sum() == 8 * V * G and f["tiers"].sum() == 8 * V * G
    assert a["launches"]["tau_group"] == 8 and f["launches"]["tau_group"] == 8 and b["launches"]["tau_group"] == 0
    # (the two forms of the screening pass use different error bounds, so they need not decide exactly the same steps)
    st = a["stats"]
    assert st["configured"] == 1
    if a["nchange"].max() <= V // 16:                # calm throughout: the groups were kept and used in every sweep
        assert st["have"] == 1 and st["calm"] == 1 and st["items"] > 0, st
        assert st["work"] + st["singles"] <= V, st   # the per-site kernel walked only part of the sites in the last sweep
    if V * S * G <= 3000 * 64 * 5:
        want = oracle_mod.update(onehot(tau0), p["gamma_true"], p["eta0"], p["counts"], 8, 4242, mu_mode=1)
        assert np.array_equal(a["tau"], np.argmax(want["tau"], 2))
        assert np.array_equal(a["nchange"], want["nchange"])
        assert np.allclose(a["ll"], want["ll_store"], rtol=1e-9, atol=0)


@pytest.mark.parametrize("V,S,G,depth", [(4000, 64, 5, 30.0), (3000, 24, 4, 60.0)])
def test_grouped_chain_identical_from_random_start_across_update_calls(eng_mod, V, S, G, depth):
    """Random start: burn-in sweeps flip almost every site (screening off, table rebuilt), then the chain calms down,
    the sites are regrouped and the screening pass takes over; flipped sites become orphans.  Split over three update()
    calls so that the groups persist across calls."""
    p = mild_problem(V, S, G, depth, 5 * V + G)
    a = run_chain(eng_mod, p, G, 1, 24, p["tau0"], p["gamma0"], p["eta0"], chunks=3)
    b = run_chain(eng_mod, p, G, 0, 24, p["tau0"], p["gamma0"], p["eta0"], chunks=3)
    for k in ("tau", "nchange", "ll", "lp", "gamma", "star"):
        assert np.array_equal(a[k], b[k]), k
    assert a["nchange"][0] > V            # a real burn-in
    st = a["stats"]
    assert st["have"] == 1 and st["calm"] == 1 and st["work"] + st["singles"] < V // 2, st


def test_grouped_chain_per_read_statistics_mode(eng_mod):
    """The grouping is independent of which statistics kernel runs (mu_mode 0: one categorical draw per read)."""
    p = mild_problem(1500, 64, 5, 15.0, 77)
    tau0 = p["tau_true"]
    a = run_chain(eng_mod, p, 5, 1, 5, tau0, p["gamma_true"], p["eta0"], mu_mode=0)
    b = run_chain(eng_mod, p, 5, 0, 5, tau0, p["gamma_true"], p["eta0"], mu_mode=0)
    for k in ("tau", "nchange", "ll", "gamma"):
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("rng_mode", ["mt", "philox"])
def test_grouped_update_tau_replay_identical(eng_mod, oracle_mod, rng_mode):
    """updateTau() replay (HaploSNP_Sampler.py:383-407) with stored gamma/eta; in MT19937 mode the screening pass must
    hand every site with a zero word (u == 0) to the per-site kernel, so results equal the reference-order oracle."""
    V, S, G, n_iter = 2000, 64, 5, 6
    p = mild_problem(V, S, G, 20.0, 99)
    tau0 = p["tau_true"]
    rng = np.random.default_rng(3)
    gs = np.stack([p["gamma_true"] * rng.uniform(0.8, 1.25, size=p["gamma_true"].shape) for _ in range(n_iter)])
    gs /= gs.sum(2)[:, :, None]
    es = np.stack([p["eta0"]] * n_iter)
    res = {}
    for group in (1, 0):
        e = eng_mod.Engine(0, seed=1234, rng_mode=eng_mod.RNG_MT19937 if rng_mode == "mt" else eng_mod.RNG_PHILOX)
        e.set_option("tau_group", group)
        e.set_counts(p["counts"])
        e.set_state(onehot(tau0), gs[0], es[0])
        out = e.update_tau(gs, es)
        res[group] = (e.get_tau_index(), out["nchange"], out["ll_store"], e.get_star_index(), e.get_tau_sum())
        e.close()
    for x, y in zip(res[1], res[0]):
        assert np.array_equal(x, y)
    tau_o = onehot(tau0)
    want = oracle_mod.update_tau(tau_o, gs, es, p["counts"], 1234, use_mt=(rng_mode == "mt"))
    assert np.array_equal(res[1][0], np.argmax(want["tau"], 2))
    assert np.array_equal(res[1][1], want["nchange"])


def test_grouped_chain_at_C2_size_matches_exact_path(eng_mod):
    """V=10000 S=64 G=8 (auto rule turns the grouping on): 12 sweeps from the true state, grouped filtered path vs the
    FP64 reference-order path for every draw."""
    p = synth_problem(10000, 64, 8, depth=100.0, seed=20240611)
    tau0 = p["tau_true"].astype(np.uint8)
    res = {}
    for exact in (0, 1):
        e = eng_mod.Engine(0, seed=23724839)
        e.set_option("tau_exact", exact)
        e.set_counts(p["counts"])
        e.set_state(onehot(tau0), p["gamma_true"], p["eta0"])
        out = e.update(12)
        res[exact] = (e.get_tau_index(), out["nchange"], out["ll_store"], e.get_group_stats(), e.get_timing()["kernel_launches"])
        e.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert res[0][4]["tau_group"] == 12 and res[1][4]["tau_group"] == 0
    assert res[0][3]["work"] + res[0][3]["singles"] < 2000


def _random_shapes(n, seed=2024):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        G = int(rng.integers(1, 13))
        S = int(rng.choice([1, 2, 5, 8, 15, 16, 17, 31, 33, 48, 64, 65, 100, 129]))
        V = int(rng.integers(150, 900))
        depth = float(rng.choice([2.0, 10.0, 60.0, 400.0]))
        out.append((V, S, G, depth))
    return out


@pytest.mark.parametrize("V,S,G,depth", _random_shapes(10))
def test_grouped_chain_identical_random_shapes(eng_mod, V, S, G, depth):
    """Seeded random shapes (ragged S around the 8/16/32-sample tile edges, G from 1 to 12, shallow to deep counts): the tensor-core
    and FFMA forms of the screening pass against the per-site kernel alone, 5 sweeps from the true state."""
    p = mild_problem(V, S, G, depth, 977 * V + 31 * S + G)
    tau0 = p["tau_true"]
    b = run_chain(eng_mod, p, G, 0, 5, tau0, p["gamma_true"], p["eta0"])
    for mma in (1, 0):
        a = run_chain(eng_mod, p, G, 1, 5, tau0, p["gamma_true"], p["eta0"], mma=mma)
        for k in ("tau", "nchange", "ll", "gamma", "tau_sum", "star"):
            assert np.array_equal(a[k], b[k]), (k, mma)
        assert a["tiers"].sum() == 5 * V * G
