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


def run_chain(eng_mod, p, G, group, n_iter, tau0, gamma0, eta0, seed=4242, mu_mode=1, chunks=1, mma=1, tc=1):
    """tc=1: tcgen05 / TMEM / TMA form of the screening pass (default); tc=0, mma=1: mma.sync form; tc=0, mma=0: FFMA form."""
    e = eng_mod.Engine(0, seed=seed)
    e.set_option("tau_group", group)
    e.set_option("tau_group_tc", tc)
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
    a = run_chain(eng_mod, p, G, 1, 8, tau0, p["gamma_true"], p["eta0"])              # tcgen05 screening pass (TMA + TMEM)
    m = run_chain(eng_mod, p, G, 1, 8, tau0, p["gamma_true"], p["eta0"], tc=0)        # mma.sync screening pass
    f = run_chain(eng_mod, p, G, 1, 8, tau0, p["gamma_true"], p["eta0"], tc=0, mma=0) # FFMA screening pass
    b = run_chain(eng_mod, p, G, 0, 8, tau0, p["gamma_true"], p["eta0"])              # per-site kernel only
    for k in ("tau", "nchange", "ll", "lp", "gamma", "tau_sum", "star"):
        assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(m[k], b[k]), k
        assert np.array_equal(f[k], b[k]), k
    assert a["tiers"].sum() == 8 * V * G and b["tiers"].sum() == 8 * V * G and f["tiers"].sum() == 8 * V * G and m["tiers"].sum() == 8 * V * G
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
    for tc, mma in ((1, 1), (0, 1), (0, 0)):
        a = run_chain(eng_mod, p, G, 1, 5, tau0, p["gamma_true"], p["eta0"], mma=mma, tc=tc)
        for k in ("tau", "nchange", "ll", "gamma", "tau_sum", "star"):
            assert np.array_equal(a[k], b[k]), (k, tc, mma)
        assert a["tiers"].sum() == 5 * V * G


# ------------------------------------------------------------------ the tensor-memory contraction itself
def _screen_reference(counts, tau_idx, gamma, eta):
    """D[v][g][j] = sum_{s,b} n[v,s,b] * (log2 q - log2 P) in float64: P = mixture of the site's pattern, q = P with strain g moved
    from its current base to candidate a_j = (cur + 1 + j) & 3 (tau_group_kernel.cuh; c_sample_tau.c:136-170 differences)."""
    V, G = tau_idx.shape
    P = np.einsum("sg,vgb->vsb", gamma, eta[tau_idx])                           # [V,S,4]
    lP = np.log2(P)
    D = np.zeros((V, G, 3))
    n = counts.astype(np.float64)
    for g in range(G):
        cur = tau_idx[:, g]
        base = P - gamma[None, :, g, None] * eta[cur][:, None, :]
        for j in range(3):
            a = (cur + 1 + j) & 3
            q = base + gamma[None, :, g, None] * eta[a][:, None, :]
            D[:, g, j] = (n * (np.log2(q) - lP)).sum((1, 2))
    return D


@pytest.mark.parametrize("V,S,G,depth", [(4000, 64, 8, 100.0), (3000, 64, 5, 20.0), (1500, 7, 3, 8.0), (1200, 130, 12, 10.0),
                                          (900, 256, 16, 30.0), (700, 33, 6, 300.0), (500, 96, 20, 15.0), (300, 5, 1, 30.0),
                                          (2000, 64, 8, 1500.0)])
def test_tc_screening_sums_against_float64(eng_mod, V, S, G, depth):
    """The tcgen05 contraction (fp16 counts x [h | l] split table, FP32 accumulation in tensor memory) against a float64 evaluation
    of the same sums for every grouped site: the observed error must stay inside the bound the gap test charges (this is
    also the measurement of the tensor-core accumulation error the bound assumes), and no step the float64 sums leave open
    (within the 26-nat gap of the current base) may be missing from the work list."""
    p = mild_problem(V, S, G, depth, 17 * V + G)
    if p["counts"].max() >= 2048:
        pytest.skip("counts not exact in fp16")
    tau = p["tau_true"]
    if G > 8:      # 2^G possible patterns: draw the state from a pool of 24 of them so that the sites do share patterns
        tau = tau[np.random.default_rng(G).integers(0, 24, size=V)]
    e = eng_mod.Engine(0, seed=1)
    e.set_option("tau_group", 1)
    e.set_counts(p["counts"])
    e.set_state(onehot(tau), p["gamma_true"], p["eta0"])
    D, mask = e.debug_screen()
    st = e.get_group_stats()
    e.close()
    ref = _screen_reference(p["counts"], tau, p["gamma_true"], p["eta0"])
    grouped = ~np.isnan(D[:, 0, 0])
    assert grouped.sum() >= V // 2 and st["items"] > 0, st
    nsite = p["counts"].sum((1, 2)).astype(np.float64)
    err = np.abs(D[grouped].astype(np.float64) - ref[grouped]).max((1, 2))
    # bound per read in log2 units (tau_group_tc_kernel.cuh: bn_scale / ln2), evaluated like the kernel does
    qmin = 0.99 * np.float32(p["gamma_true"].min()) * np.float32(p["eta0"].min())
    mq0 = max(1.0, 1.0 - np.log2(qmin))
    Sp = 2 * ((S + 3) // 4 * 4)                                                 # >= the padded sample count of any K-block split
    e_entry = (6 * 2.0 ** -24 * 1.4427 + 2 * 2.0 ** -22) + (2.0 ** -22 + 2.0 ** -24) * 2 * mq0 + (2 * G + 2) * 2.0 ** -53 * 1.4427 / qmin + 2.0 ** -25
    bound = nsite[grouped] * (e_entry + (Sp / 2 + 12) * 2.0 ** -20 * mq0) + 1e-6
    assert (err <= bound).all(), (float((err / bound).max()), int(np.argmax(err / bound)))
    print("tc screening: max |D - D64| / bound = %.4f, max abs err %.3e log2 units over %d sites" % (float((err / bound).max()), float(err.max()), int(grouped.sum())))
    # every step that is open by the float64 sums is on the work list with its bit set
    open64 = (ref.max(2) * np.log(2.0) > -26.0)                                # [V,G]  (TAU_GAP, tau_kernel.cuh)
    listed = mask != 0xFFFFFFFF
    for v in np.flatnonzero(grouped & open64.any(1)):
        assert listed[v], v
        want = sum(1 << g for g in range(G) if open64[v, g])
        assert (int(mask[v]) & want) == want, (v, hex(int(mask[v])), hex(want))
