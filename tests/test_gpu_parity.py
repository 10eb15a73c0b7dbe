"""GPU parity tests: the CUDA path, called through the C-ABI (ctypes), against the oracle and
against golden vectors recorded from the unmodified reference.  Integer results (tau, nchange,
sum_mu, Esum, tau sums) must be bit-exact; gamma/eta/ll/lp within the stated relative tolerance."""
import numpy as np
import pytest

from conftest import golden, onehot, synth_problem

pytestmark = pytest.mark.gpu

RTOL = 1e-6          # north_star tolerance for gamma / eta / log-likelihood
RTOL_TIGHT = 1e-10   # what we actually expect between CUDA and glibc double arithmetic
RTOL_LL = 1e-9       # ll/lp: summed over the pattern table in 64-bit fixed point (order-independent, 2^-k resolution)


@pytest.fixture(scope="module")
def eng_mod():
    from desman_b200 import _lib, engine
    assert _lib.device_count() >= 1
    return engine


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


# ------------------------------------------------------------------ reference ABI (MT19937 stream)
def test_dropin_sample_tau_matches_reference_C_kats():
    from desman_b200 import sampletau
    z = golden("sample_tau_kat.npz")
    for ci in range(int(z["ncases"])):
        V, G, S, ncalls, seed = [int(x) for x in z[f"c{ci}_meta"]]
        counts = z[f"c{ci}_counts"].astype(np.int64)
        tau = onehot(z[f"c{ci}_tau0"])
        sampletau.initRNG()
        sampletau.setRNG(seed)
        for k in range(ncalls):
            n = sampletau.sample_tau(tau, np.ascontiguousarray(z[f"c{ci}_gamma"][k]),
                                     np.ascontiguousarray(z[f"c{ci}_eta"][k]), counts)
            assert np.array_equal(np.argmax(tau, 2), z[f"c{ci}_tau"][k]), (ci, k)
            assert n == int(z[f"c{ci}_nchange"][k]), (ci, k)
            assert (tau.sum(2) == 1).all() and tau.min() == 0
        sampletau.freeRNG()


@pytest.mark.parametrize("tag", ["i3", "i50"])
def test_dropin_sample_tau_replays_real_reference_chain(tag):
    """Every sample_tau call the unmodified `desman -g 5 -i N` made on COG0015 (burn-in + sampling
    phases share ONE MT19937 stream): same inputs -> identical tau and nchange."""
    from desman_b200 import sampletau
    z = golden(f"cog0015_{tag}.npz")
    counts = golden("cog0015.npz")["snps"].astype(np.int64)
    seed = int(z["meta"][4])
    sampletau.initRNG()
    sampletau.setRNG(seed)
    k = 0
    for phase in ("call", "call2"):
        for i in range(z[f"{phase}_tau_in"].shape[0]):
            tau = onehot(z[f"{phase}_tau_in"][i])
            n = sampletau.sample_tau(tau, np.ascontiguousarray(z[f"{phase}_gamma"][i]),
                                     np.ascontiguousarray(z[f"{phase}_eta"][i]), counts)
            assert n == int(z["call_nchange"][k]), (phase, i)
            assert np.array_equal(np.argmax(tau, 2), z[f"{phase}_tau_out"][i]), (phase, i)
            k += 1
    sampletau.freeRNG()


def test_dropin_argument_errors_and_rejections():
    from desman_b200 import _lib, sampletau
    sampletau.initRNG()
    sampletau.setRNG(1)
    tau = onehot(np.zeros((4, 2), dtype=np.uint8))
    pi = np.full((3, 2), 0.5)
    eta = 0.96 * np.identity(4) + 0.01
    var = np.ones((4, 3, 4), dtype=np.int64)
    bad = tau.copy(); bad[1, 0, :] = 0
    with pytest.raises(_lib.DesmanB200Error, match="one-hot"):
        sampletau.sample_tau(bad, pi, eta, var)
    big = var.copy(); big[0, 0, 0] = 2**24 + 1
    with pytest.raises(_lib.DesmanB200Error, match="counts"):
        sampletau.sample_tau(tau, pi, eta, big)
    assert sampletau.sample_tau(tau, pi, eta, var) >= 0
    sampletau.freeRNG()


# ------------------------------------------------------------------ single kernels vs oracle
SHAPES = [(37, 1, 1), (50, 3, 2), (64, 64, 5), (40, 64, 8), (33, 130, 12), (20, 256, 16), (17, 96, 20), (9, 40, 32)]


@pytest.mark.parametrize("V,S,G", SHAPES)
def test_tau_philox_bit_exact_vs_oracle(eng_mod, oracle_mod, V, S, G):
    p = synth_problem(V, S, G, depth=6.0 if S > 8 else 30.0, seed=V * 1000 + G, ambiguous=True)
    p["counts"][:2] = 0                                   # all-zero rows: uniform draw t = floor(4u)
    e = eng_mod.Engine(0, seed=77, rng_mode=eng_mod.RNG_PHILOX)
    e.set_counts(p["counts"])
    tau_o = onehot(p["tau0"])
    e.set_state(tau_o, p["gamma0"], p["eta0"])
    rng = np.random.default_rng(G)
    total_flips = 0
    for k in range(4):
        gamma = rng.dirichlet(np.full(G, 0.3 if k % 2 else 1.0), size=S)
        gamma[gamma < 1e-6] = 1e-6
        gamma /= gamma.sum(1)[:, None]
        e.set_state(None, gamma, p["eta0"], G=G)
        e.set_rng(77, sweep=k)
        n_gpu = e.sample_tau()
        n_cpu = oracle_mod.sample_tau_philox(tau_o, gamma, p["eta0"], p["counts"], 77, k)
        assert n_gpu == n_cpu
        assert np.array_equal(e.get_tau_index(), np.argmax(tau_o, 2).astype(np.uint8))
        total_flips += n_cpu
    assert total_flips > 0          # the vectors are not vacuous
    e.close()


@pytest.mark.parametrize("top", [65535, 65536, 70000, 2**24])
def test_counts_upload_narrow_and_wide_cells(eng_mod, oracle_mod, top):
    """desman_set_counts ships counts < 2^16 as 4 x uint16 per cell and widens them on the device; a chunk with a larger count
    takes the int32x4 form: the tau step (which reads every cell) must equal the oracle's on both sides of the boundary."""
    V, S, G = 48, 40, 4
    p = synth_problem(V, S, G, depth=30.0, seed=5, ambiguous=True)
    p["counts"][3, 7, 1] = top
    p["counts"][V - 1, S - 1, 3] = min(top, 65535)
    e = eng_mod.Engine(0, seed=9, rng_mode=eng_mod.RNG_PHILOX)
    e.set_counts(p["counts"])
    tau_o = onehot(p["tau0"])
    e.set_state(tau_o, p["gamma0"], p["eta0"])
    for k in range(2):
        e.set_rng(9, sweep=k)
        n_gpu = e.sample_tau()
        n_cpu = oracle_mod.sample_tau_philox(tau_o, p["gamma0"], p["eta0"], p["counts"], 9, k)
        assert n_gpu == n_cpu
        assert np.array_equal(e.get_tau_index(), np.argmax(tau_o, 2).astype(np.uint8))
    ll_gpu = e.loglik()
    assert np.isclose(ll_gpu[0] if isinstance(ll_gpu, (tuple, list)) else ll_gpu,
                      oracle_mod.loglik(tau_o, p["gamma0"], p["eta0"], p["counts"]), rtol=1e-10)
    e.close()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("V,S,G", SHAPES)
def test_mu_stats_bit_exact_vs_oracle(eng_mod, oracle_mod, V, S, G, mode):
    """mode 0: one categorical draw per read; mode 1 (default): pattern-aggregated conditional binomials."""
    p = synth_problem(V, S, G, depth=40.0, seed=V + G)
    p["counts"][0] = 0
    p["counts"][1, :, 2] = 777
    p["tau0"][V // 2:] = p["tau0"][V // 2]                      # many sites share a pattern -> large aggregated counts
    eta = 0.9 * np.identity(4) + 0.025
    e = eng_mod.Engine(0, seed=2024, rng_mode=eng_mod.RNG_PHILOX)
    e.set_option("mu_mode", mode)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau0"]), p["gamma0"], eta)
    for sweep in (0, 5):
        e.set_rng(2024, sweep=sweep)
        sm, es = e.mu_stats()
        sm_o, es_o = oracle_mod.mu_stats(onehot(p["tau0"]), p["gamma0"], eta, p["counts"], 2024, sweep, mode=mode)
        assert np.array_equal(sm, sm_o)
        assert np.array_equal(es, es_o)
        assert sm.sum() == p["counts"].sum()
    e.close()


def test_mu_stats_sharded_offsets_equal_whole(eng_mod, oracle_mod):
    """Counter contract is keyed by the GLOBAL site index: two half-shards sum to the whole."""
    p = synth_problem(90, 40, 6, depth=25.0, seed=3)
    whole = oracle_mod.mu_stats(onehot(p["tau0"]), p["gamma0"], p["eta0"], p["counts"], 9, 4)
    acc_sm, acc_es = 0, 0
    for lo, hi in ((0, 41), (41, 90)):
        e = eng_mod.Engine(0, seed=9)
        e.set_option("mu_mode", 0)
        e.set_counts(p["counts"][lo:hi], v0=lo, V_total=90)
        e.set_state(onehot(p["tau0"][lo:hi]), p["gamma0"], p["eta0"])
        e.set_rng(9, sweep=4)
        sm, es = e.mu_stats()
        acc_sm, acc_es = acc_sm + sm, acc_es + es
        e.close()
    assert np.array_equal(acc_sm, whole[0]) and np.array_equal(acc_es, whole[1])
    # aggregated mode: a shard's streams are keyed by its first global site; each shard equals the oracle's shard
    for lo, hi in ((0, 41), (41, 90)):
        e = eng_mod.Engine(0, seed=9)
        e.set_option("mu_mode", 1)
        e.set_counts(p["counts"][lo:hi], v0=lo, V_total=90)
        e.set_state(onehot(p["tau0"][lo:hi]), p["gamma0"], p["eta0"])
        e.set_rng(9, sweep=4)
        sm, es = e.mu_stats()
        want = oracle_mod.mu_stats(onehot(p["tau0"][lo:hi]), p["gamma0"], p["eta0"], p["counts"][lo:hi], 9, 4, v0=lo, mode=1)
        assert np.array_equal(sm, want[0]) and np.array_equal(es, want[1])
        e.close()


@pytest.mark.parametrize("S,G", [(1, 1), (5, 3), (64, 8), (256, 16), (128, 20)])
def test_draw_gamma_eta_vs_oracle(eng_mod, oracle_mod, S, G):
    rng = np.random.default_rng(S * 100 + G)
    sm = rng.integers(0, 5000, size=(S, G)).astype(np.int64)
    sm[rng.random((S, G)) < 0.3] = 0
    es = np.diag(rng.integers(10**4, 10**6, 4)).astype(np.int64) + rng.integers(0, 50, size=(4, 4))
    p = synth_problem(4, S, G, depth=5.0)
    e = eng_mod.Engine(0, seed=31337)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
    e.set_rng(31337, sweep=12)
    g, et = e.draw_gamma_eta(sm, es)
    g_o = oracle_mod.draw_gamma(sm, 0.1, 1e-6, 31337, 12)
    e_o = oracle_mod.draw_eta(es, 0.1, 31337, 12)
    assert rel(g, g_o) < RTOL_TIGHT and rel(et, e_o) < RTOL_TIGHT
    assert np.allclose(g.sum(1), 1.0, atol=1e-12) and np.allclose(et.sum(1), 1.0, atol=1e-12)
    e.close()


def test_loglik_vs_reference_python_golden(eng_mod, oracle_mod):
    z = golden("loglik_kat.npz")
    for ci in range(int(z["ncases"])):
        counts = z[f"c{ci}_counts"].astype(np.int64)
        tau = onehot(z[f"c{ci}_tau"])
        e = eng_mod.Engine(0, seed=1)
        e.set_counts(counts)
        e.set_state(tau, z[f"c{ci}_gamma"], z[f"c{ci}_eta"])
        ll, lp = e.loglik()
        want_ll, want_lp = z[f"c{ci}_ll_lp"]
        assert abs(ll - want_ll) <= RTOL_LL * abs(want_ll), ci
        assert abs(lp - want_lp) <= RTOL_LL * abs(want_lp), ci
        e.close()


# ------------------------------------------------------------------ chains vs oracle
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("V,S,G,depth,n_iter", [(120, 16, 3, 30.0, 25), (300, 64, 8, 20.0, 12), (64, 130, 12, 8.0, 6),
                                                  (2000, 64, 5, 60.0, 10)])
def test_update_chain_vs_oracle(eng_mod, oracle_mod, V, S, G, depth, n_iter, mode):
    p = synth_problem(V, S, G, depth=depth, seed=11 + G, ambiguous=True)
    seed = 23724839
    want = oracle_mod.update(onehot(p["tau0"]), p["gamma0"], p["eta0"], p["counts"], n_iter, seed, sweep0=3, mu_mode=mode)
    e = eng_mod.Engine(0, seed=seed)
    e.set_option("mu_mode", mode)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
    e.set_rng(seed, sweep=3)
    got = e.update(n_iter)
    tau, gamma, eta = e.get_state()
    assert np.array_equal(got["nchange"], want["nchange"])
    assert np.array_equal(tau, want["tau"])                                   # bit-exact integer tau
    assert np.array_equal(e.get_tau_sum(), want["tau_sum"])
    assert rel(got["gamma_store"], want["gamma_store"]) < RTOL_TIGHT
    assert rel(got["eta_store"], want["eta_store"]) < RTOL_TIGHT
    assert rel(got["ll_store"], want["ll_store"]) < RTOL_LL
    assert rel(got["lp_store"], want["lp_store"]) < RTOL_LL
    assert rel(gamma, want["gamma"]) < RTOL_TIGHT and rel(eta, want["eta"]) < RTOL_TIGHT
    star = e.get_star()
    assert star["iter"] == want["iter_star"]
    assert abs(star["lp"] - want["lp_star"]) <= RTOL_LL * abs(want["lp_star"])
    assert np.array_equal(star["tau"], want["tau_star"])
    assert rel(star["gamma"], want["gamma_star"]) < RTOL_TIGHT and rel(star["eta"], want["eta_star"]) < RTOL_TIGHT
    assert want["nchange"].sum() > 0 and e.get_rng()[0] == 3 + n_iter
    # second update() continues the stream exactly like one long chain split in two
    want2 = oracle_mod.update(want["tau"], want["gamma"], want["eta"], p["counts"], 3, seed, sweep0=3 + n_iter, mu_mode=mode)
    got2 = e.update(3)
    assert np.array_equal(got2["nchange"], want2["nchange"])
    assert np.array_equal(e.get_state()[0], want2["tau"])
    e.close()


@pytest.mark.parametrize("use_mt", [True, False])
def test_update_tau_replay_vs_oracle(eng_mod, oracle_mod, use_mt):
    p = synth_problem(150, 24, 4, depth=10.0, seed=5, ambiguous=True)
    rng = np.random.default_rng(0)
    n_iter = 9
    gs = rng.dirichlet(np.ones(4), size=(n_iter, 24))
    gs[gs < 1e-6] = 1e-6
    gs /= gs.sum(2)[:, :, None]
    es = np.tile(p["eta0"], (n_iter, 1, 1)) * rng.uniform(0.9, 1.1, size=(n_iter, 4, 4))
    es /= es.sum(2)[:, :, None]
    want = oracle_mod.update_tau(onehot(p["tau0"]), gs, es, p["counts"], seed=4242, sweep0=0, use_mt=use_mt)
    e = eng_mod.Engine(0, seed=4242, rng_mode=eng_mod.RNG_MT19937 if use_mt else eng_mod.RNG_PHILOX)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau0"]), gs[0], es[0])
    got = e.update_tau(gs, es)
    assert np.array_equal(got["nchange"], want["nchange"])
    assert np.array_equal(e.get_state()[0], want["tau"])
    assert np.array_equal(e.get_tau_sum(), want["tau_sum"])
    assert rel(got["ll_store"], want["ll_store"]) < RTOL_LL and rel(got["lp_store"], want["lp_store"]) < RTOL_LL
    star = e.get_star()
    assert np.array_equal(star["tau"], want["tau_star"])
    assert abs(star["lp"] - want["lp_star"]) <= RTOL_LL * abs(want["lp_star"])
    e.close()


# ------------------------------------------------------------------ size-independent properties at scale
def test_properties_at_config_C2_size(eng_mod):
    """V=10000 S=64 G=8 (BASELINE config C2): conservation laws and run-to-run determinism."""
    p = synth_problem(10000, 64, 8, depth=100.0, seed=20240611)
    outs = []
    for rep in range(2):
        e = eng_mod.Engine(0, seed=23724839)
        e.set_counts(p["counts"])
        e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
        sm, es = e.mu_stats()
        assert sm.sum() == p["counts"].sum()
        assert np.array_equal(es.sum(1), p["counts"].sum((0, 1)))       # Esum rows = observed-base totals
        assert np.array_equal(sm.sum(1), p["counts"].sum((0, 2)))       # sum_mu rows = per-sample depth
        out = e.update(6)
        tau = e.get_tau_index()
        ts = e.get_tau_sum()
        assert (ts.sum(2) == 6).all()
        assert np.allclose(out["gamma_store"].sum(2), 1.0, atol=1e-12)
        assert np.allclose(out["eta_store"].sum(2), 1.0, atol=1e-12)
        assert out["lp_store"][-1] > out["lp_store"][0]
        outs.append((tau, out["gamma_store"], out["ll_store"], out["nchange"]))
        e.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)                                     # bitwise reproducible


# ------------------------------------------------------------------ filtered-exact fast path
@pytest.mark.parametrize("V,S,G,depth", [(400, 64, 8, 100.0), (300, 64, 8, 3.0), (200, 130, 12, 10.0), (150, 7, 3, 8.0),
                                          (64, 256, 16, 30.0)])
def test_tau_fast_path_draws_equal_fp64_reference_order_path(eng_mod, oracle_mod, V, S, G, depth):
    """tau_exact=1 evaluates every draw with the reference's FP64 arithmetic; the default filtered path must
    produce the identical tau (and both equal the oracle).  Tier counters prove the vectors exercise the FP32
    gap test AND the FP64 bracket test (ambiguous draws), not only deterministic ones."""
    p = synth_problem(V, S, G, depth=depth, seed=7 * V + G, ambiguous=True)
    rng = np.random.default_rng(V)
    taus = {}
    tiers = None
    for mode in (0, 1):
        e = eng_mod.Engine(0, seed=99)
        e.set_option("tau_exact", mode)
        e.set_counts(p["counts"])
        e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
        rng = np.random.default_rng(V)
        hist = []
        e.get_tier_counts()
        for k in range(6):
            gamma = rng.dirichlet(np.full(G, 0.2 if k % 2 else 1.0), size=S)
            gamma[gamma < 1e-6] = 1e-6
            gamma /= gamma.sum(1)[:, None]
            eta = rng.dirichlet(np.array([200.0, 0.3, 0.3, 0.3]), size=4)       # rough, tiny off-diagonals included
            eta = np.array([np.roll(eta[a], a) for a in range(4)])
            eta = np.maximum(eta, 1e-30); eta /= eta.sum(1)[:, None]
            e.set_state(None, gamma, eta, G=G)
            e.set_rng(99, sweep=k)
            e.sample_tau()
            hist.append(e.get_tau_index())
        taus[mode] = hist
        if mode == 0:
            tiers = e.get_tier_counts()
        e.close()
    for a, b in zip(taus[0], taus[1]):
        assert np.array_equal(a, b)
    assert tiers.sum() == 6 * V * G
    assert tiers[0] > 0 and tiers[1] > 0, tiers
    # and the oracle agrees on the last state
    tau_o = onehot(p["tau0"])
    rng = np.random.default_rng(V)
    for k in range(6):
        gamma = rng.dirichlet(np.full(G, 0.2 if k % 2 else 1.0), size=S)
        gamma[gamma < 1e-6] = 1e-6
        gamma /= gamma.sum(1)[:, None]
        eta = rng.dirichlet(np.array([200.0, 0.3, 0.3, 0.3]), size=4)
        eta = np.array([np.roll(eta[a], a) for a in range(4)])
        eta = np.maximum(eta, 1e-30); eta /= eta.sum(1)[:, None]
        oracle_mod.sample_tau_philox(tau_o, gamma, eta, p["counts"], 99, k)
    assert np.array_equal(np.argmax(tau_o, 2).astype(np.uint8), taus[0][-1])


def test_tau_fast_path_long_chain_equals_exact_at_C2_size(eng_mod):
    """V=10000 S=64 G=8: 30 full sweeps from a random start, filtered path vs FP64 path: same tau, same nchange trace."""
    p = synth_problem(10000, 64, 8, depth=100.0, seed=20240611)
    res = {}
    for mode in (0, 1):
        e = eng_mod.Engine(0, seed=23724839)
        e.set_option("tau_exact", mode)
        e.set_counts(p["counts"])
        e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
        out = e.update(30)
        res[mode] = (e.get_tau_index(), out["nchange"], out["ll_store"], e.get_tier_counts())
        e.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert np.array_equal(res[0][2], res[1][2])
    t = res[0][3]
    assert t[2] < 0.01 * t.sum()            # the FP64 recompute is the rare path


# ------------------------------------------------------------------ round-2 advisor findings
def test_dropin_masked_gamma_columns_vs_reference_C(oracle_mod):
    """Eta_Sampler.sampleTauC hands sample_tau a gamma whose columns are exactly 0.0 for the strains a gene lacks
    (maskGamma, Eta_Sampler.py:147-157,367).  The reference C accepts that; so must the drop-in, with identical draws
    (the step of a masked strain is a uniform draw over the four bases: all candidates leave the mixture unchanged)."""
    import ctypes as C
    from desman_b200 import _lib, sampletau
    V, S, G = 400, 40, 6
    p = synth_problem(V, S, G, depth=25.0, seed=77, ambiguous=True)
    rng = np.random.default_rng(3)
    use_ref = oracle_mod.have_ref()
    seed = 99
    tau_g, tau_r = onehot(p["tau0"]), onehot(p["tau0"])
    if use_ref:
        R = oracle_mod.RefSampleTau(seed)
    else:
        st = oracle_mod.MT19937()
        oracle_mod.lib().oracle_mt_seed(C.byref(st), seed)
    sampletau.initRNG(); sampletau.setRNG(seed)
    for k, masked in enumerate([(1,), (0, 4), (2, 3, 5), ()]):
        gamma = rng.dirichlet(np.ones(G), size=S)
        gamma[:, list(masked)] = 0.0
        gamma /= gamma.sum(1)[:, None]
        n_gpu = sampletau.sample_tau(tau_g, gamma, p["eta0"], p["counts"])
        if use_ref:
            n_ref = R.sample_tau(tau_r, gamma, p["eta0"], p["counts"])
        else:
            n_ref = oracle_mod.lib().oracle_sample_tau_mt(
                tau_r.ctypes.data_as(oracle_mod._p64), oracle_mod._f64(gamma)[1], oracle_mod._f64(p["eta0"])[1],
                oracle_mod._i64(p["counts"])[1], V, G, S, C.byref(st))
        assert n_gpu == n_ref and np.array_equal(tau_g, tau_r), (k, masked)
    # a sample whose whole mixture is zero stays an error (log 0 in the reference)
    gamma = rng.dirichlet(np.ones(G), size=S); gamma[3, :] = 0.0
    with pytest.raises(_lib.DesmanB200Error, match="no positive entry"):
        sampletau.sample_tau(tau_g, gamma, p["eta0"], p["counts"])
    sampletau.freeRNG()
    if use_ref:
        R.close()


def test_sampler_keeps_both_streams_across_rng_mode_switches(oracle_mod):
    """tau_rng='mt19937': sampleTau (MT19937 words) interleaved with sampleMu / sampleGamma (Philox counters) on ONE sampler.
    Every sampleTau must consume the NEXT V*G words of the GSL stream -- not restart it -- and the context is not re-created."""
    from numpy.random import RandomState
    from desman_b200 import sampletau
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    V, S, G = 300, 24, 4
    p = synth_problem(V, S, G, depth=6.0, seed=12, ambiguous=True)
    seed = 4242
    sampletau.initRNG(); sampletau.setRNG(seed)
    hs = HaploSNP_Sampler(p["counts"], G, RandomState(1), max_iter=2, tau_rng="mt19937")
    hs.tau = onehot(p["tau0"]); hs.gamma = p["gamma0"].copy()
    tau_o = onehot(p["tau0"])
    words = oracle_mod.mt_words(seed, 3 * V * G)
    eng_ids = []
    for k in range(3):
        n = hs.sampleTau()
        eng_ids.append(id(hs._eng))
        n_o = oracle_mod.sample_tau_words(tau_o, hs.gamma, hs.eta, p["counts"], words[k * V * G:(k + 1) * V * G])
        assert n == n_o and np.array_equal(hs.tau, tau_o), k
        sm, es = hs.sampleMu(hs.tau, hs.gamma, hs.eta)         # Philox-mode calls in between
        assert sm.sum() == p["counts"].sum()
        hs.sampleGamma(); hs.sampleEta()
    assert len(set(eng_ids)) == 1
    hs.close()
    sampletau.freeRNG()


def test_batched_small_calls_replay_an_eta_sampler_sequence(oracle_mod):
    """Eta_Sampler.calcTauStar-style traffic (Eta_Sampler.py:430-446): every iteration calls sample_tau once per gene, each with
    its own masked gamma (maskGamma, :147-157).  sampletau.Batch keeps the counts of all genes on the device and runs ONE launch
    per iteration; tau and nchange of every gene must equal the sequence of calls to the reference's own C (oracle/_ref; the
    oracle's restatement where that library did not travel), through several iterations of the one MT19937 stream."""
    import ctypes as C
    from desman_b200 import sampletau
    rng = np.random.default_rng(11)
    G, S, seed, n_iter = 5, 24, 20240611, 4
    sizes = [37, 1, 260, 8, 96, 15, 530, 64]                       # genes of very different lengths
    genes = []
    for k, V in enumerate(sizes):
        p = synth_problem(V, S, G, depth=12.0, seed=100 + k, ambiguous=True)
        genes.append(dict(counts=p["counts"], tau=onehot(p["tau0"]), tau_ref=onehot(p["tau0"])))
    gamma = rng.dirichlet(np.ones(G), size=S)
    eta_gene = rng.random((len(sizes), G)) < 0.7                   # presence / absence of every strain in every gene
    eta_gene[:, 0] = True
    eps = 0.96 * np.identity(4) + 0.01
    use_ref = oracle_mod.have_ref()
    if use_ref:
        R = oracle_mod.RefSampleTau(seed)
    else:
        st = oracle_mod.MT19937()
        oracle_mod.lib().oracle_mt_seed(C.byref(st), seed)
    sampletau.initRNG(); sampletau.setRNG(seed)
    b = sampletau.Batch([g["counts"] for g in genes])
    for it in range(n_iter):
        pis = []
        for k, g in enumerate(genes):
            gR = gamma.copy()
            gR[:, ~eta_gene[k]] = 0.0
            gR /= gR.sum(1)[:, None]
            pis.append(np.ascontiguousarray(gR))
        n_gpu = b.sample_tau([g["tau"] for g in genes], pis, eps)
        for k, g in enumerate(genes):
            if use_ref:
                n_ref = R.sample_tau(g["tau_ref"], pis[k], eps, g["counts"])
            else:
                n_ref = oracle_mod.lib().oracle_sample_tau_mt(
                    g["tau_ref"].ctypes.data_as(oracle_mod._p64), oracle_mod._f64(pis[k])[1], oracle_mod._f64(eps)[1],
                    oracle_mod._i64(g["counts"])[1], sizes[k], G, S, C.byref(st))
            assert n_gpu[k] == n_ref, (it, k)
            assert np.array_equal(g["tau"], g["tau_ref"]), (it, k)
        gamma = rng.dirichlet(np.ones(G), size=S)
    # the stream goes on where the batch left it: a plain call afterwards still matches
    t1, t2 = genes[2]["tau"], genes[2]["tau_ref"]
    n1 = sampletau.sample_tau(t1, gamma, eps, genes[2]["counts"])
    if use_ref:
        n2 = R.sample_tau(t2, gamma, eps, genes[2]["counts"])
        R.close()
    else:
        n2 = oracle_mod.lib().oracle_sample_tau_mt(t2.ctypes.data_as(oracle_mod._p64), oracle_mod._f64(gamma)[1], oracle_mod._f64(eps)[1],
                                                   oracle_mod._i64(genes[2]["counts"])[1], sizes[2], G, S, C.byref(st))
    assert n1 == n2 and np.array_equal(t1, t2)
    b.close()
    sampletau.freeRNG()
