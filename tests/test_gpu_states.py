"""Joint-state enumeration on the device: desman_state_logprob / desman_loglik_general through the class
surface (assignTau, logTauProb, DIC) against the golden vectors of the UNMODIFIED reference class and against the numpy oracle."""
import os

import numpy as np
import pytest
from numpy.random import RandomState

from conftest import GOLDEN, onehot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLDEN, "assign_kat.npz"))


def test_class_surface_matches_reference_golden(kat):
    """assignTau draws the reference's states from the same RandomState and leaves its stream where the reference does; conf,
    logTauProb and logLikelihood of a real-valued tau agree to 1e-9 relative (north star: 1e-6)."""
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    for ci in range(int(kat["ncases"])):
        counts, newc = kat[f"c{ci}_counts"].astype(np.int64), kat[f"c{ci}_new"].astype(np.int64)
        gamma, eta = kat[f"c{ci}_gamma"], kat[f"c{ci}_eta"]
        G = gamma.shape[1]
        h = HaploSNP_Sampler(counts, G, RandomState(3), max_iter=3)
        h.gamma_star, h.eta_star = gamma.copy(), eta.copy()
        h.tauIndices_star = kat[f"c{ci}_star"].copy()
        want = float(kat[f"c{ci}_logtauprob"])
        assert abs(h.logTauProb(gamma, eta) - want) <= 1e-9 * abs(want)
        h.randomState = RandomState(int(kat[f"c{ci}_rng_seed"]))
        aT, conf = h.assignTau(np.reshape(newc, (newc.shape[0], -1)))
        assert aT.shape == (newc.shape[0], G, 4) and aT.dtype == np.int64 and (aT.sum(2) == 1).all()
        assert np.array_equal(np.argmax(aT, axis=2), kat[f"c{ci}_assign"])
        assert np.allclose(conf, kat[f"c{ci}_conf"], rtol=1e-9, atol=0)
        assert np.array_equal(h.randomState.random_sample(2), kat[f"c{ci}_rng_after"])
        want = float(kat[f"c{ci}_ll_real"])
        assert abs(h.logLikelihood(gamma, kat[f"c{ci}_tau_real"], eta) - want) <= 1e-9 * abs(want)
        h.close()


@pytest.mark.parametrize("N,S,G", [(70, 5, 1), (130, 64, 3), (200, 7, 4), (65, 33, 6)])
def test_state_logprob_vs_oracle(N, S, G):
    """All 4^G log-probabilities, their log-sum-exp, maximum and first argmax per site, with and without an uploaded tensor,
    ragged against the 64x64 tiles of the kernel; an all-zero site gives the uniform distribution."""
    from desman_b200 import engine
    from desman_b200.synth import synth_counts
    from oracle import oracle
    p = synth_counts(N, S, G, depth=12.0, seed=100 + G)
    counts = p["counts"].copy()
    counts[N // 2] = 0
    want = oracle.state_logprob(counts, p["gamma_true"], p["eta0"])
    idx = np.random.default_rng(G).integers(0, 4 ** G, size=N)
    e = engine.Engine(0, seed=1)
    e.set_counts(counts)
    for variants in (None, counts):
        r = e.state_logprob(p["gamma_true"], p["eta0"], variants=variants, index=idx, want_logprob=True)
        assert np.allclose(r["logprob"], want, rtol=1e-11, atol=1e-9)
        assert np.allclose(r["maxlp"], want.max(1), rtol=1e-11, atol=1e-9)
        m = want.max(1)
        assert np.allclose(r["lse"], m + np.log(np.exp(want - m[:, None]).sum(1)), rtol=1e-11, atol=1e-9)
        assert np.allclose(r["lp_at_index"], want[np.arange(N), idx], rtol=1e-11, atol=1e-9)
        assert np.allclose(want[np.arange(N), r["argmax"]], want.max(1), rtol=0, atol=1e-9)
        assert abs(r["lse"][N // 2] - np.log(4.0 ** G)) < 1e-12 and r["argmax"][N // 2] == 0
    e.close()


def test_state_logprob_rejects_what_cannot_fit():
    from desman_b200 import _lib, engine
    e = engine.Engine(0, seed=1)
    e.set_counts(np.ones((4, 64, 4), dtype=np.int64))
    with pytest.raises(_lib.DesmanB200Error, match="do not fit"):
        e.state_logprob(np.full((64, 14), 1.0 / 14), np.full((4, 4), 0.25))
    with pytest.raises(_lib.DesmanB200Error, match="not a state"):
        e.state_logprob(np.full((64, 2), 0.5), np.full((4, 4), 0.25), index=np.array([0, 1, 16, 2]))
    e.close()


def test_dic_from_a_short_chain():
    """DIC = meanDeviance + 2 logLikelihood(gammaMean, tauMean, etaMean) (:486-496) on a chain of the engine, against the numpy
    oracle evaluated at the same means."""
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    from desman_b200.synth import synth_counts
    from oracle import oracle
    p = synth_counts(300, 16, 3, depth=15.0, seed=5)
    h = HaploSNP_Sampler(p["counts"], 3, RandomState(1), max_iter=6, seed=7)
    h.update()
    want = h.meanDeviance() + 2.0 * oracle.loglik_general(p["counts"], h.tauMean(), h.gammaMean(), h.etaMean())
    assert abs(h.DIC() - want) <= 1e-9 * abs(want)
    h.close()


# ------------------------------------------------------------------ sampleTauFixTau and Chib's marginal likelihood (:196-222, :538-710)
def _log_dir(x, alpha):
    from scipy.special import gammaln
    return float(gammaln(np.sum(alpha)) + np.sum((alpha - 1.0) * np.log(x)) - np.sum(gammaln(alpha)))


def _log_mean(x):
    m = np.max(x)
    return float(m + np.log(np.exp(x - m).sum()) - np.log(len(x)))


@pytest.mark.parametrize("V,S,G,H", [(200, 24, 4, 0), (150, 64, 5, 2), (90, 7, 3, 2), (64, 130, 8, 5)])
def test_sample_tau_fix_tau_vs_oracle(oracle_mod, V, S, G, H):
    """sampleTauFixTau: strains below H untouched, strains [H, G) redrawn in order, log-probabilities of strain H recorded --
    tau bit-exact, log-probabilities 1e-10 against the oracle's restatement under the same Philox counters."""
    from desman_b200 import sampletau
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    from desman_b200.synth import synth_counts
    p = synth_counts(V, S, G, depth=6.0, seed=40 + G)
    seed = 991
    sampletau.initRNG(); sampletau.setRNG(seed)
    h = HaploSNP_Sampler(p["counts"], G, RandomState(1), max_iter=2, seed=seed)
    work, work_o = onehot(p["tau0"]), onehot(p["tau0"])
    for k in range(3):
        sw = h._engine().get_rng()[0]
        lp = h.sampleTauFixTau(work, H, p["gamma_true"], p["eta0"])
        lp_o, _ = oracle_mod.sample_tau_fix_philox(work_o, H, p["gamma_true"], p["eta0"], p["counts"], seed, sw)
        assert np.array_equal(work, work_o), k
        assert np.allclose(lp, lp_o, rtol=1e-10, atol=1e-12), k
        assert np.array_equal(np.argmax(work[:, :H], 2), p["tau0"][:, :H])
        assert np.allclose(np.exp(lp).sum(1), 1.0, atol=1e-9)
    assert not np.array_equal(np.argmax(work, 2), p["tau0"])
    h.close()
    sampletau.freeRNG()


def test_chib_marginal_likelihoods_vs_oracle_composition(oracle_mod):
    """chibMarginalLogLikelihood / chibMarginalLogLikelihood2 (:538-710) after a short chain: the same compositions written with the
    CPU oracle's primitives under the same counters must give the same estimates (1e-9 relative)."""
    from desman_b200 import sampletau
    from desman_b200.engine import auto_mu_mode
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    from desman_b200.synth import synth_counts
    V, S, G, n_iter, seed = 120, 12, 3, 10, 4711
    p = synth_counts(V, S, G, depth=10.0, seed=9)
    counts = p["counts"]
    mode = auto_mu_mode(V, G)
    alpha, delta = np.full(G, 0.1), np.full(4, 0.1)
    sampletau.initRNG(); sampletau.setRNG(seed)
    h = HaploSNP_Sampler(counts, G, RandomState(1), max_iter=n_iter, seed=seed)
    h.tau, h.gamma, h.eta = onehot(p["tau0"]), p["gamma0"].copy(), p["eta0"].copy()
    h.update()
    want = oracle_mod.update(onehot(p["tau0"]), p["gamma0"], p["eta0"], counts, n_iter, seed, mu_mode=mode)
    assert np.array_equal(np.argmax(h.tau, 2), np.argmax(want["tau"], 2))
    tau_star, gamma_star, eta_star = want["tau_star"], want["gamma_star"], want["eta_star"]
    star_idx = np.argmax(tau_star, 2)
    star_code = np.zeros(V, dtype=np.int64)
    for g in range(G):
        star_code = star_code * 4 + star_idx[:, g]
    base = oracle_mod.loglik(tau_star, gamma_star, eta_star, counts)
    lgp = sum(_log_dir(gamma_star[s], alpha) for s in range(S))
    lep = sum(_log_dir(eta_star[a], delta) for a in range(4))

    # ---- chibMarginalLogLikelihood (:621-710)
    got1 = h.chibMarginalLogLikelihood()
    sw = n_iter
    gamma, eta = want["gamma"].copy(), want["eta"].copy()
    ltau = _log_mean(np.array([oracle_mod.log_tau_prob(counts, want["gamma_store"][i], want["eta_store"][i], star_code)
                               for i in range(10)]))
    lg = np.zeros(n_iter)
    for i in range(n_iter):
        sm, es = oracle_mod.mu_stats(tau_star, gamma, eta, counts, seed, sw, mode=mode)
        gamma = oracle_mod.draw_gamma(sm, 0.1, 1e-6, seed, sw)
        eta = oracle_mod.draw_eta(es, 0.1, seed, sw)
        sw += 1
        lg[i] = sum(_log_dir(gamma_star[s], alpha + sm[s]) for s in range(S))
    le = np.zeros(n_iter)
    for i in range(n_iter):
        sm, es = oracle_mod.mu_stats(tau_star, gamma_star, eta, counts, seed, sw, mode=mode)
        eta = oracle_mod.draw_eta(es, 0.1, seed, sw)
        sw += 1
        le[i] = sum(_log_dir(eta_star[a], delta + es[:, a]) for a in range(4))
    want1 = base + lgp + lep + V * np.log(1.0 / 4.0 ** G) - _log_mean(lg) - _log_mean(le) - ltau
    assert abs(got1 - want1) <= 1e-9 * abs(want1), (got1, want1)

    # ---- chibMarginalLogLikelihood2 (:538-608), continuing the same streams
    got2 = h.chibMarginalLogLikelihood2()
    # eta term from the per-sweep Esum of the chain: recompute the chain's statistics sweep by sweep
    tau_c, gamma_c, eta_c = onehot(p["tau0"]), p["gamma0"].copy(), p["eta0"].copy()
    le2 = np.zeros(n_iter)
    for i in range(n_iter):
        sm, es = oracle_mod.mu_stats(tau_c, gamma_c, eta_c, counts, seed, i, mode=mode)
        gamma_c = oracle_mod.draw_gamma(sm, 0.1, 1e-6, seed, i)
        oracle_mod.sample_tau_philox(tau_c, gamma_c, eta_c, counts, seed, i)
        eta_c = oracle_mod.draw_eta(es, 0.1, seed, i)
        le2[i] = sum(_log_dir(eta_star[a], delta + es[:, a]) for a in range(4))
    tau_w = want["tau"].copy()
    lg2 = np.zeros(n_iter)
    for i in range(n_iter):
        oracle_mod.sample_tau_philox(tau_w, gamma, eta_star, counts, seed, sw)
        sw += 1                                                    # the tau draw moves the sweep counter on
        sm, es = oracle_mod.mu_stats(tau_w, gamma, eta_star, counts, seed, sw, mode=mode)
        gamma = oracle_mod.draw_gamma(sm, 0.1, 1e-6, seed, sw)
        lg2[i] = sum(_log_dir(gamma_star[s], alpha + sm[s]) for s in range(S))
    ltau2 = 0.0
    for hh in range(G):
        work = tau_star.copy()
        st = np.zeros(n_iter)
        for i in range(n_iter):
            lp, _ = oracle_mod.sample_tau_fix_philox(work, hh, gamma_star, eta_star, counts, seed, sw)
            sw += 1
            st[i] = lp[np.arange(V), star_idx[:, hh]].sum()
        ltau2 += _log_mean(st)
    want2 = base + lep - _log_mean(le2) + lgp - _log_mean(lg2) + V * G * np.log(0.25) - ltau2
    assert abs(got2 - want2) <= 1e-9 * abs(want2), (got2, want2)
    h.close()
    sampletau.freeRNG()
