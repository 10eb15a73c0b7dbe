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
