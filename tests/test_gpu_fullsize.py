"""GPU tests at the full size of BASELINE config C3 (V=100000, S=64, G=8), through properties that do not need the
oracle to finish a chain of that size: read conservation of the statistics, chains bit-identical with and without the
pattern-grouped tau update, tier accounting, and the table-based log-likelihood against a direct numpy evaluation."""
import numpy as np
import pytest

from conftest import onehot

pytestmark = pytest.mark.gpu

V, S, G = 100000, 64, 8


@pytest.fixture(scope="module")
def problem():
    from desman_b200.synth import synth_counts
    return synth_counts(V, S, G)


@pytest.fixture(scope="module")
def eng_mod():
    from desman_b200 import _lib, engine
    assert _lib.device_count() >= 1
    return engine


def host_loglik_terms(counts, tau_idx, gamma, eta):
    """sum n*log p of HaploSNP_Sampler.logLikelihood (:435,:441) without the multinomial coefficient, in blocks."""
    tot = 0.0
    for lo in range(0, counts.shape[0], 10000):
        p = np.einsum("sg,vga->vsa", gamma, eta[tau_idx[lo:lo + 10000]])
        tot += float((counts[lo:lo + 10000] * np.log(p)).sum())
    return tot


@pytest.mark.parametrize("mode", [0, 1])
def test_statistics_conserve_every_read_at_C3(eng_mod, problem, mode):
    """Every read is assigned to exactly one strain and one true base: sum_mu and Esum both add up to the read total, and
    the row sums of sum_mu are the per-sample depths -- for the per-read and the pattern-aggregated contract alike."""
    p = problem
    e = eng_mod.Engine(0, seed=11)
    e.set_option("mu_mode", mode)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau_true"]), p["gamma_true"], p["eta0"])
    sm, es = e.mu_stats()
    e.close()
    assert sm.sum() == p["counts"].sum() and es.sum() == p["counts"].sum()
    assert np.array_equal(sm.sum(1), p["counts"].sum((0, 2)))                    # depth of every sample
    assert np.array_equal(es.sum(1), p["counts"].sum((0, 1)))                    # reads observed as each base


def test_grouped_and_per_site_chains_identical_at_C3(eng_mod, problem):
    """12 full sweeps from the random start of the benchmark (burn-in, table rebuilds, regroup, then the screening pass on
    ~3 % of the sites): tau, nchange and the log-likelihood trace must not depend on whether the sites are grouped."""
    p = problem
    res = {}
    for group in (1, 0):
        e = eng_mod.Engine(0, seed=23724839)
        e.set_option("tau_group", group)
        e.set_counts(p["counts"])
        e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
        e.get_tier_counts()
        out = e.update(12)
        res[group] = dict(tau=e.get_tau_index(), nchange=out["nchange"], ll=out["ll_store"], gamma=e.get_state(want_tau=False)[1],
                          tiers=e.get_tier_counts(), stats=e.get_group_stats(), tau_sum=e.get_tau_sum(compact=True))
        if group == 1:
            ll_dev, _ = e.loglik()
            ll_const = ll_dev - host_loglik_terms(p["counts"], res[1]["tau"], *e.get_state(want_tau=False)[1:])
            res["ll_split"] = (ll_dev, ll_const)
        e.close()
    a, b = res[1], res[0]
    for k in ("tau", "nchange", "ll", "gamma", "tau_sum"):
        assert np.array_equal(a[k], b[k]), k
    assert a["tiers"].sum() == 12 * V * G and b["tiers"].sum() == 12 * V * G
    assert a["nchange"][0] > V and a["nchange"][-1] < V // 100                  # a burn-in, then a calm chain
    st = a["stats"]
    assert st["have"] == 1 and st["calm"] == 1 and 0 < st["work"] + st["singles"] < V // 10, st
    assert a["tau_sum"].sum() == 12 * V * G                                       # every (v,g) occupies one base per sweep
    # the multinomial-coefficient constant recovered from the device value must be what lgamma gives on the host
    from scipy.special import gammaln
    c = p["counts"]
    want_const = float((gammaln(c.sum(2) + 1.0) - gammaln(c + 1.0).sum(2)).sum())
    assert abs(res["ll_split"][1] - want_const) <= 1e-9 * abs(res["ll_split"][0])
