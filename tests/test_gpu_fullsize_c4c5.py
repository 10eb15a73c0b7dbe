"""GPU tests at the full sizes of BASELINE configs C4 (V=100000, S=256, G=16) and C5 (one 8-GPU shard of V=1e6, S=128,
G=20: 125000 sites), through size-independent properties: read conservation of the mu/E statistics, run-to-run determinism
of a short chain, occupancy accounting, and the device log-likelihood against a direct numpy evaluation (1e-9 relative; the
north star asks for 1e-6).  With G >= 16 the patterns are (nearly) all distinct, so these shapes exercise the per-read
statistics kernel and the ungrouped per-site tau kernel -- the paths the C3 tests do not reach at size."""
import numpy as np
import pytest
from scipy.special import gammaln

from conftest import onehot

pytestmark = pytest.mark.gpu

SHAPES = {"C4": (100000, 256, 16), "C5_shard": (125000, 128, 20)}
N_SWEEPS = 3


@pytest.fixture(scope="module", params=sorted(SHAPES))
def case(request):
    from desman_b200 import _lib, engine
    from desman_b200.synth import synth_counts
    assert _lib.device_count() >= 1
    V, S, G = SHAPES[request.param]
    return dict(name=request.param, V=V, S=S, G=G, p=synth_counts(V, S, G), engine=engine)


def host_loglik(counts, tau_idx, gamma, eta):
    """HaploSNP_Sampler.logLikelihood (:431-442) + Desman_Utils.log_multinomial_pdf (:28-33), in blocks of sites."""
    tot = 0.0
    for lo in range(0, counts.shape[0], 5000):
        c = counts[lo:lo + 5000]
        p = np.einsum("sg,vga->vsa", gamma, eta[tau_idx[lo:lo + 5000]])
        tot += float((c * np.log(p)).sum() + (gammaln(c.sum(2) + 1.0) - gammaln(c + 1.0).sum(2)).sum())
    return tot


def test_statistics_conserve_every_read(case):
    p, eng = case["p"], case["engine"]
    e = eng.Engine(0, seed=11)
    e.set_counts(p["counts"])
    e.set_state(onehot(p["tau_true"]), p["gamma_true"], p["eta0"])
    sm, es = e.mu_stats()
    e.close()
    total = p["counts"].sum()
    assert sm.sum() == total and es.sum() == total
    assert np.array_equal(sm.sum(1), p["counts"].sum((0, 2)))                    # depth of every sample
    assert np.array_equal(es.sum(1), p["counts"].sum((0, 1)))                    # reads observed as each base


def test_short_chain_is_deterministic_and_consistent(case):
    """Two engines, same seed: identical tau, nchange, ll, gamma (the statistics are integer sums flushed with atomics and the
    log-likelihood is accumulated in fixed point, so scheduling cannot show).  The last ll of the chain equals the numpy
    evaluation at the final state, and every (v,g) occupies exactly one base per sweep."""
    p, eng, V, G = case["p"], case["engine"], case["V"], case["G"]
    runs = []
    for _ in range(2):
        e = eng.Engine(0, seed=23724839)
        e.set_counts(p["counts"])
        e.set_state(onehot(p["tau0"]), p["gamma0"], p["eta0"])
        out = e.update(N_SWEEPS)
        _, gamma, eta = e.get_state(want_tau=False)
        runs.append(dict(tau=e.get_tau_index(), nchange=out["nchange"], ll=out["ll_store"], gamma=gamma, eta=eta,
                         tau_sum=e.get_tau_sum(compact=True), ll_now=e.loglik()[0]))
        e.close()
    a, b = runs
    for k in ("tau", "nchange", "ll", "gamma", "eta", "tau_sum"):
        assert np.array_equal(a[k], b[k]), k
    assert a["tau_sum"].sum() == N_SWEEPS * V * G
    assert a["nchange"][0] > V                                                    # a random start: most sites move
    want = host_loglik(p["counts"], a["tau"], a["gamma"], a["eta"])
    assert abs(a["ll"][-1] - want) <= 1e-9 * abs(want)
    assert abs(a["ll_now"] - want) <= 1e-9 * abs(want)
