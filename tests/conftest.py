import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def onehot(idx):
    """uint8 [V,G] base indices -> int64 one-hot [V,G,4] (reference tau layout)."""
    idx = np.asarray(idx)
    out = np.zeros(idx.shape + (4,), dtype=np.int64)
    np.put_along_axis(out, idx[..., None].astype(np.int64), 1, axis=-1)
    return out


def synth_problem(V, S, G, depth=100.0, seed=20240611, ambiguous=False):
    """SURVEY.md section 8d synthetic generator (host side, numpy default_rng)."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, size=V)
    alt = (anc + rng.integers(1, 4, size=V)) % 4
    carry = rng.random((V, G)) < 0.25
    for v in range(V):
        while carry[v].all() or not carry[v].any():
            carry[v] = rng.random(G) < 0.25
            if G == 1:
                break
    tau_true = np.where(carry, alt[:, None], anc[:, None])
    gamma_true = rng.dirichlet(np.ones(G), size=S)
    if ambiguous and G > 2:
        gamma_true[:, :2] *= 0.01
        gamma_true /= gamma_true.sum(1)[:, None]
    eta_true = 0.997 * np.identity(4) + 0.001 * (1 - np.identity(4))
    p = np.einsum("sg,vga->vsa", gamma_true, eta_true[tau_true])
    p = p / p.sum(-1, keepdims=True)
    N = rng.poisson(depth, size=(V, S))
    counts = np.zeros((V, S, 4), dtype=np.int64)
    for v in range(V):
        counts[v] = rng.multinomial(N[v], p[v])
    tau0 = rng.integers(0, 4, size=(V, G)).astype(np.uint8)
    gamma0 = rng.dirichlet(np.full(G, 1.0), size=S)
    gamma0[gamma0 < 1e-6] = 1e-6
    gamma0 /= gamma0.sum(1)[:, None]
    eta0 = 0.96 * np.identity(4) + 0.01 * np.ones((4, 4))
    return dict(counts=counts, tau0=tau0, gamma0=gamma0, eta0=eta0, tau_true=tau_true, gamma_true=gamma_true)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle
