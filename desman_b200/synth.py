"""Synthetic variant-count tensors of SURVEY.md section 8d (host side, numpy; not part of the engine).

Biallelic sites, G_true = G strains with Dirichlet(1) abundances per sample, depth ~ Poisson(100),
eta_true = 0.997 I + 0.001 (1 - I).  Used by bench.py and by the parity tests at BASELINE sizes.
"""
import numpy as np

DATA_SEED = 20240611
CHAIN_SEED = 23724839  # reference default, bin/desman:55


def synth_counts(V, S, G, depth=100.0, seed=DATA_SEED, shard=0):
    """Counts [V,S,4] int64 plus a shared initial state.  `shard` selects an independent block of V
    sites of the same community: gamma_true, gamma0 and eta0 depend on (seed, S, G) only, so the
    shards of a multi-GPU run are slices of one consistent V_total = nshards*V problem."""
    rng = np.random.default_rng([seed, shard])
    shared = np.random.default_rng([seed, 0x5EED])
    anc = rng.integers(0, 4, size=V)
    alt = (anc + rng.integers(1, 4, size=V)) % 4
    carry = rng.random((V, G)) < 0.25
    if G > 1:
        for _ in range(64):  # resample until both alleles are present at every site
            bad = carry.all(1) | ~carry.any(1)
            if not bad.any():
                break
            carry[bad] = rng.random((int(bad.sum()), G)) < 0.25
        bad = carry.all(1) | ~carry.any(1)
        carry[bad, 0] = ~carry[bad, 0]
    tau_true = np.where(carry, alt[:, None], anc[:, None]).astype(np.uint8)
    gamma_true = shared.dirichlet(np.ones(G), size=S)
    eta_true = 0.997 * np.identity(4) + 0.001 * (1.0 - np.identity(4))
    counts = np.empty((V, S, 4), dtype=np.int64)
    step = max(1, (1 << 22) // max(S * G, 1))
    for lo in range(0, V, step):
        hi = min(V, lo + step)
        p = np.einsum("sg,vga->vsa", gamma_true, eta_true[tau_true[lo:hi]])
        p /= p.sum(-1, keepdims=True)
        N = rng.poisson(depth, size=(hi - lo, S))
        counts[lo:hi] = rng.multinomial(N, p)
    tau0 = rng.integers(0, 4, size=(V, G)).astype(np.uint8)
    gamma0 = shared.dirichlet(np.full(G, 1.0), size=S)
    gamma0[gamma0 < 1e-6] = 1e-6
    gamma0 /= gamma0.sum(1)[:, None]
    eta0 = 0.96 * np.identity(4) + 0.01 * np.ones((4, 4))
    return dict(counts=counts, tau0=tau0, gamma0=gamma0, eta0=eta0, tau_true=tau_true, gamma_true=gamma_true)


def onehot(idx):
    """uint8 base indices [V,G] -> int64 one-hot [V,G,4] (the reference's tau layout)."""
    idx = np.asarray(idx)
    out = np.zeros(idx.shape + (4,), dtype=np.int64)
    np.put_along_axis(out, idx[..., None].astype(np.int64), 1, axis=-1)
    return out
