"""Developer probe (CPU, uses the oracle: test infrastructure): which sampler regime do the class-split draws of the aggregated
mu/E statistics fall into at steady state (mu_agg_kernel.cuh: inversion when n*min(p,q) < 10, BTRS otherwise), per observed base?
Input for the 'separate dense passes per regime' item of DESIGN.md section 7."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from desman_b200.synth import onehot, synth_counts  # noqa: E402
from oracle import oracle  # noqa: E402

V, S, G = 100000, 64, 8
p = synth_counts(V, S, G)
# a converged state without running the chain: the generating tau and gamma (the chain sits next to them at this depth)
tau, gamma = p["tau_true"].astype(np.int64), p["gamma_true"]
eta = 0.997 * np.identity(4) + 0.001 * (1.0 - np.identity(4))
codes = (tau << (2 * np.arange(G))).sum(1)
uniq, inv = np.unique(codes, return_inverse=True)
P = len(uniq)
N = np.zeros((P, S, 4), dtype=np.int64)
np.add.at(N, inv, p["counts"])
pat = np.zeros((P, G), dtype=np.int64)
pat[inv] = tau
print("patterns", P, "sites", V)
tot = dict(zero=0, trivial=0, inversion=0, btrs=0)
steps = []
for a in range(4):
    # classes of a pattern: strains grouped by base; weights W_b = eta[b,a] * Gamma_b
    Gam = np.stack([(gamma[None, :, :] * (pat[:, None, :] == b)).sum(2) for b in range(4)], axis=2)     # [P,S,4]
    W = Gam * eta[None, None, :, a]
    present = (pat[:, :, None] == np.arange(4)).any(1)                                                  # [P,4]
    n = N[:, :, a].astype(np.float64)
    # biallelic patterns: one draw per (pattern, s, a): first present class against the rest
    first = present.argmax(1)
    w1 = np.take_along_axis(W, first[:, None, None], axis=2)[:, :, 0]
    pr = w1 / W.sum(2)
    m = n * np.minimum(pr, 1.0 - pr)
    zero = n == 0
    inv_ = (~zero) & (m < 10)
    bt = (~zero) & (m >= 10)
    print("observed base %d: cells %d  n=0 %.1f %%  inversion %.1f %% (mean search length n*min(p,q) = %.2f)  BTRS %.1f %%" % (
        a, n.size, 100 * zero.mean(), 100 * inv_.mean(), m[inv_].mean() if inv_.any() else 0.0, 100 * bt.mean()))
    # per warp-item (pattern, 32-sample chunk): does it mix regimes?
    mix = 0
    for c in range(S // 32):
        i, b = inv_[:, c * 32:(c + 1) * 32].any(1), bt[:, c * 32:(c + 1) * 32].any(1)
        mix += int((i & b).sum())
    print("   warp items (pattern, 32 samples) that mix inversion and BTRS lanes: %.1f %%" % (100.0 * mix / (P * (S // 32))))
