#!/usr/bin/env python3
"""Developer probe: NMFT iterations/s at BASELINE config C3 (two calls with different iteration counts: the difference is
the cost of the iterations without the upload / X construction / download)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from desman_b200 import engine
from desman_b200.synth import synth_counts
V, S, G = 100000, 64, 8
p = synth_counts(V, S, G)
rng = np.random.default_rng(1)
tau0 = rng.dirichlet(np.full(4, 0.01), size=V * G).reshape(V, G, 4).transpose(2, 0, 1).reshape(4 * V, G).copy()
gamma0 = rng.dirichlet(np.full(G, 0.01), size=S).T.copy()
e = engine.Engine(0, seed=1)
ts = {}
for n in (20, 220, 20, 220):
    t0 = time.perf_counter()
    tau, gamma, it, div, _ = e.nmft_factorize(p["counts"], tau0, gamma0, max_iter=n, min_change=0.0)
    ts.setdefault(n, []).append(time.perf_counter() - t0)
    print("max_iter %d -> %d iterations, div %.6g, %.1f ms" % (n, it, div, 1e3 * ts[n][-1]))
per = (min(ts[220]) - min(ts[20])) / 200
bytes_iter = 2 * 8 * 4 * V * S
print("NMFT: %.1f us per iteration = %.0f iterations/s; X passes %.0f MB per iteration -> %.0f GB/s" % (1e6 * per, 1 / per, bytes_iter / 1e6, bytes_iter / per / 1e9))
