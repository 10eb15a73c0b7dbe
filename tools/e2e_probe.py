#!/usr/bin/env python3
"""Where does the end-to-end overhead of HaploSNP_Sampler.update() go?  (developer probe)"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
from numpy.random import RandomState

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler  # noqa: E402
from desman_b200.synth import CHAIN_SEED, onehot, synth_counts  # noqa: E402

p = synth_counts(100000, 64, 8)
for rep in range(2):
    hs = HaploSNP_Sampler(p["counts"], 8, RandomState(1), max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 20, seed=CHAIN_SEED)
    hs.tau, hs.gamma, hs.eta = onehot(p["tau0"]), p["gamma0"].copy(), p["eta0"].copy()
    t0 = time.perf_counter()
    if rep == 1:
        pr = cProfile.Profile(); pr.enable()
    hs.update()
    if rep == 1:
        pr.disable()
    print("update() wall %.1f ms (device %.1f ms)" % (1e3 * (time.perf_counter() - t0), hs._timing["elapsed_ms"]))
    hs.close()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
