#!/usr/bin/env python3
"""Developer probe: cost of creating/destroying an engine context and of the first calls."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from desman_b200 import engine
from desman_b200.synth import synth_counts, onehot
p = synth_counts(100000, 64, 8)
tau = onehot(p["tau0"])
for rep in range(4):
    t0 = time.perf_counter(); e = engine.Engine(0, seed=1); t1 = time.perf_counter()
    e.set_counts(p["counts"]); t2 = time.perf_counter()
    e.set_state(tau, p["gamma0"], p["eta0"]); t3 = time.perf_counter()
    out = e.update(20); t4 = time.perf_counter()
    ts = e.get_tau_sum(); t5 = time.perf_counter()
    st = e.get_state(); t6 = time.perf_counter()
    e.close(); t7 = time.perf_counter()
    print("rep %d: create %.1f set_counts %.1f set_state %.1f update(20) %.1f tau_sum %.1f get_state %.1f close %.1f ms" % (
        rep, *(1e3 * x for x in (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5, t7 - t6))))
