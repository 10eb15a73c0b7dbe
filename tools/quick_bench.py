#!/usr/bin/env python3
"""Developer timing probe (not the contract bench): per-kernel device times for one config."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from desman_b200 import engine  # noqa: E402
from desman_b200.synth import synth_counts  # noqa: E402


def main():
    V, S, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    n_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    t0 = time.time()
    p = synth_counts(V, S, G)
    t1 = time.time()
    e = engine.Engine(0, seed=23724839)
    e.set_counts(p["counts"])
    e.set_state(None, p["gamma0"], p["eta0"], G=G)
    e.set_tau_index(p["tau0"])
    t2 = time.time()
    e.update(3)
    e.get_tier_counts()
    e.set_profiling(True, False)
    out = e.update(n_iter)
    tiers = e.get_tier_counts().tolist()
    tm = e.get_timing()
    res = dict(V=V, S=S, G=G, n_iter=n_iter, gen_s=t1 - t0, upload_s=t2 - t1, ms_per_sweep=tm["elapsed_ms"] / n_iter,
               kernel_ms_per_sweep={k: v / n_iter for k, v in tm["kernel_ms"].items()},
               launches=tm["kernel_launches"], tiers=tiers, nchange=out["nchange"].tolist()[:10], lp=out["lp_store"][[0, -1]].tolist())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
