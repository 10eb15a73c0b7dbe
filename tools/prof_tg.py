import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from desman_b200 import engine
from desman_b200.synth import synth_counts, CHAIN_SEED
p = synth_counts(100000, 64, 8)
e = engine.Engine(0, seed=CHAIN_SEED)
e.set_counts(p["counts"]); e.set_state(None, p["gamma0"], p["eta0"], G=8); e.set_tau_index(p["tau0"])
e.update(12)
print("stats", e.get_group_stats(), flush=True)
import os
os.environ["X"]="1"
e.set_profiling(True, True)
e.update(1)
e.synchronize()
print(e.get_timing()["kernel_ms"])
