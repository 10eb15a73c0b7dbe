"""Developer probe (CPU, uses the oracle: test infrastructure): would "lazy screening" of the tau update pay off?

Keeps per site the margin of its gap test against a reference table and asks, sweep after sweep, whether the cheap rigorous
bound N_v * max|Wd - Wd_ref| still covers it.  Result at V=20000 of config C3: it does not -- the bound
is ~340 nats after ONE sweep against median margins of ~120 nats, while the true drift of the sums is 1-6 nats over 12 sweeps.
"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from desman_b200.synth import synth_counts, onehot
from oracle import oracle
V, S, G = 20000, 64, 8
p = synth_counts(V, S, G)
seed = 23724839
t0 = time.time()
res = oracle.update(onehot(p["tau0"]), p["gamma0"], p["eta0"], p["counts"], 25, seed, mu_mode=0)
print("burn-in 25 sweeps", time.time() - t0, "s; nchange tail", res["nchange"][-5:])
tau = np.argmax(res["tau"], 2) if res["tau"].ndim == 3 else res["tau"]
gam = [res["gamma"]]; etas = [res["eta"]]
state = dict(tau=res["tau"], gamma=res["gamma"], eta=res["eta"])
# continue the chain sweep by sweep, recording gamma/eta
K = 12
cur = state
for k in range(K):
    r = oracle.update(cur["tau"], cur["gamma"], cur["eta"], p["counts"], 1, seed, sweep0=25 + k, mu_mode=0)
    cur = dict(tau=r["tau"], gamma=r["gamma"], eta=r["eta"])
    gam.append(r["gamma"]); etas.append(r["eta"])
print("flips per sweep after burn-in ~", r["nchange"])
counts = p["counts"].astype(np.float64)
N = counts.sum((1, 2))
tau_idx = tau.astype(np.int64)

def wd_tables(gamma, eta):
    """per site: Wd[v,g,j,s,b] = log2(P - eta[cur]g + eta[a_j]g) - log2 P  (a_j = (cur+1+j)&3)"""
    out = np.empty((V, G, 3, S, 4), dtype=np.float32)
    for lo in range(0, V, 2000):
        t = tau_idx[lo:lo + 2000]
        P = np.einsum("sg,vgb->vsb", gamma, eta[t])                      # [v,S,4]
        for g in range(G):
            c = t[:, g]
            base = P - eta[c][:, None, :] * gamma[None, :, g, None]
            for j in range(3):
                a = (c + 1 + j) & 3
                q = base + eta[a][:, None, :] * gamma[None, :, g, None]
                out[lo:lo + 2000, g, j] = np.log2(q) - np.log2(P)
    return out

W0 = wd_tables(gam[0], etas[0])
D0 = np.einsum("vgjsb,vsb->vgj", W0, counts.astype(np.float32)) * np.log(2.0)     # nats, L_a - L_cur
margin = (-60.0 - D0.max((1, 2)))                                                 # > 0: all strains decided "stay"
print("sites decided at the reference sweep: %.2f %%" % (100.0 * (margin > 0).mean()))
print("margin quantiles (nats) of decided sites:", np.percentile(margin[margin > 0], [1, 5, 25, 50]))
for k in range(1, K + 1):
    Wk = wd_tables(gam[k], etas[k])
    delta = np.abs(Wk - W0).max((1, 2, 3, 4))                                     # per site = per pattern
    need = N * delta * np.log(2.0)
    safe = margin > need
    Dk = np.einsum("vgjsb,vsb->vgj", Wk, counts.astype(np.float32)) * np.log(2.0)
    truly = (-60.0 - Dk.max((1, 2))) > 0
    print("sweep +%2d: max-entry bound: safe %.2f %% (truly decided %.2f %%), median N*delta %.1f nats, median |D-D0| %.1f nats" % (
        k, 100.0 * safe.mean(), 100.0 * truly.mean(), np.median(need), np.median(np.abs(Dk - D0).max((1, 2)))))
