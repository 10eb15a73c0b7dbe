#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerates the committed golden fixtures.

Runs ONLY in the build container (needs /root/reference).  It imports the UNMODIFIED
reference Python (desman/*.py, bin/desman) and drives the reference's own compiled
sampletau/c_sample_tau.c (oracle/_ref/libref_sampletau.so, built by oracle/Makefile
against oracle/gsl_shim because GSL is absent here).  Three non-invasive shims, none
touching reference source:
  * np.int / np.float aliases (removed in numpy 2),
  * a synthetic 'desman' package object whose __path__ points at /root/reference/desman
    (desman/__init__.py demands an installed distribution),
  * a 'sampletau' module backed by oracle/_ref via ctypes (the shipped Cython output is
    for Cython 0.28 / py2-era CPython); it also records every call for the fixtures.

Outputs (all under tests/golden/):
  sample_tau_kat.npz    kernel-level known answers from the reference C
  loglik_kat.npz        logLikelihood/logPosterior from the reference Python
  mu_stats_ref.npz      Monte-Carlo moments of the reference sampleMu -> (sum_mu, Esum)
  assign_kat.npz        assignTau / logTauProb / logLikelihood(real-valued tau) of the reference Python
  cog0015.npz           the COG0015 count tensor after Variant_Filter (config C1 input)
  cog0015_i3.npz        recorded NMFT + Gibbs chain of `desman ... -g 5 -i 3`
  cog0015_i3/           the CLI's output files of that run
"""
import io
import logging
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle  # noqa: E402  (test infrastructure)

np.int = int      # shim 1
np.float = float

pkg = types.ModuleType("desman")  # shim 2
pkg.__path__ = [os.path.join(REF, "desman")]
sys.modules["desman"] = pkg

CALLS = []  # recorded sample_tau calls


def _tau_idx(tau):
    return np.argmax(tau, axis=2).astype(np.uint8)


class _SampleTauShim(types.ModuleType):  # shim 3
    def __init__(self):
        super().__init__("sampletau")
        self._ref = None
        self.record = True

    def initRNG(self):
        self._ref = oracle.ref()
        self._ref.c_initRNG()

    def setRNG(self, seed):
        self._ref.c_setRNG(seed)

    def freeRNG(self):
        self._ref.c_freeRNG()

    def sample_tau(self, tau, pi, eta, variants):
        assert tau.dtype == np.int64 and tau.flags.c_contiguous and pi.flags.c_contiguous
        assert eta.flags.c_contiguous and variants.dtype == np.int64 and variants.flags.c_contiguous
        before = _tau_idx(tau)
        V, G = tau.shape[0], tau.shape[1]
        n = self._ref.c_sample_tau(tau.ctypes.data_as(oracle._p64), pi.ctypes.data_as(oracle._pd),
                                   eta.ctypes.data_as(oracle._pd), variants.ctypes.data_as(oracle._p64),
                                   V, G, pi.shape[0])
        if self.record:
            CALLS.append(dict(tau_in=before, gamma=np.array(pi), eta=np.array(eta), tau_out=_tau_idx(tau),
                              nchange=n))
        return n


sampletau = _SampleTauShim()
sys.modules["sampletau"] = sampletau

import desman.HaploSNP_Sampler as hsnp  # noqa: E402
import desman.Init_NMFT as inmft  # noqa: E402
import desman.Variant_Filter as vf  # noqa: E402
from numpy.random import RandomState  # noqa: E402

FREQ = os.path.join(REF, "data", "contig_6or16_genesL_scgCOG0015.freq")


# ------------------------------------------------------------------ A. kernel KATs
def synth_counts(rng, V, S, depth, zero_rows=0):
    n = rng.poisson(depth / 4.0, size=(V, S, 4)).astype(np.int64)
    # biallelic-ish structure: suppress two random bases per site most of the time
    for v in range(V):
        if rng.random() < 0.7:
            drop = rng.choice(4, 2, replace=False)
            n[v, :, drop] = rng.binomial(1, 0.02, size=(2, S))
    if zero_rows:
        n[:zero_rows] = 0
    return n


def make_sample_tau_kat():
    rng = np.random.default_rng(20240611)
    cases = []
    spec = [  # (V, G, S, depth, zero_rows, ncalls, seed)
        (40, 1, 1, 20, 2, 2, 1), (40, 2, 3, 5, 3, 2, 7), (64, 5, 64, 100, 2, 2, 23724839),
        (64, 8, 64, 30, 2, 2, 42), (33, 8, 3, 5, 4, 2, 0), (50, 5, 64, 5, 0, 2, 99),
        (16, 3, 7, 8, 1, 40, 12345), (24, 12, 33, 60, 1, 2, 2**31 - 1), (12, 20, 130, 40, 0, 1, 5),
    ]
    out = {}
    for ci, (V, G, S, depth, zr, ncalls, seed) in enumerate(spec):
        counts = synth_counts(rng, V, S, depth, zr)
        tau = np.zeros((V, G, 4), dtype=np.int64)
        idx = rng.integers(0, 4, size=(V, G))
        for v in range(V):
            for g in range(G):
                tau[v, g, idx[v, g]] = 1
        r = oracle.RefSampleTau(seed)
        gammas, etas, taus, nch = [], [], [], []
        for k in range(ncalls):
            conc = 1.0 if k % 2 == 0 else 0.1
            gamma = rng.dirichlet(np.full(G, conc), size=S)
            gamma[gamma < 1e-6] = 1e-6
            gamma = gamma / gamma.sum(axis=1)[:, None]
            if k % 3 == 2:
                eta = rng.dirichlet(np.array([50.0, 1.0, 1.0, 1.0]), size=4)
                eta = np.array([np.roll(eta[a], a) for a in range(4)])
            else:
                eta = 0.96 * np.identity(4) + 0.01 * np.ones((4, 4))
            n = r.sample_tau(tau, np.ascontiguousarray(gamma), np.ascontiguousarray(eta), counts)
            gammas.append(gamma); etas.append(eta); taus.append(_tau_idx(tau)); nch.append(n)
        r.close()
        out[f"c{ci}_meta"] = np.array([V, G, S, ncalls, seed], dtype=np.int64)
        out[f"c{ci}_counts"] = counts.astype(np.int32)
        out[f"c{ci}_tau0"] = idx.astype(np.uint8)
        out[f"c{ci}_gamma"] = np.array(gammas)
        out[f"c{ci}_eta"] = np.array(etas)
        out[f"c{ci}_tau"] = np.array(taus)
        out[f"c{ci}_nchange"] = np.array(nch, dtype=np.int64)
        cases.append(ci)
        print(f"  sample_tau KAT case {ci}: V={V} G={G} S={S} calls={ncalls} flips={nch[:4]}")
    out["ncases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "sample_tau_kat.npz"), **out)


# ------------------------------------------------------------------ B. ll / lp KATs
def _bare_sampler(counts, G):
    """Reference HaploSNP_Sampler without running its (4^G-sized) constructor."""
    h = object.__new__(hsnp.HaploSNP_Sampler)
    h.V, h.S, h.G = counts.shape[0], counts.shape[1], G
    h.variants = np.copy(counts, order="C")
    h.alpha = np.full(G, 0.1)
    h.delta = np.full(4, 0.1)
    h.epsilon = 1e-6
    return h


def make_loglik_kat():
    rng = np.random.default_rng(77)
    out = {}
    spec = [(30, 3, 8, 40), (25, 8, 64, 100), (10, 1, 5, 12), (8, 16, 40, 30)]
    for ci, (V, G, S, depth) in enumerate(spec):
        counts = synth_counts(rng, V, S, depth, 1)
        idx = rng.integers(0, 4, size=(V, G))
        tau = np.zeros((V, G, 4), dtype=np.int64)
        for v in range(V):
            for g in range(G):
                tau[v, g, idx[v, g]] = 1
        gamma = rng.dirichlet(np.full(G, 0.5), size=S)
        gamma[gamma < 1e-6] = 1e-6
        gamma = gamma / gamma.sum(axis=1)[:, None]
        eta = rng.dirichlet(np.array([80.0, 1.0, 1.0, 1.0]), size=4)
        eta = np.array([np.roll(eta[a], a) for a in range(4)])
        h = _bare_sampler(counts, G)
        ll = float(h.logLikelihood(gamma, tau, eta))
        lp = float(h.logPosterior(gamma, tau, eta))
        out[f"c{ci}_counts"] = counts.astype(np.int32)
        out[f"c{ci}_tau"] = idx.astype(np.uint8)
        out[f"c{ci}_gamma"] = gamma
        out[f"c{ci}_eta"] = eta
        out[f"c{ci}_ll_lp"] = np.array([ll, lp])
        print(f"  loglik KAT case {ci}: V={V} G={G} S={S} ll={ll:.6f} lp={lp:.6f}")
    out["ncases"] = np.array(len(spec))
    np.savez_compressed(os.path.join(HERE, "loglik_kat.npz"), **out)


# ------------------------------------------------------------------ C. sampleMu moments
def make_mu_stats_ref():
    rng = np.random.default_rng(5)
    V, G, S, depth, reps = 12, 3, 6, 30, 600
    counts = synth_counts(rng, V, S, depth, 1)
    idx = rng.integers(0, 4, size=(V, G))
    tau = np.zeros((V, G, 4), dtype=np.int64)
    for v in range(V):
        for g in range(G):
            tau[v, g, idx[v, g]] = 1
    gamma = rng.dirichlet(np.full(G, 1.0), size=S)
    eta = 0.90 * np.identity(4) + 0.025 * np.ones((4, 4))  # noisy on purpose: exercises E off-diagonals
    h = _bare_sampler(counts, G)
    h.randomState = RandomState(1234)
    h.E = np.zeros((V, S, 4, 4), dtype=np.int64)
    h.mu = np.zeros((V, S, 4, G), dtype=np.int64)
    sm = np.zeros((reps, S, G))
    es = np.zeros((reps, 4, 4))
    for r in range(reps):
        h.sampleMu(tau, gamma, eta)
        sm[r] = h.mu.sum(axis=(0, 2))
        es[r] = h.E.sum(axis=(0, 1))
    np.savez_compressed(os.path.join(HERE, "mu_stats_ref.npz"), counts=counts.astype(np.int32),
                        tau=idx.astype(np.uint8), gamma=gamma, eta=eta, reps=np.array(reps),
                        sum_mu_mean=sm.mean(0), sum_mu_var=sm.var(0, ddof=1),
                        esum_mean=es.mean(0), esum_var=es.var(0, ddof=1))
    print("  sampleMu reference moments done: sum_mu_mean[0] =", sm.mean(0)[0])


# ------------------------------------------------------------------ D. COG0015 CLI run
def make_cog0015(n_iter=3, G=5, seed=23724839, tag="i3"):
    rec = dict(div=[], nmft_init=None, nmft_final=None, llp=[])

    orig_ri = inmft.Init_NMFT.random_initialize
    orig_fac = inmft.Init_NMFT.factorize
    orig_obj = inmft.Init_NMFT.div_objective
    orig_ll = hsnp.HaploSNP_Sampler.logLikelihood
    orig_lp = hsnp.HaploSNP_Sampler.logPosterior
    orig_rd = hsnp.HaploSNP_Sampler.removeDegenerate

    def ri(self):
        orig_ri(self)
        rec["nmft_init"] = (np.array(self.tau), np.array(self.gamma))
        rec["freq"] = np.array(self.freq_matrix)

    def fac(self):
        orig_fac(self)
        rec["nmft_final"] = (np.array(self.tau), np.array(self.gamma))

    def obj(self):
        d = orig_obj(self)
        rec["div"].append(float(d))
        return d

    depth = [0]

    def ll(self, g, t, e):
        depth[0] += 1
        x = orig_ll(self, g, t, e)
        depth[0] -= 1
        if depth[0] == 0:
            rec["llp"].append(("ll", float(x)))
        return x

    def lp(self, g, t, e):
        depth[0] += 1
        x = orig_lp(self, g, t, e)
        depth[0] -= 1
        if depth[0] == 0:
            rec["llp"].append(("lp", float(x)))
        return x

    def rd(self):
        rec["G_before"] = self.G
        orig_rd(self)
        rec["G_after"] = self.G
        rec["haplo"] = self

    inmft.Init_NMFT.random_initialize = ri
    inmft.Init_NMFT.factorize = fac
    inmft.Init_NMFT.div_objective = obj
    hsnp.HaploSNP_Sampler.logLikelihood = ll
    hsnp.HaploSNP_Sampler.logPosterior = lp
    hsnp.HaploSNP_Sampler.removeDegenerate = rd

    outdir = tempfile.mkdtemp(prefix="desman_golden_")
    del CALLS[:]
    argv = sys.argv
    sys.argv = ["desman", FREQ, "-g", str(G), "-i", str(n_iter), "-s", str(seed), "-o", outdir]
    for hdl in list(logging.root.handlers):
        logging.root.removeHandler(hdl)
    try:
        runpy.run_path(os.path.join(REF, "bin", "desman"), run_name="__main__")
    finally:
        sys.argv = argv
        inmft.Init_NMFT.random_initialize = orig_ri
        inmft.Init_NMFT.factorize = orig_fac
        inmft.Init_NMFT.div_objective = orig_obj
        hsnp.HaploSNP_Sampler.logLikelihood = orig_ll
        hsnp.HaploSNP_Sampler.logPosterior = orig_lp
        hsnp.HaploSNP_Sampler.removeDegenerate = orig_rd
        logging.shutdown()
        for hdl in list(logging.root.handlers):
            logging.root.removeHandler(hdl)

    h = rec["haplo"]
    dst = os.path.join(HERE, f"cog0015_{tag}")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(outdir)):
        if f == "log_file.txt":
            # strip timestamps so the fixture is reproducible
            with open(os.path.join(outdir, f)) as fi, open(os.path.join(dst, f), "w") as fo:
                for line in fi:
                    fo.write(line.split(":INFO:root:", 1)[-1] if ":INFO:root:" in line else line)
        else:
            shutil.copy(os.path.join(outdir, f), os.path.join(dst, f))
    shutil.rmtree(outdir)

    ll_lp = np.array([x for _, x in rec["llp"]])
    kinds = "".join("l" if k == "ll" else "p" for k, _ in rec["llp"])
    np.savez_compressed(
        os.path.join(HERE, f"cog0015_{tag}.npz"),
        meta=np.array([h.V, rec["G_before"], h.S, n_iter, seed, rec["G_after"]], dtype=np.int64),
        nmft_tau0=rec["nmft_init"][0], nmft_gamma0=rec["nmft_init"][1],
        nmft_tau=rec["nmft_final"][0], nmft_gamma=rec["nmft_final"][1],
        nmft_div=np.array(rec["div"]),
        call_tau_in=np.array([c["tau_in"] for c in CALLS[:n_iter]]),
        call_tau_out=np.array([c["tau_out"] for c in CALLS[:n_iter]]),
        call_gamma=np.array([c["gamma"] for c in CALLS[:n_iter]]),
        call_eta=np.array([c["eta"] for c in CALLS[:n_iter]]),
        call_nchange=np.array([c["nchange"] for c in CALLS], dtype=np.int64),
        call2_tau_in=np.array([c["tau_in"] for c in CALLS[n_iter:]]),
        call2_tau_out=np.array([c["tau_out"] for c in CALLS[n_iter:]]),
        call2_gamma=np.array([c["gamma"] for c in CALLS[n_iter:]]),
        call2_eta=np.array([c["eta"] for c in CALLS[n_iter:]]),
        ll_lp=ll_lp, ll_lp_kinds=np.array(kinds),
        tau_star=_tau_idx(h.tau_star), gamma_star=h.gamma_star, eta_star=h.eta_star,
        tau_mean=h.tauMean(), gamma_mean=h.gammaMean(), eta_mean=h.etaMean(),
        lp_star=np.array(h.lp_star), mean_dev=np.array(h.meanDeviance()), ll_store=h.ll_store,
        gamma_store=h.gamma_store, eta_store=h.eta_store)
    print(f"  COG0015 -g {G} -i {n_iter}: G_after={rec['G_after']} lp_star={h.lp_star:.6f} "
          f"nchange={[c['nchange'] for c in CALLS]} nmft_iters={len(rec['div']) - 1}")


def make_cog0015_input():
    import pandas as p
    variants = p.read_csv(FREQ, header=0, index_col=0)
    f = vf.Variant_Filter(variants, randomState=RandomState(238329), optimise=True, threshold=None,
                          min_coverage=5.0, qvalue_cutoff=1.0e-3)
    snps = np.asarray(f.snps_filter)
    assert snps.max() < 32768
    cols = variants.columns.values.tolist()
    np.savez_compressed(os.path.join(HERE, "cog0015.npz"), snps=snps.astype(np.int16),
                        eta=np.asarray(f.eta), position=np.asarray(variants["Position"]),
                        contigs=np.array([str(x) for x in variants.index]), columns=np.array([str(c) for c in cols]))
    print("  COG0015 input:", snps.shape, "max", snps.max(), "zeros %.3f" % (snps == 0).mean())


# ------------------------------------------------------------------ F. joint-state enumeration (assignTau, logTauProb), DIC
def make_assign_kat():
    """assignTau (:233-261), logTauProb (:498-524) and logLikelihood of a real-valued tau (DIC, :486-496) of the UNMODIFIED
    reference class (full constructor: tauStates[4^G,G,4] is small for G <= 5)."""
    rng = np.random.default_rng(4242)
    out = {}
    spec = [(20, 3, 6, 30), (10, 5, 16, 40), (8, 1, 4, 12), (12, 2, 3, 25)]
    for ci, (V, G, S, depth) in enumerate(spec):
        counts = synth_counts(rng, V, S, depth, 1)
        newc = synth_counts(rng, V + 3, S, depth, 1)
        h = hsnp.HaploSNP_Sampler(np.copy(counts), G, RandomState(11 + ci), max_iter=3)
        gamma = rng.dirichlet(np.full(G, 0.7), size=S)
        gamma[gamma < 1e-6] = 1e-6
        gamma = gamma / gamma.sum(axis=1)[:, None]
        eta = rng.dirichlet(np.array([60.0, 1.0, 1.0, 1.0]), size=4)
        eta = np.array([np.roll(eta[a], a) for a in range(4)])
        h.gamma_star, h.eta_star = np.copy(gamma), np.copy(eta)
        # the most probable joint state of every site, the runner-up at every third one (stays clear of exp underflow: the
        # reference takes math.log of the normalised probability)
        sp = np.array([h.baseProbabilityGivenTau(h.tauStates[t], gamma, eta) for t in range(h.nTauStates)])
        lp = np.einsum("tsb,vsb->vt", np.log(sp), counts.astype(float))
        order = np.argsort(-lp, axis=1, kind="stable")
        star = order[:, 0].copy()
        if h.nTauStates > 1:
            star[::3] = order[::3, 1]
        h.tauIndices_star = star.astype(np.int64)
        ltp = float(h.logTauProb(gamma, eta))
        h.randomState = RandomState(9000 + ci)
        aT, conf = h.assignTau(np.reshape(newc, (V + 3, S * 4)))
        after = h.randomState.random_sample(2)
        tau_real = rng.dirichlet(np.full(4, 0.3), size=(V, G))
        ll_real = float(h.logLikelihood(gamma, tau_real, eta))
        out.update({f"c{ci}_counts": counts.astype(np.int32), f"c{ci}_new": newc.astype(np.int32), f"c{ci}_gamma": gamma,
                    f"c{ci}_eta": eta, f"c{ci}_star": star.astype(np.int64), f"c{ci}_logtauprob": np.array(ltp),
                    f"c{ci}_assign": np.argmax(aT, axis=2).astype(np.uint8), f"c{ci}_conf": conf, f"c{ci}_rng_after": after,
                    f"c{ci}_tau_real": tau_real, f"c{ci}_ll_real": np.array(ll_real), f"c{ci}_rng_seed": np.array(9000 + ci)})
        print(f"  assign KAT case {ci}: V={V} G={G} S={S} logTauProb={ltp:.6f} ll(real tau)={ll_real:.6f} conf[:3]={conf[:3]}")
    out["ncases"] = np.array(len(spec))
    np.savez_compressed(os.path.join(HERE, "assign_kat.npz"), **out)


def make_fix_tau_kat():
    """storeHLogProb of the UNMODIFIED sampleTauFixTau (HaploSNP_Sampler.py:196-222): the normalised log-probabilities of the
    step at strain H under the tau the call came in with -- a deterministic function of the inputs (the draws that follow come
    from numpy's stream and are not pinned).  Pins oracle_sample_tau_fix_philox's logp."""
    out = {}
    for ci, (V, S, G, H, seed) in enumerate([(30, 12, 4, 0, 5), (24, 40, 6, 3, 9), (16, 9, 3, 2, 2), (12, 20, 5, 4, 7)]):
        rng = np.random.default_rng(seed)
        counts = synth_counts(rng, V, S, 40.0, zero_rows=1)
        tau = np.zeros((V, G, 4), dtype=np.int64)
        idx = rng.integers(0, 4, size=(V, G))
        np.put_along_axis(tau, idx[:, :, None], 1, axis=2)
        gamma = rng.dirichlet(np.ones(G), size=S)
        eta = 0.95 * np.identity(4) + 0.0125 + 0.002 * rng.random((4, 4))
        eta /= eta.sum(1)[:, None]
        h = _bare_sampler(counts, G)
        h.randomState = RandomState(seed)
        work = tau.copy()
        logp = h.sampleTauFixTau(work, H, gamma, eta)
        out[f"c{ci}_meta"] = np.array([V, S, G, H, seed], dtype=np.int64)
        out[f"c{ci}_counts"] = counts.astype(np.int32)
        out[f"c{ci}_tau"] = idx.astype(np.uint8); out[f"c{ci}_gamma"] = gamma; out[f"c{ci}_eta"] = eta
        out[f"c{ci}_logp"] = logp
        assert np.array_equal(work[:, :H], tau[:, :H])                 # strains below H are pinned
    np.savez_compressed(os.path.join(HERE, "fix_tau_kat.npz"), **out)
    print("fix_tau_kat.npz written")


def make_host_helpers_kat():
    """Small host-side methods of the UNMODIFIED class that the mirror re-implements in numpy: normaliseLogProb (:186-194),
    logMean (:526-540), tauDist (:137-147), baseProbabilityGivenTau (:129-134), updateTauIndices / mapTauState (:224-231)."""
    rng = np.random.default_rng(77)
    V, S, G = 9, 7, 6
    counts = synth_counts(rng, V, S, 20.0)
    h = _bare_sampler(counts, G)
    h.tauMap = np.zeros((G, 4), dtype=np.int64)
    for g in range(G):
        for a in range(4):
            h.tauMap[g, a] = a * (4 ** (G - g - 1))
    idx = rng.integers(0, 4, size=(V, G))
    tau = np.zeros((V, G, 4), dtype=np.int64)
    np.put_along_axis(tau, idx[:, :, None], 1, axis=2)
    h.tau = tau
    h.tauIndices = np.zeros(V, dtype=np.int64)
    h.updateTauIndices()
    lv = [rng.normal(-50.0, 30.0, size=4) for _ in range(5)] + [np.array([-1e4, -1e4 - 1.0, -2e4, -1e4 - 0.5])]
    ls = [rng.normal(-1e3, 5.0, size=n) for n in (3, 17, 50)]
    gamma = rng.dirichlet(np.ones(G), size=S)
    eta = 0.9 * np.identity(4) + 0.025
    np.savez_compressed(os.path.join(HERE, "host_helpers_kat.npz"), tau=idx.astype(np.uint8), tauIndices=h.tauIndices,
                        lv=np.array(lv), lv_out=np.array([h.normaliseLogProb(x) for x in lv]),
                        ls0=ls[0], ls1=ls[1], ls2=ls[2], ls_out=np.array([h.logMean(x) for x in ls]),
                        dist=np.array([h.tauDist(tau[i], tau[i + 1]) for i in range(V - 1)], dtype=np.int64),
                        gamma=gamma, eta=eta, base_prob=np.array([h.baseProbabilityGivenTau(tau[i], gamma, eta) for i in range(V)]))
    print("host_helpers_kat.npz written")


def make_degenerate_kat():
    """calculateSND (:712-730), compSND (:747-770), variableTau (:732-745) and removeDegenerate (:771-832) of the UNMODIFIED
    class (real constructor, G = 5) on a tau with duplicated haplotypes."""
    rng = np.random.default_rng(31)
    V, S, G = 14, 6, 5
    counts = synth_counts(rng, V, S, 20.0)
    h = hsnp.HaploSNP_Sampler(counts, G, RandomState(31), max_iter=2)
    idx = rng.integers(0, 4, size=(V, G))
    idx[:, 3] = idx[:, 0]                          # strain 3 == strain 0, strain 4 == strain 1: two merges
    idx[:, 4] = idx[:, 1]
    idx[5, :] = 2                                  # one site without variation
    tau = np.zeros((V, G, 4), dtype=np.int64)
    np.put_along_axis(tau, idx[:, :, None], 1, axis=2)
    gamma = rng.dirichlet(np.ones(G), size=S)
    h.tau = tau.copy(); h.gamma = gamma.copy()
    other = np.zeros((V, 3, 4), dtype=np.int64)
    np.put_along_axis(other, rng.integers(0, 4, size=(V, 3))[:, :, None], 1, axis=2)
    out = dict(tau=idx.astype(np.uint8), gamma=gamma, other=np.argmax(other, 2).astype(np.uint8),
               snd=h.calculateSND(h.tau), comp=h.compSND(h.tau, other), variable=h.variableTau(h.tau))
    h.removeDegenerate()
    out.update(G_after=np.array(h.G), tau_after=np.argmax(h.tau, 2).astype(np.uint8), gamma_after=h.gamma,
               tauIndices_after=np.asarray(h.tauIndices), alpha_after=h.alpha, gamma_store_shape=np.array(h.gamma_store.shape))
    np.savez_compressed(os.path.join(HERE, "degenerate_kat.npz"), **out)
    print("degenerate_kat.npz written, G %d -> %d" % (G, h.G))


def make_nmft_steps_kat():
    """Single steps of the UNMODIFIED Init_NMFT class (div_objective, div_update + _adjustment, div_update_tau,
    div_update_gamma, factorize_gamma) on small problems: pins desman_b200.Init_NMFT's step methods."""
    out = {}
    for ci, (V, S, G, seed) in enumerate([(40, 12, 3, 7), (25, 40, 5, 11), (30, 9, 1, 3)]):
        rng = np.random.default_rng(seed)
        counts = synth_counts(rng, V, S, 30.0)
        n = inmft.Init_NMFT(counts, G, RandomState(seed), max_iter=25)
        n.random_initialize()
        n._adjustment()
        out[f"c{ci}_meta"] = np.array([V, S, G, seed], dtype=np.int64)
        out[f"c{ci}_counts"] = counts.astype(np.int32)
        out[f"c{ci}_tau0"] = n.tau.copy(); out[f"c{ci}_gamma0"] = n.gamma.copy()
        out[f"c{ci}_div0"] = np.array(n.div_objective())
        n.div_update(); n._adjustment()
        out[f"c{ci}_tau1"] = n.tau.copy(); out[f"c{ci}_gamma1"] = n.gamma.copy(); out[f"c{ci}_div1"] = np.array(n.div_objective())
        n.div_update_tau()
        out[f"c{ci}_tau2"] = n.tau.copy(); out[f"c{ci}_div2"] = np.array(n.div_objective())
        n.div_update_gamma()
        out[f"c{ci}_gamma3"] = n.gamma.copy(); out[f"c{ci}_div3"] = np.array(n.div_objective())
        buf = io.StringIO()
        so, sys.stdout = sys.stdout, buf
        try:
            n.factorize_gamma()
        finally:
            sys.stdout = so
        out[f"c{ci}_gamma4"] = n.gamma.copy(); out[f"c{ci}_div4"] = np.array(n.div_objective())
    np.savez_compressed(os.path.join(HERE, "nmft_steps_kat.npz"), **out)
    print("nmft_steps_kat.npz written")


if __name__ == "__main__":
    what = sys.argv[1:] or ["kat", "loglik", "mu", "input", "i3"]
    oracle.build()
    if "kat" in what:
        make_sample_tau_kat()
    if "loglik" in what:
        make_loglik_kat()
    if "mu" in what:
        make_mu_stats_ref()
    if "assign" in what:
        make_assign_kat()
    if "nmft" in what:
        make_nmft_steps_kat()
    if "fixtau" in what:
        make_fix_tau_kat()
    if "helpers" in what:
        make_host_helpers_kat()
    if "degenerate" in what:
        make_degenerate_kat()
    if "input" in what:
        make_cog0015_input()
    if "i3" in what:
        make_cog0015(3, 5, 23724839, "i3")
    if "i50" in what:
        make_cog0015(50, 5, 23724839, "i50")
