"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE.

ctypes front-end to oracle/liboracle.so (our C restatement of the DESMAN hot path,
desman_oracle.c) and to oracle/_ref/libref_sampletau.so (the reference's own
sampletau/c_sample_tau.c compiled unmodified against oracle/gsl_shim).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package desman_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None

STAGE_TAU, STAGE_MU, STAGE_GAMMA, STAGE_ETA, STAGE_GAMMA_BOOST, STAGE_ETA_BOOST = 1, 2, 3, 4, 5, 6


def build(force=False):
    """Compile liboracle.so (always possible) and _ref (only where /root/reference exists)."""
    need = force or not os.path.exists(os.path.join(_HERE, "liboracle.so"))
    if not need:
        src = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("desman_oracle.c", "desman_oracle.h"))
        need = src > os.path.getmtime(os.path.join(_HERE, "liboracle.so"))
    if need:
        subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if os.path.exists("/root/reference/sampletau/c_sample_tau.c") and (
            force or not os.path.exists(os.path.join(_HERE, "_ref", "libref_sampletau.so"))):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


_p64 = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)
_pu32 = C.POINTER(C.c_uint32)


class MT19937(C.Structure):
    _fields_ = [("mt", C.c_uint32 * 624), ("mti", C.c_int)]


class ChainCfg(C.Structure):
    _fields_ = [("V", C.c_int), ("G", C.c_int), ("S", C.c_int), ("n_iter", C.c_int),
                ("alpha", C.c_double), ("delta", C.c_double), ("epsilon", C.c_double),
                ("seed", C.c_uint64), ("sweep0", C.c_uint32), ("mu_mode", C.c_int)]


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        L.oracle_mt_seed.argtypes = [C.POINTER(MT19937), C.c_ulong]
        L.oracle_mt_next.argtypes = [C.POINTER(MT19937)]
        L.oracle_mt_next.restype = C.c_uint32
        L.oracle_mt_fill.argtypes = [C.POINTER(MT19937), _pu32, C.c_int64]
        L.oracle_philox4x32_10.argtypes = [_pu32, _pu32, _pu32]
        L.oracle_sample_tau_words.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int, _pu32]
        L.oracle_sample_tau_words.restype = C.c_int
        L.oracle_sample_tau_mt.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int, C.POINTER(MT19937)]
        L.oracle_sample_tau_mt.restype = C.c_int
        L.oracle_sample_tau_philox.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int,
                                               C.c_uint64, C.c_uint32, C.c_int64]
        L.oracle_sample_tau_philox.restype = C.c_int
        L.oracle_sample_tau_fix_philox.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int,
                                                   C.c_uint64, C.c_uint32, C.c_int64, C.c_int, _pd]
        L.oracle_sample_tau_fix_philox.restype = C.c_int
        L.oracle_tau_step_probs.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int, _pd, _pd]
        L.oracle_mu_stats.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int,
                                      C.c_uint64, C.c_uint32, C.c_int64, _p64, _p64]
        L.oracle_mu_stats_agg.argtypes = L.oracle_mu_stats.argtypes
        L.oracle_draw_gamma.argtypes = [_p64, C.c_int, C.c_int, C.c_double, C.c_double,
                                        C.c_uint64, C.c_uint32, _pd]
        L.oracle_draw_eta.argtypes = [_p64, C.c_double, C.c_uint64, C.c_uint32, _pd]
        L.oracle_gamma_variate.argtypes = [C.c_double, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.oracle_gamma_variate.restype = C.c_double
        L.oracle_loglik.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int]
        L.oracle_loglik.restype = C.c_double
        L.oracle_logprior.argtypes = [_pd, _pd, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
        L.oracle_logprior.restype = C.c_double
        L.oracle_update.argtypes = [C.POINTER(ChainCfg), _p64, _pd, _pd, _p64,
                                    _pd, _pd, _pd, _pd, _p64, _p64, _p64, _pd, _pd, _pd,
                                    C.POINTER(C.c_int), _p64, _p64]
        L.oracle_update_tau.argtypes = [C.POINTER(ChainCfg), C.c_int, C.POINTER(MT19937), _p64, _pd, _pd,
                                        _p64, _pd, _pd, _p64, _p64, _p64, _pd]
        L.oracle_nmft_freq.argtypes = [_p64, C.c_int, C.c_int, _pd]
        L.oracle_nmft_objective.argtypes = [_pd, _pd, _pd, C.c_int, C.c_int, C.c_int]
        L.oracle_nmft_objective.restype = C.c_double
        L.oracle_nmft_update.argtypes = [_pd, _pd, _pd, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_nmft_factorize.argtypes = [_pd, _pd, _pd, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.c_int, _pd, _pd]
        L.oracle_nmft_factorize.restype = C.c_int
        L.oracle_nmft_get_tau.argtypes = [_pd, C.c_int, C.c_int, _p64]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_sampletau.so"))


def ref():
    """The reference's own c_sample_tau.c (c_initRNG/c_setRNG/c_freeRNG/c_sample_tau)."""
    global _REF
    if _REF is None:
        build()
        R = C.CDLL(os.path.join(_HERE, "_ref", "libref_sampletau.so"))
        R.c_setRNG.argtypes = [C.c_ulong]
        R.c_sample_tau.argtypes = [_p64, _pd, _pd, _p64, C.c_int, C.c_int, C.c_int]
        R.c_sample_tau.restype = C.c_int
        _REF = R
    return _REF


def _i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_p64)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_pd)


def _opt(a, ptype):
    return a.ctypes.data_as(ptype) if a is not None else None


# ---------------------------------------------------------------- RNG
def mt_words(seed, n, skip=0):
    st = MT19937()
    lib().oracle_mt_seed(C.byref(st), seed)
    out = np.empty(skip + n, dtype=np.uint32)
    lib().oracle_mt_fill(C.byref(st), out.ctypes.data_as(_pu32), skip + n)
    return out[skip:]


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*[int(x) & 0xffffffff for x in ctr])
    k = (C.c_uint32 * 2)(*[int(x) & 0xffffffff for x in key])
    o = (C.c_uint32 * 4)()
    lib().oracle_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def tau_words_philox(V, G, seed, sweep, v0=0):
    out = np.empty((V, G), dtype=np.uint32)
    key = (seed & 0xffffffff, (seed >> 32) & 0xffffffff)
    for v in range(V):
        for g in range(G):
            out[v, g] = philox((v0 + v, g, sweep, STAGE_TAU << 28), key)[0]
    return out


# ---------------------------------------------------------------- tau
def sample_tau_words(tau, pi, eta, variants, words):
    """In place on a C-contiguous int64 tau; returns nchange."""
    assert tau.dtype == np.int64 and tau.flags.c_contiguous
    V, G = tau.shape[0], tau.shape[1]
    S = pi.shape[0]
    pi, ppi = _f64(pi)
    eta, pe = _f64(eta)
    variants, pv = _i64(variants)
    words = np.ascontiguousarray(words, dtype=np.uint32)
    return lib().oracle_sample_tau_words(tau.ctypes.data_as(_p64), ppi, pe, pv, V, G, S,
                                         words.ctypes.data_as(_pu32))


def sample_tau_philox(tau, pi, eta, variants, seed, sweep, v0=0):
    assert tau.dtype == np.int64 and tau.flags.c_contiguous
    V, G = tau.shape[0], tau.shape[1]
    S = pi.shape[0]
    pi, ppi = _f64(pi)
    eta, pe = _f64(eta)
    variants, pv = _i64(variants)
    return lib().oracle_sample_tau_philox(tau.ctypes.data_as(_p64), ppi, pe, pv, V, G, S, seed, sweep, v0)


def sample_tau_fix_philox(tau, H, pi, eta, variants, seed, sweep, v0=0):
    """sampleTauFixTau (HaploSNP_Sampler.py:196-222) under the Philox contract; tau int64 one-hot, in place.
    Returns (storeHLogProb [V,4], nchange)."""
    assert tau.dtype == np.int64 and tau.flags.c_contiguous
    V, G = tau.shape[0], tau.shape[1]
    S = pi.shape[0]
    pi, ppi = _f64(pi)
    eta, pe = _f64(eta)
    variants, pv = _i64(variants)
    logp = np.zeros((V, 4))
    n = lib().oracle_sample_tau_fix_philox(tau.ctypes.data_as(_p64), ppi, pe, pv, V, G, S, seed, sweep, v0, H,
                                           logp.ctypes.data_as(_pd))
    return logp, n


class RefSampleTau:
    """Drives the reference's own compiled c_sample_tau (process-global GSL-compatible RNG)."""

    def __init__(self, seed):
        self.R = ref()
        self.R.c_initRNG()
        self.R.c_setRNG(seed)

    def sample_tau(self, tau, pi, eta, variants):
        assert tau.dtype == np.int64 and tau.flags.c_contiguous
        V, G = tau.shape[0], tau.shape[1]
        S = pi.shape[0]
        pi, ppi = _f64(pi)
        eta, pe = _f64(eta)
        variants, pv = _i64(variants)
        return self.R.c_sample_tau(tau.ctypes.data_as(_p64), ppi, pe, pv, V, G, S)

    def close(self):
        self.R.c_freeRNG()


def tau_step_probs(tau_index_v, pi, eta, variants_v, g):
    G = len(tau_index_v)
    S = pi.shape[0]
    ti, pti = _i64(tau_index_v)
    pi, ppi = _f64(pi)
    eta, pe = _f64(eta)
    nv, pnv = _i64(variants_v)
    lp = np.zeros(4)
    pr = np.zeros(4)
    lib().oracle_tau_step_probs(pti, ppi, pe, pnv, G, S, g, lp.ctypes.data_as(_pd), pr.ctypes.data_as(_pd))
    return lp, pr


# ---------------------------------------------------------------- sweep pieces
def mu_stats(tau, gamma, eta, variants, seed, sweep, v0=0, mode=0):
    V, G = tau.shape[0], tau.shape[1]
    S = gamma.shape[0]
    tau, pt = _i64(tau)
    gamma, pg = _f64(gamma)
    eta, pe = _f64(eta)
    variants, pv = _i64(variants)
    sum_mu = np.zeros((S, G), dtype=np.int64)
    esum = np.zeros((4, 4), dtype=np.int64)
    fn = lib().oracle_mu_stats_agg if mode == 1 else lib().oracle_mu_stats
    fn(pt, pg, pe, pv, V, G, S, seed, sweep, v0, sum_mu.ctypes.data_as(_p64), esum.ctypes.data_as(_p64))
    return sum_mu, esum


def draw_gamma(sum_mu, alpha, epsilon, seed, sweep):
    S, G = sum_mu.shape
    sum_mu, pm = _i64(sum_mu)
    out = np.zeros((S, G))
    lib().oracle_draw_gamma(pm, S, G, alpha, epsilon, seed, sweep, out.ctypes.data_as(_pd))
    return out


def draw_eta(esum, delta, seed, sweep):
    esum, pe = _i64(esum)
    out = np.zeros((4, 4))
    lib().oracle_draw_eta(pe, delta, seed, sweep, out.ctypes.data_as(_pd))
    return out


def gamma_variate(shape, seed, sweep, idx, stage=STAGE_GAMMA, boost_stage=STAGE_GAMMA_BOOST):
    return lib().oracle_gamma_variate(shape, seed, sweep, idx, stage, boost_stage)


def loglik(tau, gamma, eta, variants):
    V, G = tau.shape[0], tau.shape[1]
    S = gamma.shape[0]
    tau, pt = _i64(tau)
    gamma, pg = _f64(gamma)
    eta, pe = _f64(eta)
    variants, pv = _i64(variants)
    return lib().oracle_loglik(pt, pg, pe, pv, V, G, S)


def logprior(gamma, eta, V, alpha=0.1, delta=0.1):
    S, G = gamma.shape
    gamma, pg = _f64(gamma)
    eta, pe = _f64(eta)
    return lib().oracle_logprior(pg, pe, V, G, S, alpha, delta)


def logpost(tau, gamma, eta, variants, alpha=0.1, delta=0.1):
    return loglik(tau, gamma, eta, variants) + logprior(gamma, eta, tau.shape[0], alpha, delta)


def update(tau, gamma, eta, variants, n_iter, seed, sweep0=0, alpha=0.1, delta=0.1, epsilon=1e-6, mu_mode=0):
    """Runs the restated update() chain; returns a dict of outputs (inputs are copied)."""
    tau = np.array(tau, dtype=np.int64, order="C")
    gamma = np.array(gamma, dtype=np.float64, order="C")
    eta = np.array(eta, dtype=np.float64, order="C")
    variants, pv = _i64(variants)
    V, G = tau.shape[0], tau.shape[1]
    S = gamma.shape[0]
    cfg = ChainCfg(V, G, S, n_iter, alpha, delta, epsilon, seed, sweep0, mu_mode)
    out = dict(
        gamma_store=np.zeros((n_iter, S, G)), eta_store=np.zeros((n_iter, 4, 4)),
        ll_store=np.zeros(n_iter), lp_store=np.zeros(n_iter), nchange=np.zeros(n_iter, dtype=np.int64),
        tau_sum=np.zeros((V, G, 4), dtype=np.int64), tau_star=np.zeros((V, G, 4), dtype=np.int64),
        gamma_star=np.zeros((S, G)), eta_star=np.zeros((4, 4)),
        sum_mu=np.zeros((S, G), dtype=np.int64), esum=np.zeros((4, 4), dtype=np.int64))
    lp_star = C.c_double(0.0)
    iter_star = C.c_int(0)
    lib().oracle_update(C.byref(cfg), tau.ctypes.data_as(_p64), gamma.ctypes.data_as(_pd),
                        eta.ctypes.data_as(_pd), pv,
                        out["gamma_store"].ctypes.data_as(_pd), out["eta_store"].ctypes.data_as(_pd),
                        out["ll_store"].ctypes.data_as(_pd), out["lp_store"].ctypes.data_as(_pd),
                        out["nchange"].ctypes.data_as(_p64), out["tau_sum"].ctypes.data_as(_p64),
                        out["tau_star"].ctypes.data_as(_p64), out["gamma_star"].ctypes.data_as(_pd),
                        out["eta_star"].ctypes.data_as(_pd), C.byref(lp_star), C.byref(iter_star),
                        out["sum_mu"].ctypes.data_as(_p64), out["esum"].ctypes.data_as(_p64))
    out.update(tau=tau, gamma=gamma, eta=eta, lp_star=lp_star.value, iter_star=iter_star.value)
    return out


def update_tau(tau, gamma_store, eta_store, variants, seed, sweep0=0, use_mt=False, mt_state=None,
               alpha=0.1, delta=0.1):
    tau = np.array(tau, dtype=np.int64, order="C")
    gamma_store, pgs = _f64(gamma_store)
    eta_store, pes = _f64(eta_store)
    variants, pv = _i64(variants)
    n_iter, S, G = gamma_store.shape
    V = tau.shape[0]
    cfg = ChainCfg(V, G, S, n_iter, alpha, delta, 1e-6, seed, sweep0, 0)
    out = dict(ll_store=np.zeros(n_iter), lp_store=np.zeros(n_iter), nchange=np.zeros(n_iter, dtype=np.int64),
               tau_sum=np.zeros((V, G, 4), dtype=np.int64), tau_star=np.zeros((V, G, 4), dtype=np.int64))
    lp_star = C.c_double(0.0)
    st = mt_state
    if use_mt and st is None:
        st = MT19937()
        lib().oracle_mt_seed(C.byref(st), seed)
    lib().oracle_update_tau(C.byref(cfg), 1 if use_mt else 0, C.byref(st) if st is not None else None,
                            tau.ctypes.data_as(_p64), pgs, pes, pv,
                            out["ll_store"].ctypes.data_as(_pd), out["lp_store"].ctypes.data_as(_pd),
                            out["nchange"].ctypes.data_as(_p64), out["tau_sum"].ctypes.data_as(_p64),
                            out["tau_star"].ctypes.data_as(_p64), C.byref(lp_star))
    out.update(tau=tau, lp_star=lp_star.value, mt_state=st)
    return out


# ---------------------------------------------------------------- NMFT
def nmft_freq(snps):
    V, S = snps.shape[0], snps.shape[1]
    snps, ps = _i64(snps)
    out = np.zeros((4 * V, S))
    lib().oracle_nmft_freq(ps, V, S, out.ctypes.data_as(_pd))
    return out


def nmft_objective(freq, tau, gamma):
    G, S = gamma.shape
    V = freq.shape[0] // 4
    freq, pf = _f64(freq)
    tau, pt = _f64(tau)
    gamma, pg = _f64(gamma)
    return lib().oracle_nmft_objective(pf, pt, pg, V, G, S)


def nmft_update(freq, tau, gamma, update_gamma=True, update_tau=True):
    """Returns new (tau, gamma) after one div_update (inputs copied)."""
    G, S = gamma.shape
    V = freq.shape[0] // 4
    freq, pf = _f64(freq)
    tau = np.array(tau, dtype=np.float64, order="C")
    gamma = np.array(gamma, dtype=np.float64, order="C")
    lib().oracle_nmft_update(pf, tau.ctypes.data_as(_pd), gamma.ctypes.data_as(_pd), V, G, S,
                             int(update_gamma), int(update_tau))
    return tau, gamma


def nmft_factorize(freq, tau0, gamma0, max_iter=5000, min_change=1e-5, fix_gamma=False):
    G, S = gamma0.shape
    V = freq.shape[0] // 4
    freq, pf = _f64(freq)
    tau = np.array(tau0, dtype=np.float64, order="C")
    gamma = np.array(gamma0, dtype=np.float64, order="C")
    trace = np.zeros(max_iter)
    div = C.c_double(0.0)
    it = lib().oracle_nmft_factorize(pf, tau.ctypes.data_as(_pd), gamma.ctypes.data_as(_pd), V, G, S,
                                     max_iter, min_change, int(fix_gamma), trace.ctypes.data_as(_pd),
                                     C.byref(div))
    return tau, gamma, it, trace[:it], div.value


def nmft_get_tau(tau, V, G):
    tau, pt = _f64(tau)
    out = np.zeros((V, G, 4), dtype=np.int64)
    lib().oracle_nmft_get_tau(pt, V, G, out.ctypes.data_as(_p64))
    return out


# ------------------------------------------------------------------ joint-state enumeration (numpy restatement, small G)
def tau_states(G):
    """tauStates [4^G,G] as base indices: Desman_Utils.cartesian order, strain 0 the slowest digit
    (HaploSNP_Sampler.py:95-103; tauMap :112-116)."""
    t = np.arange(4 ** G, dtype=np.int64)
    return np.stack([(t >> (2 * (G - 1 - g))) & 3 for g in range(G)], axis=1)


def state_logprob(variants, gamma, eta):
    """stateLogProb[n,t] = sum(log(siteProb[t]) * variants[n]) with siteProb[t] = baseProbabilityGivenTau(tauStates[t])
    (HaploSNP_Sampler.py:129-135, :239-255, :503-517)."""
    gamma, eta = np.asarray(gamma, dtype=np.float64), np.asarray(eta, dtype=np.float64)
    st = tau_states(gamma.shape[1])                                             # [T,G]
    site = np.einsum("sg,tgb->tsb", gamma, eta[st])                             # [T,S,4]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.einsum("tsb,nsb->nt", np.log(site), np.asarray(variants, dtype=np.float64))


def log_tau_prob(variants, gamma, eta, index):
    """logTauProb (:498-524): sum_v log(dP[v, index[v]]), dP = exp(lp - max) / sum."""
    lp = state_logprob(variants, gamma, eta)
    ret = 0.0
    for v in range(lp.shape[0]):
        dP = np.exp(lp[v] - np.max(lp[v]))
        dP = dP / np.sum(dP, axis=0)
        ret += np.log(dP[index[v]])
    return float(ret)


def assign_tau(variants, gamma, eta, random_state):
    """assignTau (:233-261) with the reference's own numpy draw per site; returns (base indices [N,G], conf [N])."""
    lp = state_logprob(variants, gamma, eta)
    st = tau_states(np.asarray(gamma).shape[1])
    out = np.zeros((lp.shape[0], st.shape[1]), dtype=np.uint8)
    conf = np.zeros(lp.shape[0])
    for n in range(lp.shape[0]):
        dP = np.exp(lp[n] - np.max(lp[n]))
        dP = dP / np.sum(dP, axis=0)
        t = np.flatnonzero(random_state.multinomial(1, dP, 1))[0]
        conf[n] = np.amax(dP)
        out[n] = st[t]
    return out, conf


def loglik_general(variants, tau_real, gamma, eta):
    """logLikelihood (:431-442, Desman_Utils.py:28-33) for a real-valued tau [V,G,4]."""
    from scipy.special import gammaln
    c = np.asarray(variants, dtype=np.float64)
    p = np.einsum("vga,sg,ab->vsb", np.asarray(tau_real, dtype=np.float64), gamma, eta)
    return float((gammaln(c.sum(2) + 1.0) - gammaln(c + 1.0).sum(2) + (c * np.log(p)).sum(2)).sum())


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(n)
