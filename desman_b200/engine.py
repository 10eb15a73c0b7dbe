"""Thin object wrapper around a desman_ctx handle (include/desman_b200.h, Part 2).

Host-side mirror only: all arithmetic happens in the CUDA kernels of libdesman_b200.so.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import RNG_MT19937, RNG_PHILOX, check


def auto_mu_mode(V, G):
    """The engine's default choice of the mu/E statistics contract (engine.cu resolved_mu_mode): pattern-aggregated
    binomials (1) when the ~12*2^G biallelic patterns are at most V/2, else one categorical draw per read (0)."""
    return 1 if (G <= 24 and 12.0 * 2.0 ** G <= V / 2.0) else 0


class Engine:
    """One device-resident chain: counts, tau, gamma, eta and RNG position on one GPU."""

    def __init__(self, device=0, seed=0, rng_mode=RNG_PHILOX):
        self._L = _lib.lib()
        self._h = _lib._ctx()
        check(self._L.desman_ctx_create(C.byref(self._h), device, seed & 0xFFFFFFFFFFFFFFFF, rng_mode),
              "desman_ctx_create")
        self.device = device
        self.rng_mode = rng_mode
        self.V = self.S = self.G = 0
        self.v0 = 0
        self.V_total = 0

    def close(self):
        if self._h:
            self._L.desman_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ configuration
    def set_counts(self, variants, v0=0, V_total=None):
        variants, p = _lib.as_i64(variants)
        if variants.ndim != 3 or variants.shape[2] != 4:
            raise ValueError("variants must be [V,S,4]")
        V, S = variants.shape[0], variants.shape[1]
        check(self._L.desman_set_counts(self._h, p, V, S, v0, V_total if V_total is not None else V),
              "desman_set_counts")
        self.V, self.S, self.v0 = V, S, v0
        self.V_total = V_total if V_total is not None else V

    def set_hyper(self, alpha=0.1, delta=0.1, epsilon=1e-6):
        check(self._L.desman_set_hyper(self._h, alpha, delta, epsilon), "desman_set_hyper")

    def set_rng(self, seed, sweep=0, mt_words_consumed=0):
        check(self._L.desman_set_rng(self._h, seed & 0xFFFFFFFFFFFFFFFF, sweep, mt_words_consumed), "desman_set_rng")

    def get_rng(self):
        sw, mt = C.c_uint32(0), C.c_uint64(0)
        check(self._L.desman_get_rng(self._h, C.byref(sw), C.byref(mt)), "desman_get_rng")
        return sw.value, mt.value

    def set_state(self, tau=None, gamma=None, eta=None, G=None):
        pt = pg = pe = None
        if tau is not None:
            tau, pt = _lib.as_i64(tau)
            G = tau.shape[1] if G is None else G
        if gamma is not None:
            gamma, pg = _lib.as_f64(gamma)
            G = gamma.shape[1] if G is None else G
        if eta is not None:
            eta, pe = _lib.as_f64(eta)
        if G is None:
            G = self.G
        check(self._L.desman_set_state(self._h, pt, pg, pe, G), "desman_set_state")
        self.G = G

    def set_tau_index(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.uint8)
        check(self._L.desman_set_tau_index(self._h, idx.ctypes.data_as(_lib._pu8), idx.shape[1]),
              "desman_set_tau_index")
        self.G = idx.shape[1]

    def get_tau_index(self):
        out = np.empty((self.V, self.G), dtype=np.uint8)
        check(self._L.desman_get_tau_index(self._h, out.ctypes.data_as(_lib._pu8)), "desman_get_tau_index")
        return out

    def get_state(self, want_tau=True):
        tau = np.empty((self.V, self.G, 4), dtype=np.int64) if want_tau else None
        gamma = np.empty((self.S, self.G))
        eta = np.empty((4, 4))
        check(self._L.desman_get_state(self._h, _lib.ptr_i64(tau), _lib.ptr_d(gamma), _lib.ptr_d(eta)),
              "desman_get_state")
        return tau, gamma, eta

    # ------------------------------------------------------------------ single steps
    def sample_tau(self):
        n = C.c_int64(0)
        check(self._L.desman_sample_tau(self._h, C.byref(n)), "desman_sample_tau")
        return n.value

    def sample_tau_fix(self, H, want_logp=True):
        """sampleTauFixTau (:196-222): redraw strains [H, G); returns (logp [V,4] or None, nchange)."""
        logp = np.empty((self.V, 4)) if want_logp else None
        n = C.c_int64(0)
        check(self._L.desman_sample_tau_fix(self._h, int(H), _lib.ptr_d(logp), C.byref(n)), "desman_sample_tau_fix")
        return logp, n.value

    def mu_stats(self):
        sm = np.zeros((self.S, self.G), dtype=np.int64)
        es = np.zeros((4, 4), dtype=np.int64)
        check(self._L.desman_mu_stats(self._h, _lib.ptr_i64(sm), _lib.ptr_i64(es)), "desman_mu_stats")
        return sm, es

    def draw_gamma_eta(self, sum_mu, esum):
        sum_mu, pm = _lib.as_i64(sum_mu)
        esum, pe = _lib.as_i64(esum)
        g = np.empty((self.S, self.G))
        e = np.empty((4, 4))
        check(self._L.desman_draw_gamma_eta(self._h, pm, pe, _lib.ptr_d(g), _lib.ptr_d(e)), "desman_draw_gamma_eta")
        return g, e

    def loglik(self):
        ll, lp = C.c_double(0), C.c_double(0)
        check(self._L.desman_loglik(self._h, C.byref(ll), C.byref(lp)), "desman_loglik")
        return ll.value, lp.value

    def loglik_general(self, tau, gamma, eta):
        """logLikelihood (:431-442) under a real-valued tau [V,G,4] (the tauMean of DIC, :486-496)."""
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        eta = np.ascontiguousarray(eta, dtype=np.float64)
        if tau.ndim != 3 or tau.shape[0] != self.V or tau.shape[2] != 4 or gamma.shape != (self.S, tau.shape[1]) or eta.shape != (4, 4):
            raise ValueError("loglik_general: tau [V,G,4], gamma [S,G], eta [4,4] expected")
        ll = C.c_double(0)
        check(self._L.desman_loglik_general(self._h, _lib.ptr_d(tau), _lib.ptr_d(gamma), _lib.ptr_d(eta), tau.shape[1], C.byref(ll)),
              "desman_loglik_general")
        return ll.value

    def state_logprob(self, gamma, eta, variants=None, index=None, want_logprob=False):
        """Log-probabilities of all 4^G joint states per site (assignTau :233-261, logTauProb :498-524).
        variants: int64 [N,S,4] or None for the counts of the engine.  Returns a dict: maxlp [N], lse [N] (log sum exp over the
        states), argmax [N], and, if asked for, lp_at_index [N] (index [N] given) and logprob [N,4^G]."""
        gamma = np.ascontiguousarray(gamma, dtype=np.float64)
        eta = np.ascontiguousarray(eta, dtype=np.float64)
        if gamma.ndim != 2 or eta.shape != (4, 4):
            raise ValueError("state_logprob: gamma [S,G], eta [4,4] expected")
        S, G = gamma.shape
        pv, N = None, self.V
        if variants is not None:
            variants, pv = _lib.as_i64(variants, "variants")
            if variants.ndim != 3 or variants.shape[1] != S or variants.shape[2] != 4:
                raise ValueError("state_logprob: variants [N,%d,4] expected" % S)
            N = variants.shape[0]
        elif S != self.S:
            raise ValueError("state_logprob: gamma has %d samples, the engine's counts %d" % (S, self.S))
        out = dict(maxlp=np.empty(N), lse=np.empty(N), argmax=np.empty(N, dtype=np.int64))
        pi = pl = pf = None
        if index is not None:
            index, pi = _lib.as_i64(index, "index")
            if index.shape != (N,):
                raise ValueError("state_logprob: index [N] expected")
            out["lp_at_index"] = np.empty(N)
            pl = _lib.ptr_d(out["lp_at_index"])
        if want_logprob:
            out["logprob"] = np.empty((N, 4 ** G))
            pf = _lib.ptr_d(out["logprob"])
        check(self._L.desman_state_logprob(self._h, pv, N, S, _lib.ptr_d(gamma), _lib.ptr_d(eta), G, pi, pf, pl,
                                           _lib.ptr_d(out["maxlp"]), _lib.ptr_d(out["lse"]), _lib.ptr_i64(out["argmax"])),
              "desman_state_logprob")
        return out

    # ------------------------------------------------------------------ chains
    def update(self, n_iter):
        S, G = self.S, self.G
        out = dict(gamma_store=np.zeros((n_iter, S, G)), eta_store=np.zeros((n_iter, 4, 4)),
                   ll_store=np.zeros(n_iter), lp_store=np.zeros(n_iter), nchange=np.zeros(n_iter, dtype=np.int64))
        check(self._L.desman_update(self._h, n_iter, _lib.ptr_d(out["gamma_store"]), _lib.ptr_d(out["eta_store"]),
                                    _lib.ptr_d(out["ll_store"]), _lib.ptr_d(out["lp_store"]),
                                    _lib.ptr_i64(out["nchange"])), "desman_update")
        return out

    def update_tau(self, gamma_store, eta_store):
        gamma_store, pg = _lib.as_f64(gamma_store)
        eta_store, pe = _lib.as_f64(eta_store)
        n_iter = gamma_store.shape[0]
        out = dict(ll_store=np.zeros(n_iter), lp_store=np.zeros(n_iter), nchange=np.zeros(n_iter, dtype=np.int64))
        check(self._L.desman_update_tau(self._h, n_iter, pg, pe, _lib.ptr_d(out["ll_store"]),
                                        _lib.ptr_d(out["lp_store"]), _lib.ptr_i64(out["nchange"])),
              "desman_update_tau")
        return out

    def get_star(self, want_tau=True):
        tau = np.empty((self.V, self.G, 4), dtype=np.int64) if want_tau else None
        gamma = np.empty((self.S, self.G))
        eta = np.empty((4, 4))
        lp, it = C.c_double(0), C.c_int(0)
        check(self._L.desman_get_star(self._h, _lib.ptr_i64(tau), _lib.ptr_d(gamma), _lib.ptr_d(eta), C.byref(lp),
                                      C.byref(it)), "desman_get_star")
        return dict(tau=tau, gamma=gamma, eta=eta, lp=lp.value, iter=it.value)

    def get_esum_store(self, n_iter):
        """E_store[i].sum(axis=(0,1)) [n_iter,4,4] of the last update() (Esum[a_obs,b_true] per sweep)."""
        out = np.empty((n_iter, 4, 4), dtype=np.int64)
        check(self._L.desman_get_esum_store(self._h, _lib.ptr_i64(out)), "desman_get_esum_store")
        return out

    def get_star_index(self):
        out = np.empty((self.V, self.G), dtype=np.uint8)
        check(self._L.desman_get_star_index(self._h, out.ctypes.data_as(_lib._pu8)), "desman_get_star_index")
        return out

    def get_tau_sum(self, compact=False):
        """tau_store.sum(axis=0) of the last update()/update_tau(); compact=True: the device's uint32 counters as they are."""
        if compact:
            out = np.empty((self.V, self.G, 4), dtype=np.uint32)
            check(self._L.desman_get_tau_sum_u32(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))), "desman_get_tau_sum_u32")
            return out
        out = np.empty((self.V, self.G, 4), dtype=np.int64)
        check(self._L.desman_get_tau_sum(self._h, _lib.ptr_i64(out)), "desman_get_tau_sum")
        return out

    def nmft_factorize(self, snps, tau0, gamma0, max_iter=5000, min_change=1e-5, fix_gamma=False, want_trace=False):
        snps, ps = _lib.as_i64(snps)
        V, S = snps.shape[0], snps.shape[1]
        tau = np.array(tau0, dtype=np.float64, order="C")
        gamma = np.array(gamma0, dtype=np.float64, order="C")
        G = gamma.shape[0]
        if tau.shape != (4 * V, G) or gamma.shape != (G, S):
            raise ValueError("tau must be [4V,G] and gamma [G,S]")
        trace = np.zeros(max(max_iter, 1)) if want_trace else None
        it, div = C.c_int(0), C.c_double(0)
        check(self._L.desman_nmft_factorize(self._h, ps, V, S, G, _lib.ptr_d(tau), _lib.ptr_d(gamma), max_iter,
                                            min_change, int(fix_gamma), C.byref(it), C.byref(div), _lib.ptr_d(trace)),
              "desman_nmft_factorize")
        return tau, gamma, it.value, div.value, (trace[:it.value] if want_trace else None)

    # ------------------------------------------------------------------ multi-GPU / measurement
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        check(_lib.lib().desman_comm_unique_id(buf), "desman_comm_unique_id")
        return buf.raw

    def comm_kind(self):
        """Data plane of the per-sweep exchange: 'none', 'nccl-allreduce' or 'p2p-mailbox' (one-shot over NVLink peer memory)."""
        return ("none", "nccl-allreduce", "p2p-mailbox")[self._L.desman_comm_kind(self._h)]

    @staticmethod
    def nmft_last_timing():
        ms, it = C.c_double(0), C.c_int(0)
        check(_lib.lib().desman_nmft_last_timing(C.byref(ms), C.byref(it)), "desman_nmft_last_timing")
        return ms.value, it.value

    def comm_init(self, uid, rank, nranks):
        check(self._L.desman_comm_init(self._h, uid, rank, nranks), "desman_comm_init")

    def set_option(self, name, value):
        check(self._L.desman_set_option(self._h, name.encode(), int(value)), "desman_set_option")

    def get_group_stats(self):
        """Site groups of the tau screening pass: dict(have, calm, items, singles, work, orphans, slots, configured)."""
        out = np.zeros(8, dtype=np.int64)
        check(self._L.desman_get_group_stats(self._h, _lib.ptr_i64(out)), "desman_get_group_stats")
        d = dict(zip(("have", "calm", "items", "singles", "work", "orphans", "slots", "configured"), out.tolist()))
        d["worth"] = (d["configured"] >> 1) & 1          # the realised groups pay for a table each (decided on the device)
        d["configured"] &= 1
        return d

    def debug_screen(self):
        """Validation of the tensor-memory screening pass: (D [V,G,3] float32 log2-units sums, NaN where a site is in no
        group; mask [V] uint32 of undecided strains, 0xffffffff where the site is not on the work list)."""
        D = np.empty((self.V, self.G, 3), dtype=np.float32)
        mask = np.empty(self.V, dtype=np.uint32)
        check(self._L.desman_debug_screen(self._h, D.ctypes.data_as(C.POINTER(C.c_float)),
                                          mask.ctypes.data_as(C.POINTER(C.c_uint32))), "desman_debug_screen")
        return D, mask

    def get_tier_counts(self, reset=True):
        out = np.zeros(3, dtype=np.int64)
        check(self._L.desman_get_tier_counts(self._h, _lib.ptr_i64(out), int(reset)), "desman_get_tier_counts")
        return out

    def set_profiling(self, per_kernel_events=False, flush_l2=False):
        check(self._L.desman_set_profiling(self._h, int(per_kernel_events), int(flush_l2)), "desman_set_profiling")   # per_kernel_events: 0, 1 or 2 (see the header)

    def get_timing(self):
        el = C.c_double(0)
        kms = (C.c_double * len(_lib.K_NAMES))()
        kl = (C.c_int64 * len(_lib.K_NAMES))()
        check(self._L.desman_get_timing(self._h, C.byref(el), kms, kl), "desman_get_timing")
        return dict(elapsed_ms=el.value, kernel_ms=dict(zip(_lib.K_NAMES, list(kms))),
                    kernel_launches=dict(zip(_lib.K_NAMES, list(kl))))

    def synchronize(self):
        check(self._L.desman_synchronize(self._h), "desman_synchronize")
