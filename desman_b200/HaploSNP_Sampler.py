"""Mirror of the reference class desman/HaploSNP_Sampler.py (same constructor, attributes and
method names) whose sweeps run on the GPU through libdesman_b200.so.

What differs from the reference, and why:
  * update()/updateTau() keep tau, gamma, eta and the count tensor resident on the device for the
    whole call; the externally assignable attributes (tau, gamma, eta, gamma_store, eta_store,
    bin/desman:140-146,194-199) are uploaded on entry and written back on exit.
  * mu[V,S,4,G] / E[V,S,4,4] are never materialised -- only their sums are consumed
    (HaploSNP_Sampler.py:266,276).  mu_store/E_store/tau_store/tauStates (0.8 TB / 0.4 TB / 13 GB /
    2 TB at the BASELINE configs, :89-103) are not allocated; tauMean()/probabilisticTau() come
    from per-site occupancy counters kept by the tau kernel.
  * the mu/gamma/eta draws follow the Philox counter contract instead of numpy's sequential legacy
    stream (which has no parallel form); tau draws use Philox too, or the GSL-compatible MT19937
    stream when tau_rng="mt19937".
"""
import logging
import sys

import numpy as np

from . import sampletau as _sampletau
from ._lib import RNG_MT19937, RNG_PHILOX
from .engine import Engine


def du_log_dirichlet(x, alpha):
    """Desman_Utils.log_dirichlet_pdf (Desman_Utils.py:35-44)."""
    from scipy.special import gammaln
    ret = gammaln(np.sum(alpha))
    for i in range(len(alpha)):
        ret += (alpha[i] - 1.0) * np.log(x[i])
        ret -= gammaln(alpha[i])
    return float(ret)


class Constants(object):
    MAX_LOG_DIR_PROB = 100.0


class HaploSNP_Sampler():

    def __init__(self, snps, G, randomState, fixed_tau=None, burn_iter=None, max_iter=None, alpha_constant=0.1,
                 delta_constant=0.1, epsilon=1.0e-6, device=0, seed=None, tau_rng="philox", shard=None, comm=None):
        # reference defaults, HaploSNP_Sampler.py:33-43
        self.burn_iter = 250 if burn_iter is None else burn_iter
        self.max_iter = 250 if max_iter is None else max_iter
        self.tau_comp_iter = 10

        self.randomState = randomState
        self.G = G
        self.V = snps.shape[0]
        self.S = snps.shape[1]
        self.variants = np.copy(snps, order='C').astype(np.int64, copy=False)
        self.epsilon = epsilon

        self.delta = np.empty(4); self.delta.fill(delta_constant)
        self.delta_constant = delta_constant
        self.alpha = np.empty(self.G); self.alpha.fill(alpha_constant)
        self.alpha_constant = alpha_constant

        # the constructor consumes the caller's stream exactly like the reference (:63, :72)
        self._tau_oh = self._tau_ix = self._tau_star_oh = self._tau_star_ix = None
        self._tauIndices = self._tauIndices_star = None
        self.gamma = self.randomState.dirichlet(self.alpha, size=self.S)
        self.gamma_store = np.zeros((self.max_iter, self.S, self.G))
        if fixed_tau is None:
            tri = self.randomState.randint(0, 4, self.V * self.G)
            self.tau = np.zeros((self.V, self.G, 4), dtype=np.int64)
            np.put_along_axis(self.tau, np.reshape(tri, (self.V, self.G, 1)).astype(np.int64), 1, axis=2)
        else:
            self.tau = np.reshape(fixed_tau, (self.V, self.G, 4))
        self.tauIndices = np.zeros((self.V), dtype=np.int64)

        self.eta = 0.96 * np.identity((4)) + 0.01 * np.ones((4, 4))        # :84
        self.eta_store = np.zeros((self.max_iter, 4, 4))

        self.amatrix = np.identity(4, dtype=np.int64)
        self.ll = 0.0
        self.lp = 0.0
        self.ll_store = np.zeros(self.max_iter)
        self.lp_store = np.zeros(self.max_iter)
        self.nchange_store = np.zeros(self.max_iter, dtype=np.int64)
        self.nTauStates = 4 ** self.G
        self._set_tau_map()

        self._tau_sum = None          # tau_store.sum(axis=0) of the last update()/updateTau()
        self._Esum_store = None       # E_store[i].sum(axis=(0,1)) per sweep of the last update()
        self._sum_mu = None           # mu.sum(axis=(0,2)) of the last sampleMu()
        self._Esum = None             # E.sum(axis=(0,1))
        self._timing = None

        # device side
        self._device = device
        self._seed = _sampletau.current_seed() if seed is None else int(seed)
        self._tau_rng = tau_rng
        self._shard = shard           # (v0, V_total) when this object holds one V-shard of a larger chain
        self._comm = comm             # (uid, rank, nranks)
        self._eng = None
        self._eng_mode = None
        self._counts_up = False
        if comm is not None:          # a rank of a sharded chain: context and communicator are set up with the object;
            self._context()           # counts and state still travel with the first driver call

    # ------------------------------------------------------------------ tau / tau_star: one-hot views built on demand
    # The engine keeps tau as uint8 base indices [V,G]; the reference's int64 one-hot [V,G,4] layout (32x larger) is
    # materialised only when the attribute is read, and whatever the caller assigns is uploaded at the next driver.
    @staticmethod
    def _onehot(idx):
        out = np.zeros(idx.shape + (4,), dtype=np.int64)
        np.put_along_axis(out, idx[..., None].astype(np.int64), 1, axis=-1)
        return out

    @property
    def tau(self):
        if self._tau_oh is None:
            self._tau_oh = self._onehot(self._tau_ix)
        return self._tau_oh

    @tau.setter
    def tau(self, value):
        self._tau_oh = value
        self._tau_ix = None

    @property
    def tau_star(self):
        if self._tau_star_oh is None:
            self._tau_star_oh = self._onehot(self._tau_star_ix)
        return self._tau_star_oh

    @tau_star.setter
    def tau_star(self, value):
        self._tau_star_oh = value
        self._tau_star_ix = None

    def _tau_index(self, star=False):
        ix, oh = (self._tau_star_ix, self._tau_star_oh) if star else (self._tau_ix, self._tau_oh)
        if oh is not None:          # the caller may have assigned or modified the one-hot array
            return np.argmax(np.asarray(oh), axis=2).astype(np.uint8)
        return ix

    # ------------------------------------------------------------------ device plumbing
    def _context(self, mode=RNG_PHILOX):
        """Device context (+ communicator of a sharded chain) without any data on it."""
        if self._eng is None:
            self._eng = Engine(self._device, self._seed, mode)
            if self._comm is not None:
                self._eng.comm_init(*self._comm)
            self._eng.set_rng(self._seed, sweep=_sampletau.global_sweep())
            self._eng_mode = mode
            self._counts_up = False
        elif self._eng_mode != mode:
            # ONE context for the life of the object: the stream of the tau draws is a per-call choice.  The Philox sweep
            # counter and the position in the MT19937 stream both survive the switch (alternating sampleMu / update with
            # sampleTau / updateTau under tau_rng="mt19937" continues both streams), the counts stay on the device and the
            # communicator of a sharded sampler is not initialised twice.
            self._eng.set_option("rng_mode", mode)
            self._eng.rng_mode = mode
            self._eng_mode = mode
        return self._eng

    def _engine(self, mode=RNG_PHILOX):
        eng = self._context(mode)
        if not self._counts_up:
            v0, vt = self._shard if self._shard is not None else (0, self.V)
            eng.set_counts(self.variants, v0=v0, V_total=vt)
            self._counts_up = True
        eng.set_hyper(self.alpha_constant, self.delta_constant, self.epsilon)
        return eng

    def _push(self, eng, gamma=None, tau=None, eta=None):
        gamma = self.gamma if gamma is None else gamma
        eta = self.eta if eta is None else eta
        if gamma.shape[1] != self.G:
            raise ValueError("gamma does not match G = %d" % self.G)
        if tau is None and self._tau_oh is None and self._tau_ix is not None:
            if self._tau_ix.shape[1] != self.G:
                raise ValueError("tau does not match G = %d" % self.G)
            eng.set_state(None, gamma, eta, G=self.G)
            eng.set_tau_index(self._tau_ix)                       # untouched since the last driver: upload the indices
            return
        tau = self.tau if tau is None else tau
        if tau.shape[1] != self.G:
            raise ValueError("tau does not match G = %d" % self.G)
        eng.set_state(np.ascontiguousarray(tau, dtype=np.int64), gamma, eta, G=self.G)

    def _pull(self, eng, full=True):
        # the process-global sweep counter follows every driver, so a sampler created while this one is still open (the
        # not-selected sampler of bin/desman -r) starts past the counters this chain has used
        _sampletau.advance_global_sweep(eng.get_rng()[0])
        self._tau_ix, self._tau_oh = eng.get_tau_index(), None
        if full:
            _, self.gamma, self.eta = eng.get_state(want_tau=False)

    def close(self):
        if self._eng is not None:
            _sampletau.advance_global_sweep(self._eng.get_rng()[0])
            self._eng.close()
            self._eng = None

    def _set_tau_map(self):
        self.tauMap = np.zeros((self.G, 4), dtype=object if self.G > 31 else np.int64)
        for g in range(self.G):
            for a in range(4):
                self.tauMap[g, a] = a * (4 ** (self.G - g - 1))              # :115-117

    # ------------------------------------------------------------------ small helpers of the reference
    def calcK(self):
        return self.V * self.G + self.S * (self.G - 1)

    def mapTauState(self, tauState):
        return np.einsum('ga,ga', self.tauMap, tauState)                     # :224-226

    def updateTauIndices(self):
        self.tauIndices = self._site_codes(self._tau_index())               # :228-231, vectorised

    # tauIndices / tauIndices_star are derived from tau / tau_star (2 ms per 1e5 sites on the host): after a chain driver they
    # are worked out when first read, not inside the driver
    @property
    def tauIndices(self):
        if self._tauIndices is None:
            self._tauIndices = self._site_codes(self._tau_index())
        return self._tauIndices

    @tauIndices.setter
    def tauIndices(self, value):
        self._tauIndices = value

    @property
    def tauIndices_star(self):
        if self._tauIndices_star is None:
            self._tauIndices_star = self._site_codes(self._tau_index(star=True))
        return self._tauIndices_star

    @tauIndices_star.setter
    def tauIndices_star(self, value):
        self._tauIndices_star = value

    def _site_codes(self, idx):
        if self.G > 31:
            w = np.array([4 ** (self.G - g - 1) for g in range(self.G)], dtype=object)
            return idx.astype(object) @ w
        code = np.zeros(idx.shape[0], dtype=np.int64)                      # base-4 digits, strain 0 most significant
        for g in range(self.G):
            code <<= 2
            code |= idx[:, g]
        return code

    def tauDist(self, tau1, tau2):
        return int((np.argmax(tau1, axis=1) != np.argmax(tau2, axis=1)).sum())

    def baseProbabilityGivenTau(self, tauState, gamma, eta):
        return np.einsum('jk,lj,km->lm', tauState, gamma, eta)               # :129-135

    def storeStarState(self, iter):
        self.gamma_star = np.copy(self.gamma)
        self.tau_star = np.copy(self.tau)
        self.tauIndices_star = np.copy(self.tauIndices)
        self.eta_star = np.copy(self.eta)
        self.iter_star = iter
        self.lp_star = self.lp

    # ------------------------------------------------------------------ single Gibbs steps
    def sampleTau(self, gamma=None, eta=None):
        """One tau update with (gamma, eta); returns the number of changed (v,g) entries (:148-184)."""
        eng = self._engine(RNG_MT19937 if self._tau_rng == "mt19937" else RNG_PHILOX)
        self._push(eng, gamma=gamma, eta=eta)
        n = eng.sample_tau()
        self._pull(eng, full=False)
        self.updateTauIndices()
        return n

    def sampleMu(self, tauC, gammaC, etaC):
        """mu/E data augmentation (:284-309); only sum_mu[S,G] and Esum[4,4] are produced."""
        eng = self._engine()
        self._push(eng, gamma=gammaC, tau=tauC, eta=etaC)
        self._sum_mu, self._Esum = eng.mu_stats()
        return self._sum_mu, self._Esum

    def sampleGamma(self):
        """gamma[s,:] ~ Dir(alpha + sum_mu[s,:]), clip at epsilon, renormalise (:263-273)."""
        eng = self._engine()
        self._push(eng)
        self.gamma, self._eta_draw = eng.draw_gamma_eta(self._sum_mu, self._Esum)

    def sampleEta(self):
        """eta[a,:] ~ Dir(delta + Esum[:,a]) (:275-281); drawn together with gamma from the same statistics."""
        if getattr(self, "_eta_draw", None) is None:
            eng = self._engine()
            self._push(eng)
            _, self._eta_draw = eng.draw_gamma_eta(self._sum_mu, self._Esum)
        self.eta = self._eta_draw
        self._eta_draw = None

    def logLikelihood(self, cGamma, cTau, cEta):
        """Data log likelihood (:431-442)."""
        return self._ll_lp(cGamma, cTau, cEta)[0]

    def logPosterior(self, cGamma, cTau, cEta):
        """Log posterior (:444-461)."""
        return self._ll_lp(cGamma, cTau, cEta)[1]

    def _ll_lp(self, cGamma, cTau, cEta):
        cTau = np.asarray(cTau)
        if not (((cTau == 0) | (cTau == 1)).all() and (cTau.sum(axis=2) == 1).all()):
            # a real-valued tau (the tauMean of DIC, :486-496): general kernel; the tau prior term of lp (:459) is the same
            eng = self._engine()
            self._push(eng)
            ll = eng.loglik_general(cTau, cGamma, cEta)
            prior = sum(du_log_dirichlet(np.asarray(cGamma)[s], self.alpha) for s in range(self.S))
            prior += sum(du_log_dirichlet(np.asarray(cEta)[a], self.delta) for a in range(4))
            return ll, ll + prior + self.V * self.G * np.log(0.25)
        eng = self._engine()
        self._push(eng, gamma=np.asarray(cGamma), tau=cTau, eta=np.asarray(cEta))
        return eng.loglik()

    # ------------------------------------------------------------------ chain drivers
    def _log_progress(self, res, what):
        for it in range(0, len(res["nchange"]), 10):                           # :360-361, :402-403
            logging.info('Gibbs Iter %d, no. changed = %d, %s = %f' % (it, res["nchange"][it], what, res["lp_store"][it]))

    def _finish(self, eng, res, n_iter, full):
        self._pull(eng, full)
        star = eng.get_star(want_tau=False)
        self._tau_star_ix, self._tau_star_oh = eng.get_star_index(), None
        self.lp_star = star["lp"]
        self.iter_star = star["iter"]
        if full:
            self.gamma_star, self.eta_star = star["gamma"], star["eta"]
        self.ll_store = res["ll_store"]
        self.lp_store = res["lp_store"]
        self.nchange_store = res["nchange"]
        if n_iter > 0:
            self.ll, self.lp = float(res["ll_store"][-1]), float(res["lp_store"][-1])
        self._tau_sum = eng.get_tau_sum(compact=True)        # uint32 occupancy counters; tauMean() divides them
        self._timing = eng.get_timing()
        self._tauIndices = self._tauIndices_star = None                       # lazily, from the new tau / tau_star

    def update(self):
        """max_iter Gibbs sweeps mu/E -> gamma -> tau -> eta -> ll/lp with MAP tracking (:334-365)."""
        eng = self._engine()
        self._push(eng)
        res = eng.update(self.max_iter)
        self.gamma_store, self.eta_store = res["gamma_store"], res["eta_store"]
        self._Esum_store = eng.get_esum_store(self.max_iter) if self.max_iter > 0 else None
        self._finish(eng, res, self.max_iter, True)
        self._log_progress(res, "nlp")

    def update_fixed_tau(self):
        """max_iter sweeps of mu/E -> gamma -> eta with tau held fixed (:409-428)."""
        eng = self._engine()
        self._push(eng)
        eng.set_option("fixed_tau", 1)
        try:
            res = eng.update(self.max_iter)
        finally:
            eng.set_option("fixed_tau", 0)
        self.gamma_store, self.eta_store = res["gamma_store"], res["eta_store"]
        self._finish(eng, res, self.max_iter, True)

    def updateTau(self):
        """tau-only replay against gamma_store/eta_store (:383-407)."""
        if self.gamma_store.shape[0] < self.max_iter or self.eta_store.shape[0] < self.max_iter:
            raise ValueError("gamma_store/eta_store hold fewer than max_iter iterations")
        eng = self._engine(RNG_MT19937 if self._tau_rng == "mt19937" else RNG_PHILOX)
        self._push(eng, gamma=self.gamma_store[0], eta=self.eta_store[0])
        res = eng.update_tau(self.gamma_store[:self.max_iter], self.eta_store[:self.max_iter])
        self._finish(eng, res, self.max_iter, False)
        self._log_progress(res, "nll")
        sys.stdout.flush()

    def burn(self):
        """burn_iter sweeps printing `iter ll lp` (:313-324)."""
        eng = self._engine()
        self._push(eng)
        res = eng.update(self.burn_iter)
        self._pull(eng)
        for it in range(self.burn_iter):
            print(str(it) + " " + str(res["ll_store"][it]) + " " + str(res["lp_store"][it]))
        if self.burn_iter > 0:
            self.ll, self.lp = float(res["ll_store"][-1]), float(res["lp_store"][-1])

    def burnTau(self):
        """burn_iter tau-only updates at (gamma_star, eta_star) (:367-380)."""
        gs = np.tile(self.gamma_star, (self.burn_iter, 1, 1))
        es = np.tile(self.eta_star, (self.burn_iter, 1, 1))
        eng = self._engine(RNG_MT19937 if self._tau_rng == "mt19937" else RNG_PHILOX)
        self._push(eng, gamma=self.gamma_star, eta=self.eta_star)
        res = eng.update_tau(gs, es)
        self._pull(eng, full=False)
        for it in range(self.burn_iter):
            print(str(it) + "," + str(res["nchange"][it]) + "," + str(res["lp_store"][it]))
            sys.stdout.flush()
        if self.burn_iter > 0:
            self.ll, self.lp = float(res["ll_store"][-1]), float(res["lp_store"][-1])

    # ------------------------------------------------------------------ summaries
    def meanDeviance(self):
        return -2.0 * np.mean(self.ll_store)                                   # :463-465

    def gammaMean(self):
        return np.mean(self.gamma_store, axis=0)                               # :467-471

    def etaMean(self):
        return np.mean(self.eta_store, axis=0)                                 # :473-477

    def tauMean(self):
        """np.mean(tau_store, axis=0) (:479-483) from the kernel's occupancy counters."""
        if self._tau_sum is None:
            return np.zeros((self.V, self.G, 4))
        return self._tau_sum / float(self.max_iter)

    def probabilisticTau(self):
        return self.tauMean()                                                  # :834-840

    @property
    def tau_store(self):
        raise AttributeError("tau_store[max_iter,V,G,4] is not materialised by desman_b200 "
                             "(use tauMean()/probabilisticTau(); see DESIGN.md)")

    mu_store = E_store = tauStates = tau_store

    # ------------------------------------------------------------------ host-side bookkeeping (V*G^2 integer work)
    def calculateSND(self, tau):
        """Pairwise single-nucleotide differences between strains (:712-730)."""
        idx = np.argmax(np.asarray(tau), axis=2)
        G = idx.shape[1]
        snd = np.zeros((G, G), dtype=np.int64)
        for g in range(G):
            snd[g, :] = (idx != idx[:, g:g + 1]).sum(axis=0)
        return snd

    def compSND(self, tau1, tau2):
        i1, i2 = np.argmax(np.asarray(tau1), axis=2), np.argmax(np.asarray(tau2), axis=2)
        snd = np.zeros((i1.shape[1], i2.shape[1]), dtype=np.int64)
        for g in range(i1.shape[1]):
            snd[g, :] = (i2 != i1[:, g:g + 1]).sum(axis=0)
        return snd

    def variableTau(self, tau):
        idx = np.argmax(np.asarray(tau), axis=2)
        return (idx != idx[:, :1]).any(axis=1)

    def removeDegenerate(self):
        """Merge identical haplotypes (SND == 0), adding their gamma columns (:771-832)."""
        snd = self.calculateSND(self.tau)
        deleted = np.zeros(self.G, dtype=bool)
        allmapped = []
        for g in range(self.G):
            gmap = []
            for h in range(g + 1, self.G):
                if not deleted[h] and snd[g, h] == 0:
                    deleted[h] = True
                    gmap.append(h)
            allmapped.append(gmap)
        NU = self.G - int(deleted.sum())
        tau_new = np.zeros((self.V, NU, 4), dtype=np.int64)
        gamma_new = np.zeros((self.S, NU))
        k = 0
        for g in range(self.G):
            if not deleted[g]:
                tau_new[:, k, :] = self.tau[:, g, :]
                gamma_new[:, k] = self.gamma[:, g]
                for h in allmapped[g]:
                    gamma_new[:, k] += self.gamma[:, h]
                k += 1
        self.gamma = gamma_new
        self.tau = tau_new
        self.G = NU
        self.alpha = np.empty(self.G); self.alpha.fill(self.alpha_constant)
        self.gamma_store = np.zeros((self.max_iter, self.S, self.G))          # stores are re-allocated (:809-814)
        self._tau_sum = None
        self.nTauStates = 4 ** self.G
        self._set_tau_map()
        self.updateTauIndices()

    # ------------------------------------------------------------------ joint-state enumeration
    def sampleLogProb(self, adLogProbS):
        dP = np.exp(adLogProbS - np.max(adLogProbS))                           # :124-127
        dP = dP / np.sum(dP, axis=0)
        return np.flatnonzero(self.randomState.multinomial(1, dP, 1))[0]

    def assignTau(self, assignMatrix):
        """Tau for new sets of variants (:233-261): the log-probabilities of all 4^G joint states per site come from the
        device (FP64 tiled product); the draw per site is the reference's numpy multinomial from self.randomState, in site order."""
        assignMatrix = np.asarray(assignMatrix)
        N = assignMatrix.shape[0]
        variants = np.ascontiguousarray(np.reshape(assignMatrix, (N, self.S, 4)), dtype=np.int64)
        eng = self._engine()
        self._push(eng)
        T = 4 ** self.G
        assign = np.zeros((N, self.G, 4), dtype=np.int64)
        conf = np.zeros(N)
        shifts = 2 * (self.G - 1 - np.arange(self.G))
        step = max(1, (1 << 25) // T)                                          # <= 256 MB of log-probabilities at a time
        for lo in range(0, N, step):
            lp = eng.state_logprob(self.gamma_star, self.eta_star, variants=variants[lo:lo + step], want_logprob=True)["logprob"]
            for i in range(lp.shape[0]):
                dP = np.exp(lp[i] - np.max(lp[i]))
                dP = dP / np.sum(dP, axis=0)
                t = np.flatnonzero(self.randomState.multinomial(1, dP, 1))[0]
                conf[lo + i] = np.amax(dP)
                assign[lo + i, np.arange(self.G), (int(t) >> shifts) & 3] = 1      # tauStates[t] (:95-103)
        return (assign, conf)

    def logTauProb(self, cGamma, cEta):
        """sum_v log P(tau_star[v] | counts[v], gamma, eta) over the 4^G joint states of a site (:498-524)."""
        eng = self._engine()
        self._push(eng)
        idx = np.ascontiguousarray(self.tauIndices_star, dtype=np.int64)
        r = eng.state_logprob(cGamma, cEta, index=idx)
        return float(np.sum(r["lp_at_index"] - r["lse"]))

    def DIC(self):
        return self.meanDeviance() + 2.0 * self.logLikelihood(self.gammaMean(), self.tauMean(), self.etaMean())   # :486-496

    # ------------------------------------------------------------------ Chib's marginal likelihood (:538-710)
    # Compositions of the device primitives above under the Philox counter contract: the reference
    # draws every step from numpy's sequential stream, so its estimates are reproduced in law, not draw for draw; the same
    # compositions of the CPU oracle's primitives give these values to 1e-9 (tests/test_gpu_states.py).
    def normaliseLogProb(self, logProb):
        logProb = logProb - np.max(logProb)                                    # :186-194
        return logProb - np.log(np.exp(logProb).sum())

    def logMean(self, logStore):
        maxLog = np.max(logStore)                                              # :526-536
        return maxLog + np.log(np.exp(logStore - maxLog).sum()) - np.log(logStore.shape[0])

    def tauOne(self, tauSlice):
        return int(np.argmax(np.asarray(tauSlice) == 1))                       # :610-618

    def sampleTauFixTau(self, fixedTau, H, gammaStar, etaStar):
        """Redraw strains [H, G) of every site in order at (gammaStar, etaStar), in place on fixedTau; returns the normalised
        log-probabilities [V,4] of strain H's bases before its draw (:196-222).  One launch of the per-site tau kernel."""
        eng = self._engine()
        self._push(eng, gamma=gammaStar, tau=fixedTau, eta=etaStar)
        logp, _ = eng.sample_tau_fix(H)
        _sampletau.advance_global_sweep(eng.get_rng()[0])
        idx = eng.get_tau_index()
        fixedTau[...] = 0
        np.put_along_axis(fixedTau, idx[..., None].astype(np.int64), 1, axis=2)
        return logp

    def _log_gamma_post(self, sum_mu):
        return sum(du_log_dirichlet(self.gamma_star[s, :], self.alpha + sum_mu[s, :]) for s in range(self.S))

    def _log_eta_post(self, sum_E):
        return sum(du_log_dirichlet(self.eta_star[a, :], self.delta + sum_E[:, a]) for a in range(4))

    def chibMarginalLogLikelihood(self):
        """Chib's estimator with the joint-state tau term (:621-710)."""
        cMLogL = self.logLikelihood(self.gamma_star, self.tau_star, self.eta_star)
        for s in range(self.S):
            cMLogL += du_log_dirichlet(self.gamma_star[s, :], self.alpha)
        for a in range(4):
            cMLogL += du_log_dirichlet(self.eta_star[a, :], self.delta)
        cMLogL += self.V * np.log(1.0 / float(self.nTauStates))
        storeLogTau = np.array([self.logTauProb(self.gamma_store[i], self.eta_store[i]) for i in range(self.tau_comp_iter)])
        logTauHat = self.logMean(storeLogTau)
        eng = self._engine()
        storeLogGamma = np.zeros(self.max_iter)
        for i in range(self.max_iter):                                          # pi term (:664-681)
            sum_mu, _ = self.sampleMu(self.tau_star, self.gamma, self.eta)
            self.sampleGamma()
            self.sampleEta()
            eng.set_option("advance_sweep", 1)
            storeLogGamma[i] = self._log_gamma_post(sum_mu)
        logGammaHat = self.logMean(storeLogGamma)
        storeLogEpsilon = np.zeros(self.max_iter)
        for i in range(self.max_iter):                                          # epsilon term (:690-703)
            _, sum_E = self.sampleMu(self.tau_star, self.gamma_star, self.eta)
            self._eta_draw = None
            self.sampleEta()
            eng.set_option("advance_sweep", 1)
            storeLogEpsilon[i] = self._log_eta_post(sum_E)
        logEpsilonHat = self.logMean(storeLogEpsilon)
        _sampletau.advance_global_sweep(eng.get_rng()[0])
        print(str(cMLogL) + "," + str(logGammaHat) + "," + str(logEpsilonHat) + "," + str(logTauHat))
        return cMLogL - logGammaHat - logEpsilonHat - logTauHat

    def chibMarginalLogLikelihood2(self):
        """Chib's estimator with the strain-by-strain tau term (:538-608); the eta term reads the per-sweep Esum of the
        last update() (E_store[i].sum(axis=(0,1)), :557)."""
        if getattr(self, "_Esum_store", None) is None or self._Esum_store.shape[0] < self.max_iter:
            raise ValueError("chibMarginalLogLikelihood2 needs the E statistics of a preceding update()")
        cMLogL = self.logLikelihood(self.gamma_star, self.tau_star, self.eta_star)
        logGammaPrior = sum(du_log_dirichlet(self.gamma_star[s, :], self.alpha) for s in range(self.S))
        logEtaPrior = sum(du_log_dirichlet(self.eta_star[a, :], self.delta) for a in range(4))
        logTauPrior = self.V * self.G * np.log(1.0 / 4.0)
        storeLogEpsilon = np.array([self._log_eta_post(self._Esum_store[i]) for i in range(self.max_iter)])
        logEpsilonHat = self.logMean(storeLogEpsilon)
        storeLogGamma = np.zeros(self.max_iter)
        for i in range(self.max_iter):                                          # (:566-583)
            self.sampleTau(self.gamma, self.eta_star)
            sum_mu, _ = self.sampleMu(self.tau, self.gamma, self.eta_star)
            self.sampleGamma()
            self._eta_draw = None
            print(str(i) + ",GC," + str(self._log_gamma_post(sum_mu)))
            storeLogGamma[i] = self._log_gamma_post(sum_mu)
        logGammaHat = self.logMean(storeLogGamma)
        logTauHat = 0.0
        star_ix = self._tau_index(star=True)
        for h in range(self.G):                                                 # (:587-603)
            workingTau = np.copy(self.tau_star)
            storeLogTau = np.zeros(self.max_iter)
            for i in range(self.max_iter):
                tauLogProb = self.sampleTauFixTau(workingTau, h, self.gamma_star, self.eta_star)
                temp = float(tauLogProb[np.arange(self.V), star_ix[:, h]].sum())
                storeLogTau[i] = temp
                print(str(i) + ",GT," + str(h) + "," + str(temp))
            logTauHat += self.logMean(storeLogTau)
        return cMLogL + logEtaPrior - logEpsilonHat + logGammaPrior - logGammaHat + logTauPrior - logTauHat
