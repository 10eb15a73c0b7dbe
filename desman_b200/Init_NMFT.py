"""Mirror of the reference class desman/Init_NMFT.py: same constructor, attributes (tau [4V,G],
gamma [G,S], both assignable from outside -- bin/desman:188, Eta_Sampler.py:132) and methods; the
multiplicative-update loop runs in the batched CUDA kernels of libdesman_b200.so.

The random initial factors are drawn on the host from the caller's RandomState in the reference's
order (Init_NMFT.py:66-86), so the initialisation is stream-compatible with the reference.
"""
import logging

import numpy as np

from .engine import Engine


class Init_NMFT:
    """Initialises tau and gamma based on tensor non-negative matrix factorization"""

    BASE_PRIOR = 1.0

    def __init__(self, snps, rank, randomState, n_run=1, max_iter=5000, min_change=1.0e-5, alpha_constant=0.01,
                 device=0):
        self.V = snps.shape[0]
        self.S = snps.shape[1]
        self.G = rank
        self.randomState = randomState
        self.n_run = n_run
        self.max_iter = max_iter
        self.min_change = min_change
        self.alpha = np.empty(self.G); self.alpha.fill(alpha_constant)
        self.alpha4 = np.empty(4); self.alpha4.fill(alpha_constant)
        self.snps = np.ascontiguousarray(snps, dtype=np.int64)
        self.N = self.V * 4
        self.tau = np.zeros((self.N, self.G))
        self.gamma = np.zeros((self.G, self.S))
        self.div = None
        self.n_iter = 0
        self.div_trace = None
        self._device = device
        self._freq = None

    @property
    def freq_matrix(self):
        """(snps + 1) / sum_b (snps + 1), rows v + a*V (Init_NMFT.py:49-60); built lazily on the host for inspection."""
        if self._freq is None:
            x = self.snps.astype(np.float64) + self.BASE_PRIOR
            f = x / x.sum(axis=2)[:, :, np.newaxis]
            self._freq = np.ascontiguousarray(np.transpose(f, (2, 0, 1)).reshape(self.N, self.S))
        return self._freq

    # ------------------------------------------------------------------ random starts (host RNG, reference order)
    def random_initialize(self):
        if self.G > 1:
            temp = self.randomState.dirichlet(self.alpha, size=self.S)                 # :69
        else:
            temp = np.ones((self.S, self.G))
        self.gamma = np.transpose(temp)
        self.random_initialize_tau()

    def random_initialize_tau(self):
        # V*G successive dirichlet(alpha4) draws in (v, g) order (:74-78) == one bulk draw of that many rows
        d = self.randomState.dirichlet(self.alpha4, size=self.V * self.G).reshape(self.V, self.G, 4)
        self.tau = np.ascontiguousarray(np.transpose(d, (2, 0, 1)).reshape(self.N, self.G))

    def _adjustment(self):
        self.tau = np.maximum(self.tau, np.finfo(self.tau.dtype).eps)
        self.gamma = np.maximum(self.gamma, np.finfo(self.gamma.dtype).eps)

    # ------------------------------------------------------------------ device loops
    def _run(self, fix_gamma):
        eng = Engine(self._device, seed=0)
        try:
            tau, gamma, it, div, trace = eng.nmft_factorize(self.snps, self.tau, np.ascontiguousarray(self.gamma),
                                                            max_iter=self.max_iter, min_change=self.min_change,
                                                            fix_gamma=fix_gamma, want_trace=True)
        finally:
            eng.close()
        self.tau, self.gamma, self.n_iter, self.div, self.div_trace = tau, gamma, it, div, trace
        if fix_gamma != 2:
            for i in range(0, it, 100):                                               # :112-113 / :146-147
                logging.info('NTF Iter %d, div = %f' % (i, trace[i]))

    def factorize(self):
        """Init_NMFT.py:98-115"""
        for run in range(self.n_run):
            self.random_initialize()
            self._run(False)

    def factorize_tau(self):
        """Init_NMFT.py:134-149: gamma stays fixed (set by the caller), no eps clamp."""
        for run in range(self.n_run):
            self.random_initialize_tau()
            self._run(True)

    def factorize_gamma(self):
        """Init_NMFT.py:117-132: tau stays fixed (set by the caller), no random start, no eps clamp."""
        for run in range(self.n_run):
            self._run(2)
            for i in range(0, self.n_iter, 100):                                      # :129-130
                print(str(i) + "," + str(self.div_trace[i]))

    # ------------------------------------------------------------------ single steps (the bodies of the reference's loops)
    def _adjustment_input(self, X):
        return np.maximum(X, np.finfo(self.tau.dtype).eps)                            # :93-97

    def _step(self, mode, max_iter):
        eng = Engine(self._device, seed=0)
        try:
            tau, gamma, it, div, _ = eng.nmft_factorize(self.snps, self.tau, np.ascontiguousarray(self.gamma), max_iter=max_iter,
                                                        min_change=-1.0, fix_gamma=mode)
        finally:
            eng.close()
        return tau, gamma, div

    def div_objective(self):
        """KL divergence of X from tau gamma with the factors as they are (:152-156)."""
        return self._step(1, 0)[2]

    def div_update(self):
        """One multiplicative update of gamma, then tau (:158-181) -- followed by the eps clamp of _adjustment(), which the
        reference's factorize() applies right after every div_update() (:107-108)."""
        self.tau, self.gamma, self.div = self._step(0, 1)

    def div_update_tau(self):
        """One multiplicative update of tau with gamma fixed (:192-205)."""
        self.tau, self.gamma, self.div = self._step(1, 1)

    def div_update_gamma(self):
        """One multiplicative update of gamma with tau fixed (:183-190)."""
        self.tau, self.gamma, self.div = self._step(2, 1)

    # ------------------------------------------------------------------ results
    def get_gamma(self):
        return np.transpose(self.gamma)                                                # :209-210

    def get_tau(self):
        """One-hot argmax over the 4 bases per (v,g); strict '>' from 0.0, ties -> lowest base (:230-245)."""
        t = self.tau.reshape(4, self.V, self.G)
        best = np.zeros((self.V, self.G), dtype=np.int64)
        maxt = np.zeros((self.V, self.G))
        for a in range(4):
            upd = t[a] > maxt
            best[upd] = a
            maxt[upd] = t[a][upd]
        ret = np.zeros((self.V, self.G, 4), dtype=np.int64)
        np.put_along_axis(ret, best[:, :, None], 1, axis=2)
        return ret

    def discretise_tau(self):
        d = self.get_tau()
        self.tau = np.ascontiguousarray(np.transpose(d, (2, 0, 1)).reshape(self.N, self.G)).astype(np.float64)
