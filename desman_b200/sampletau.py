"""Drop-in for the reference's `sampletau` extension module (sampletau/sampletau.pyx:21-57).

Same four functions, same argument checks, same in-place mutation of `tau`, same GSL-compatible
MT19937 stream -- but `sample_tau` runs the sm_100a kernel of libdesman_b200.so.  Callers in the
reference: bin/desman:131-132,242; HaploSNP_Sampler.py:345,392; Eta_Sampler.py:367.

Argument errors mirror what Cython's buffer acquisition raises for the reference module
(SURVEY.md section 8b): TypeError for None / non-arrays, ValueError for wrong dtype, rank or
memory order, OverflowError for a seed that does not fit a C int.
"""
import ctypes as C

import numpy as np

from . import _lib


_seed = 0          # last seed given to setRNG: the device chains are keyed by it, like the GSL stream
_sweep = 0         # process-global Philox sweep counter (the analogue of the single global GSL stream)


def current_seed():
    return _seed


def global_sweep():
    return _sweep


def advance_global_sweep(sweep):
    global _sweep
    _sweep = max(_sweep, int(sweep))


def initRNG():
    """c_initRNG (c_sample_tau.c:26-34)"""
    _lib.lib().c_initRNG()
    msg = _lib.last_error()
    if msg and "c_initRNG" in msg:
        raise _lib.DesmanB200Error(msg)


def setRNG(seed):
    """c_setRNG (c_sample_tau.c:36-40); `seed` is a C int in the reference signature (sampletau.pyx:28)."""
    seed = int(seed)
    if seed > 2**31 - 1:
        raise OverflowError("value too large to convert to int")
    if seed < -2**31:
        raise OverflowError("value too small to convert to int")
    global _seed, _sweep
    _seed, _sweep = seed & 0xFFFFFFFFFFFFFFFF, 0
    _lib.lib().c_setRNG(C.c_ulong(seed & 0xFFFFFFFFFFFFFFFF))  # int -> unsigned long wrap, as in C


def freeRNG():
    """c_freeRNG (c_sample_tau.c:42-45)"""
    _lib.lib().c_freeRNG()


def _check(arr, name, ndim, dtype, cname):
    if arr is None or not isinstance(arr, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(arr).__name__))
    if arr.ndim != ndim:
        raise ValueError("Buffer has wrong number of dimensions (expected %d, got %d)" % (ndim, arr.ndim))
    if arr.dtype != dtype:
        got = {"int32": "int", "float32": "float", "int64": "long", "float64": "double"}.get(arr.dtype.name,
                                                                                             arr.dtype.name)
        raise ValueError("Buffer dtype mismatch, expected '%s' but got '%s'" % (cname, got))
    if not arr.flags.c_contiguous:
        raise ValueError("ndarray is not C-contiguous")


def sample_tau(tau, pi, eta, variants):
    """sample_tau(tau, pi, eta, variants) -> nchange   (sampletau.pyx:38-57)

    tau      int64  [V,G,4] one-hot, mutated in place
    pi       float64 [S,G]  strain abundances gamma
    eta      float64 [4,4]  error matrix, row = true base
    variants int64  [V,S,4] base counts
    """
    _check(tau, "tau", 3, np.int64, "long")
    _check(pi, "pi", 2, np.float64, "double")
    _check(eta, "eta", 2, np.float64, "double")
    _check(variants, "variants", 3, np.int64, "long")
    nV, nG, nS = tau.shape[0], tau.shape[1], pi.shape[0]      # sampletau.pyx:51-53
    n = _lib.lib().c_sample_tau(tau.ctypes.data_as(_lib._p64), pi.ctypes.data_as(_lib._pd),
                                eta.ctypes.data_as(_lib._pd), variants.ctypes.data_as(_lib._p64), nV, nG, nS)
    if n < 0:
        raise _lib.DesmanB200Error("c_sample_tau failed: " + _lib.last_error())
    return n


class Batch:
    """Many small sample_tau problems per launch (an addition to the reference module): what Eta_Sampler.sampleTauC does gene by
    gene (Eta_Sampler.py:355-369, :430-446), with the counts of all genes resident on the device.

        b = sampletau.Batch([variants_gene0, variants_gene1, ...])        # int64 [V_k,S,4] each, uploaded once
        nchange = b.sample_tau([tau_gene0, ...], [gammaR_gene0, ...], epsilon)   # one launch; tau arrays mutated in place

    Draw for draw (the process-global MT19937 stream of setRNG, gene after gene) the result equals
    [sample_tau(tau_k, pi_k, eta, variants_k) for k in range(len(variants))]."""

    def __init__(self, variants):
        self._var = []
        for v in variants:
            _check(v, "variants", 3, np.int64, "long")
            self._var.append(v)
        if not self._var:
            raise ValueError("Batch needs at least one problem")
        self.S = self._var[0].shape[1]
        if any(v.shape[1] != self.S or v.shape[2] != 4 for v in self._var):
            raise ValueError("every problem must be [V_k,%d,4]" % self.S)
        n = len(self._var)
        ptrs = (_lib._p64 * n)(*[v.ctypes.data_as(_lib._p64) for v in self._var])
        nV = (C.c_int * n)(*[v.shape[0] for v in self._var])
        self._h = C.c_void_p()
        if _lib.lib().desman_batch_create(C.byref(self._h), n, ptrs, nV, self.S) != 0:
            raise _lib.DesmanB200Error("desman_batch_create failed: " + _lib.last_error())

    def sample_tau(self, taus, pis, eta):
        n = len(self._var)
        if len(taus) != n or len(pis) != n:
            raise ValueError("one tau and one pi per problem")
        _check(eta, "eta", 2, np.float64, "double")
        G = pis[0].shape[1]
        for t, pi, v in zip(taus, pis, self._var):
            _check(t, "tau", 3, np.int64, "long")
            _check(pi, "pi", 2, np.float64, "double")
            if t.shape != (v.shape[0], G, 4) or pi.shape != (self.S, G):
                raise ValueError("tau must be [V_k,%d,4] and pi [%d,%d] for every problem" % (G, self.S, G))
        tp = (_lib._p64 * n)(*[t.ctypes.data_as(_lib._p64) for t in taus])
        pp = (_lib._pd * n)(*[x.ctypes.data_as(_lib._pd) for x in pis])
        out = (C.c_int * n)()
        if _lib.lib().desman_batch_sample_tau(self._h, tp, pp, eta.ctypes.data_as(_lib._pd), G, out) != 0:
            raise _lib.DesmanB200Error("desman_batch_sample_tau failed: " + _lib.last_error())
        return list(out)

    def close(self):
        if self._h:
            _lib.lib().desman_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
