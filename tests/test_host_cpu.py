"""CPU tests of the host logic and of the C-ABI surface (no compute calls: no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_is_built_and_exports_every_declared_symbol():
    from desman_b200 import _lib, build
    build.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    header = open(os.path.join(ROOT, "include", "desman_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b((?:c_|desman_)\w+)\s*\(", header))
    declared.discard("desman_ctx")
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in _lib.lib().desman_build_info()


def test_reference_abi_names_are_exact():
    # sampletau.pyx:13-19 binds exactly these four C symbols
    from desman_b200 import _lib
    for name in ("c_sample_tau", "c_initRNG", "c_setRNG", "c_freeRNG"):
        assert name in _lib.SYMBOLS


def test_sampletau_argument_checks_mirror_cython():
    from desman_b200 import sampletau
    tau = np.zeros((3, 2, 4), dtype=np.int64)
    pi = np.full((5, 2), 0.5)
    eta = np.full((4, 4), 0.25)
    var = np.zeros((3, 5, 4), dtype=np.int64)
    with pytest.raises(TypeError, match="Argument 'tau' has incorrect type"):
        sampletau.sample_tau(None, pi, eta, var)
    with pytest.raises(TypeError):
        sampletau.sample_tau(tau.tolist(), pi, eta, var)
    with pytest.raises(ValueError, match="expected 'long' but got 'int'"):
        sampletau.sample_tau(tau.astype(np.int32), pi, eta, var)
    with pytest.raises(ValueError, match="expected 'double'"):
        sampletau.sample_tau(tau, pi.astype(np.float32), eta, var)
    with pytest.raises(ValueError, match="not C-contiguous"):
        sampletau.sample_tau(np.asfortranarray(tau), pi, eta, var)
    with pytest.raises(ValueError, match="wrong number of dimensions \\(expected 2, got 1\\)"):
        sampletau.sample_tau(tau, pi[0], eta, var)
    with pytest.raises(OverflowError):
        sampletau.setRNG(2**31)


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly, not compute on the host."""
    from desman_b200 import _lib, engine
    try:
        n = _lib.device_count()
    except _lib.DesmanB200Error:
        n = 0
    if n > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.DesmanB200Error):
        engine.Engine(0, seed=1)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under desman_b200/ may import, include, link or load it
    (comments may cite the oracle function a kernel mirrors)."""
    pkg = os.path.join(ROOT, "desman_b200")
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(#\s*include\s*[<\"][^>\"]*oracle)|(liboracle)|(dlopen\([^)]*oracle)|(CDLL\([^)]*oracle)",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), f


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one line on stdout, valid JSON, our arm's
    metric/unit/config names, impl = reference, a cpu_baseline describing the run, zero-copy e2e mirror of the value."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["higher_is_better"] is True and d["steps"] == 2
    assert d["metric"].startswith("Gibbs sweeps/sec") and d["config"]["workload"].startswith("BASELINE config C3")
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    # "reference-tau+port": tau by the reference's own c_sample_tau.c (oracle/_ref), the numpy steps by their C port
    assert d["cpu_baseline"]["kind"] in ("reference-tau+port", "port")
    import bench
    assert d["config"]["workload"] == bench.workload_name("c3", 100000, 64, 8)          # the string our arm prints too
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_batch_and_nmft_host_checks_need_no_device():
    """Argument checks of the additions to the reference surfaces happen on the host, before any device call."""
    import numpy as np
    from numpy.random import RandomState

    from desman_b200 import sampletau
    from desman_b200.Init_NMFT import Init_NMFT
    with pytest.raises(ValueError, match="at least one problem"):
        sampletau.Batch([])
    a, b = np.ones((3, 5, 4), dtype=np.int64), np.ones((2, 6, 4), dtype=np.int64)
    with pytest.raises(ValueError, match="every problem must be"):
        sampletau.Batch([a, b])
    with pytest.raises(ValueError, match="dtype mismatch"):
        sampletau.Batch([a.astype(np.int32)])
    n = Init_NMFT(a, 2, RandomState(1))
    x = np.array([[0.0, 1.0], [1e-300, 0.5]])
    y = n._adjustment_input(x)                                   # Init_NMFT.py:93-97: floor at the machine epsilon, input untouched
    assert np.array_equal(y, np.maximum(x, np.finfo(np.float64).eps)) and x[0, 0] == 0.0
    n.random_initialize()                                        # host RNG in the reference's order: shapes of the factors
    assert n.tau.shape == (12, 2) and n.gamma.shape == (2, 5) and np.allclose(n.gamma.sum(0), 1.0)


def test_class_host_helpers_match_the_reference_class():
    """The small numpy methods of the class mirror against golden vectors of the UNMODIFIED reference class
    (tests/golden/make_golden.py helpers); the constructor and these methods need no device."""
    import numpy as np
    from numpy.random import RandomState

    from conftest import golden, onehot
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    z = golden("host_helpers_kat.npz")
    tau = onehot(z["tau"])
    V, G = tau.shape[0], tau.shape[1]
    S = z["gamma"].shape[0]
    hs = HaploSNP_Sampler(np.ones((V, S, 4), dtype=np.int64), G, RandomState(1))
    hs.tau = tau
    hs.updateTauIndices()
    assert np.array_equal(np.asarray(hs.tauIndices), z["tauIndices"])            # :224-231: base-4 code, strain 0 most significant
    for x, y in zip(z["lv"], z["lv_out"]):
        assert np.allclose(hs.normaliseLogProb(x), y, rtol=1e-12, atol=1e-12)    # :186-194
    for k in range(3):
        assert np.isclose(hs.logMean(z["ls%d" % k]), z["ls_out"][k], rtol=1e-12)  # :526-540
    assert [hs.tauDist(tau[i], tau[i + 1]) for i in range(V - 1)] == z["dist"].tolist()   # :137-147
    for i in range(V):
        assert np.allclose(hs.baseProbabilityGivenTau(tau[i], z["gamma"], z["eta"]), z["base_prob"][i], rtol=1e-12)   # :129-134


def test_snd_and_remove_degenerate_match_the_reference_class():
    """calculateSND / compSND / variableTau / removeDegenerate of the mirror against golden vectors of the UNMODIFIED class
    (tests/golden/make_golden.py degenerate): two pairs of identical haplotypes are merged and their gamma columns added."""
    import numpy as np
    from numpy.random import RandomState

    from conftest import golden, onehot
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    z = golden("degenerate_kat.npz")
    tau, other = onehot(z["tau"]), onehot(z["other"])
    V, G = tau.shape[0], tau.shape[1]
    S = z["gamma"].shape[0]
    hs = HaploSNP_Sampler(np.ones((V, S, 4), dtype=np.int64), G, RandomState(1), max_iter=2)
    hs.tau, hs.gamma = tau.copy(), z["gamma"].copy()
    assert np.array_equal(hs.calculateSND(hs.tau), z["snd"])                    # :712-730
    assert np.array_equal(hs.compSND(hs.tau, other), z["comp"])                 # :747-770
    assert np.array_equal(hs.variableTau(hs.tau), z["variable"])                # :732-745
    hs.removeDegenerate()                                                       # :771-832
    assert hs.G == int(z["G_after"]) == 3
    assert np.array_equal(np.argmax(hs.tau, 2), z["tau_after"])
    assert np.allclose(hs.gamma, z["gamma_after"], rtol=0, atol=0)
    assert np.array_equal(np.asarray(hs.tauIndices), z["tauIndices_after"])
    assert np.array_equal(hs.alpha, z["alpha_after"]) and tuple(hs.gamma_store.shape) == tuple(z["gamma_store_shape"])
