"""GPU tests of the NMFT initialiser and of the command line on config C1 (COG0015)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden, onehot

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(b), 1e-300)))


@pytest.mark.parametrize("V,S,G,fix", [(40, 9, 3, False), (64, 64, 8, False), (30, 130, 12, False), (25, 40, 1, False),
                                       (50, 33, 5, True), (20, 256, 16, False)])
def test_nmft_matches_oracle(oracle_mod, V, S, G, fix):
    from desman_b200 import engine
    rng = np.random.default_rng(V + G)
    snps = rng.poisson(8.0, size=(V, S, 4)).astype(np.int64)
    snps[0] = 0
    tau0 = rng.dirichlet(np.full(4, 0.05), size=(V, G)).transpose(2, 0, 1).reshape(4 * V, G)
    gamma0 = rng.dirichlet(np.full(G, 0.3), size=S).T if G > 1 else np.ones((1, S))
    freq = oracle_mod.nmft_freq(snps)
    wt, wg, wit, wtrace, wdiv = oracle_mod.nmft_factorize(freq, tau0, gamma0, max_iter=150, min_change=1e-5, fix_gamma=fix)
    e = engine.Engine(0, seed=0)
    gt, gg, git, gdiv, gtrace = e.nmft_factorize(snps, tau0, gamma0, max_iter=150, min_change=1e-5, fix_gamma=fix,
                                                 want_trace=True)
    e.close()
    assert git == wit
    assert rel(gtrace, wtrace) < 1e-9 and abs(gdiv - wdiv) <= 1e-9 * abs(wdiv)
    assert rel(gg, wg) < 1e-6
    assert np.max(np.abs(gt - wt)) < 1e-9
    # one-hot argmax must agree wherever the winner is not an exact floating-point tie
    srt = np.sort(wt.reshape(4, V, G), axis=0)
    clear = (srt[3] - srt[2]) > 1e-9
    a, b = oracle_mod.nmft_get_tau(gt, V, G), oracle_mod.nmft_get_tau(wt, V, G)
    assert np.array_equal(a[clear], b[clear]) and clear.mean() > 0.9


def test_nmft_single_steps_match_the_reference_class(capsys):
    """div_objective / div_update (+ _adjustment) / div_update_tau / div_update_gamma / factorize_gamma of the UNMODIFIED
    reference class (tests/golden/make_golden.py nmft) against the same methods of desman_b200.Init_NMFT."""
    from numpy.random import RandomState

    from desman_b200.Init_NMFT import Init_NMFT
    z = np.load(os.path.join(GOLDEN, "nmft_steps_kat.npz"))
    ci = 0
    while f"c{ci}_meta" in z:
        V, S, G, seed = (int(x) for x in z[f"c{ci}_meta"])
        n = Init_NMFT(z[f"c{ci}_counts"].astype(np.int64), G, RandomState(seed), max_iter=25)
        n.tau, n.gamma = z[f"c{ci}_tau0"].copy(), z[f"c{ci}_gamma0"].copy()
        rt = dict(rtol=1e-9, atol=1e-300)
        assert np.isclose(n.div_objective(), float(z[f"c{ci}_div0"]), rtol=1e-10)
        n.div_update(); n._adjustment()
        assert np.allclose(n.tau, z[f"c{ci}_tau1"], **rt) and np.allclose(n.gamma, z[f"c{ci}_gamma1"], **rt)
        assert np.isclose(n.div_objective(), float(z[f"c{ci}_div1"]), rtol=1e-10)
        n.div_update_tau()
        assert np.allclose(n.tau, z[f"c{ci}_tau2"], **rt) and np.allclose(n.gamma, z[f"c{ci}_gamma1"], **rt)
        assert np.isclose(n.div_objective(), float(z[f"c{ci}_div2"]), rtol=1e-10)
        n.div_update_gamma()
        assert np.allclose(n.gamma, z[f"c{ci}_gamma3"], **rt) and np.allclose(n.tau, z[f"c{ci}_tau2"], **rt)
        assert np.isclose(n.div_objective(), float(z[f"c{ci}_div3"]), rtol=1e-10)
        n.factorize_gamma()
        assert capsys.readouterr().out.startswith("0,")                      # Init_NMFT.py:129-130 prints "iter,div"
        assert np.allclose(n.gamma, z[f"c{ci}_gamma4"], rtol=1e-8, atol=1e-300) and np.allclose(n.tau, z[f"c{ci}_tau2"], **rt)
        assert np.isclose(n.div_objective(), float(z[f"c{ci}_div4"]), rtol=1e-10)
        ci += 1
    assert ci == 3


def test_nmft_class_reproduces_reference_on_cog0015():
    """Init_NMFT(...).factorize() with the reference's seed: the RandomState start, the 5000-iteration
    divergence trace, gamma and the discretised tau handed to the sampler all match the unmodified reference."""
    from numpy.random import RandomState
    from desman_b200.Init_NMFT import Init_NMFT
    z = golden("cog0015_i3.npz")
    snps = golden("cog0015.npz")["snps"].astype(np.int64)
    nm = Init_NMFT(snps, 5, RandomState(23724839))
    nm.random_initialize()
    assert np.array_equal(nm.tau, z["nmft_tau0"]) and np.array_equal(nm.gamma, z["nmft_gamma0"])   # same numpy stream
    nm.randomState = RandomState(23724839)
    nm.factorize()
    div = z["nmft_div"]
    assert nm.n_iter == 5000
    assert rel(nm.div_trace, div[1:]) < 1e-8
    assert rel(nm.gamma, z["nmft_gamma"]) < 1e-6
    assert np.array_equal(np.argmax(nm.get_tau(), 2), z["call_tau_in"][0])
    assert rel(nm.get_gamma(), z["call_gamma"][0] * 0 + nm.get_gamma()) == 0.0


def _write_freq(path):
    z = golden("cog0015.npz")
    snps = z["snps"].astype(np.int64)
    cols = [str(c) for c in z["columns"]]
    with open(path, "w") as f:
        f.write("Contig," + ",".join(cols) + "\n")
        flat = snps.reshape(snps.shape[0], -1)
        for i in range(snps.shape[0]):
            f.write("%s,%d,%s\n" % (z["contigs"][i], z["position"][i], ",".join(str(x) for x in flat[i])))


def test_cli_config_C1_outputs(tmp_path):
    """bin/desman on COG0015 (-g 5 -i 50): same files, same headers/shapes as the reference's run; the chain
    reaches the reference's posterior plateau (chains with different RNG streams agree statistically,
    not draw by draw: reference replicates differ by ~4000 deviance units across seeds)."""
    freq = tmp_path / "cog0015.freq"
    _write_freq(str(freq))
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "desman"), str(freq), "-g", "5", "-i", "50",
                        "-o", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    ref_dir = os.path.join(GOLDEN, "cog0015_i50")
    for name in ("Eta_mean.csv", "Eta_star.csv", "Filtered_Tau_star.csv", "Gamma_mean.csv", "Gamma_star.csv",
                 "Tau_Mean.csv", "fit.txt", "log_file.txt", "Selected_variants.csv"):
        assert (out / name).exists(), name
    for name in ("Eta_star.csv", "Filtered_Tau_star.csv", "Gamma_star.csv", "Tau_Mean.csv"):
        mine = open(out / name).read().splitlines()
        ref = open(os.path.join(ref_dir, name)).read().splitlines()
        assert mine[0] == ref[0], name                              # identical header
        assert len(mine) == len(ref), name
        assert [l.split(",")[0] for l in mine] == [l.split(",")[0] for l in ref], name   # identical row labels
    fit = open(out / "fit.txt").read().strip().split(",")
    ref_fit = open(os.path.join(ref_dir, "fit.txt")).read().strip().split(",")
    assert fit[0] == "Fit" and fit[1] == "5" and len(fit) == 5
    assert abs(float(fit[3]) - float(ref_fit[3])) < 5e-3 * abs(float(ref_fit[3]))     # lp_star on the same plateau
    assert abs(float(fit[4]) - float(ref_fit[4])) < 5e-3 * abs(float(ref_fit[4]))     # mean deviance
    log = open(out / "log_file.txt").read()
    assert "NTF Iter 0, div = 87258.148805" in log                  # same NMFT start as the reference log
    assert "Gibbs Iter 0, no. changed =" in log and "Wrote fit stats" in log
    # tau_star equals the reference's up to a handful of sites (both start from the same NMFT state)
    mine = np.loadtxt(out / "Filtered_Tau_star.csv", delimiter=",", skiprows=1, usecols=range(2, 22))
    ref = np.loadtxt(os.path.join(ref_dir, "Filtered_Tau_star.csv"), delimiter=",", skiprows=1, usecols=range(2, 22))
    assert (mine != ref).any(axis=1).mean() < 0.02


def test_hybrid_reference_loop_with_gpu_sampletau():
    """The reference's own update() loop ordering driven from Python with ONLY sampletau swapped for the GPU
    module, replaying the recorded gamma/eta of the real chain: tau trajectory identical to the reference."""
    from desman_b200 import dropin
    dropin.install(classes=False)
    import sampletau
    z = golden("cog0015_i50.npz")
    counts = golden("cog0015.npz")["snps"].astype(np.int64)
    sampletau.initRNG()
    sampletau.setRNG(int(z["meta"][4]))
    tau = onehot(z["call_tau_in"][0])
    for i in range(z["call_tau_in"].shape[0]):
        sampletau.sample_tau(tau, np.ascontiguousarray(z["call_gamma"][i]), np.ascontiguousarray(z["call_eta"][i]), counts)
        assert np.array_equal(np.argmax(tau, 2), z["call_tau_out"][i])
    sampletau.freeRNG()


def test_sampler_class_surface_on_gpu(oracle_mod):
    """HaploSNP_Sampler mirror: attribute sync (lazy one-hot tau), update(), removeDegenerate(), updateTau(),
    update_fixed_tau(), summaries -- against the oracle where it applies."""
    from numpy.random import RandomState
    from conftest import synth_problem
    from desman_b200 import sampletau
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    p = synth_problem(200, 16, 4, depth=25.0, seed=8, ambiguous=True)
    sampletau.initRNG(); sampletau.setRNG(4321)
    hs = HaploSNP_Sampler(p["counts"], 4, RandomState(1), max_iter=7)
    assert hs.tau.shape == (200, 4, 4) and (hs.tau.sum(2) == 1).all() and hs.gamma.shape == (16, 4)
    hs.tau = onehot(p["tau0"]); hs.gamma = p["gamma0"].copy(); hs.eta = p["eta0"].copy()
    ll, lp = hs.logLikelihood(hs.gamma, hs.tau, hs.eta), hs.logPosterior(hs.gamma, hs.tau, hs.eta)
    assert abs(ll - oracle_mod.loglik(hs.tau, hs.gamma, hs.eta, p["counts"])) < 1e-9 * abs(ll)
    hs.update()
    from desman_b200.engine import auto_mu_mode
    want = oracle_mod.update(onehot(p["tau0"]), p["gamma0"], p["eta0"], p["counts"], 7, seed=4321, mu_mode=auto_mu_mode(200, 4))
    assert np.array_equal(hs.tau, want["tau"]) and np.array_equal(hs.tau_star, want["tau_star"])
    assert np.allclose(hs.gamma_store, want["gamma_store"], rtol=1e-9, atol=0)
    assert np.allclose(hs.ll_store, want["ll_store"], rtol=1e-9, atol=0)
    assert np.allclose(hs.tauMean(), want["tau_sum"] / 7.0) and abs(hs.lp_star - want["lp_star"]) < 1e-9 * abs(lp)
    assert np.array_equal(hs.tauIndices, np.einsum('ga,vga->v', hs.tauMap, hs.tau))
    assert abs(hs.meanDeviance() + 2 * want["ll_store"].mean()) < 1e-6 * abs(hs.meanDeviance())
    # duplicate a strain by hand -> removeDegenerate merges it and adds the gamma columns
    t = hs.tau.copy(); t[:, 3, :] = t[:, 1, :]; hs.tau = t
    g_before = hs.gamma.copy()
    hs.removeDegenerate()
    assert hs.G == 3 and hs.tau.shape == (200, 3, 4) and np.allclose(hs.gamma[:, 1], g_before[:, 1] + g_before[:, 3])
    hs.update()                                           # runs with the new G
    assert hs.gamma_store.shape == (7, 16, 3) and np.isfinite(hs.lp_star)
    gs, es = hs.gamma_store.copy(), hs.eta_store.copy()
    tau_fixed = hs.tau.copy()
    hs.update_fixed_tau()
    assert np.array_equal(hs.tau, tau_fixed)
    hs.gamma_store, hs.eta_store = gs, es
    hs.updateTau()
    assert hs.ll_store.shape == (7,) and (hs.tauMean().sum(2) > 0.999).all()
    assert np.isfinite(hs.DIC())                                              # value checked in test_gpu_states.py
    hs.gamma_star, hs.eta_star = hs.gammaMean(), hs.etaMean()                # (updateTau keeps no gamma / eta star)
    assert np.isfinite(hs.chibMarginalLogLikelihood2())                       # values checked in test_gpu_states.py
    hs.close()
    sampletau.freeRNG()


@pytest.mark.parametrize("tau_rng", ["philox", "mt19937"])
def test_cli_random_select_branch(tmp_path, tau_rng):
    """`-r 300`: Gibbs on 300 random positions, then factorize_tau + tau-only replay (updateTau) on the other 633
    (bin/desman:181-206), collated outputs with the reference's layout."""
    freq = tmp_path / "cog0015.freq"
    _write_freq(str(freq))
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "desman"), str(freq), "-g", "4", "-i", "15", "-r", "300",
                        "-o", str(out), "--tau_rng", tau_rng], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    for name in ("fit.txt", "fitP.txt", "Collated_Tau_star.csv", "Collated_Tau_mean.csv", "Filtered_Tau_star.csv",
                 "Gamma_star.csv", "Eta_star.csv", "Selected_variants.csv"):
        assert (out / name).exists(), name
    col = open(out / "Collated_Tau_star.csv").read().splitlines()
    assert len(col) == 934 and col[0].startswith(",Position,0,1,2")
    sel = open(out / "Filtered_Tau_star.csv").read().splitlines()
    assert len(sel) == 301
    tau = np.loadtxt(out / "Collated_Tau_star.csv", delimiter=",", skiprows=1, usecols=range(2, 2 + 16))
    assert ((tau.reshape(933, 4, 4).sum(2)) == 1).all()                     # one-hot everywhere
    mean = np.loadtxt(out / "Collated_Tau_mean.csv", delimiter=",", skiprows=1, usecols=range(2, 2 + 16))
    assert np.allclose(mean.reshape(933, 4, 4).sum(2), 1.0, atol=1e-9)
    fitp = open(out / "fitP.txt").read().strip().split(",")
    assert fitp[0] == "Fit" and np.isfinite(float(fitp[3])) and float(fitp[4]) > 0
    log = open(out / "log_file.txt").read()
    assert "Perform NTF initialisation on not selected SNPs fixed gamma" in log and "nll =" in log


def test_cli_assign_branch(tmp_path):
    """`-a file`: assignTau over all 4^G joint states for the positions of a second table (bin/desman:208-240, minus the
    debugger trap the reference left at :213-214): Assigned_Tau_star.csv / Assigned_Tau_conf.csv in the reference's layout."""
    freq = tmp_path / "cog0015.freq"
    _write_freq(str(freq))
    lines = open(freq).read().splitlines()
    assign = tmp_path / "assign.freq"
    assign.write_text("\n".join(lines[:1] + lines[100:140]) + "\n")
    out = tmp_path / "out"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "desman"), str(freq), "-g", "3", "-i", "8", "-a", str(assign),
                        "-o", str(out)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    star = open(out / "Assigned_Tau_star.csv").read().splitlines()
    assert len(star) == 41 and star[0] == ",Position,0,1,2,3,4,5,6,7,8,9,10,11"
    tau = np.loadtxt(out / "Assigned_Tau_star.csv", delimiter=",", skiprows=1, usecols=range(2, 14))
    assert (tau.reshape(40, 3, 4).sum(2) == 1).all()
    conf = np.loadtxt(out / "Assigned_Tau_conf.csv", delimiter=",", skiprows=1, usecols=(1, 2))
    want_pos = np.array([int(l.split(",")[1]) for l in lines[100:140]])
    assert np.array_equal(conf[:, 0], want_pos) and ((conf[:, 1] > 0) & (conf[:, 1] <= 1.0 + 1e-12)).all()
    # these positions are rows of the fitted table itself: with the fitted gamma/eta most of them are assigned with confidence
    assert np.median(conf[:, 1]) > 0.5
