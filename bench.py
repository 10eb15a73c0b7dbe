#!/usr/bin/env python3
"""bench.py -- Gibbs sweeps/sec of the DESMAN haplotype-inference hot path on B200.

    python bench.py --gpus N --steps K --warmup W            (our arm; N>1 under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   (CPU arm: reference/port)

A "step" is one full Gibbs sweep (HaploSNP_Sampler.update() loop body, HaploSNP_Sampler.py:341-358:
mu/E statistics, gamma draw, tau draw, eta draw, ll+lp, MAP tracking, running tau sums) over a
synthetic V x S x 4 count tensor (SURVEY.md section 8d generator).  Workload at N GPUs: config C3 of
BASELINE.json per GPU (V=100000, S=64, G=8), the sites of ONE chain sharded over the ranks
(V_total = N*100000) with one NCCL all-reduce of the S*G+16 statistics (+ ll, nchange) per sweep.
`value` counts sweeps of the 100000-site unit: value = (V_total/100000) * K / t, so that N=1 is plain
sweeps/s on C3 and the aggregate grows with N under weak scaling.

Timing: per-sweep CUDA events on the engine's own stream (device time), L2 flushed between sweeps
(256 MiB write, outside the timed events), summed over the K sweeps, max over ranks.  `e2e` is the
same K sweeps through the public plugin call HaploSNP_Sampler.update() from HOST numpy arrays
(counts/state upload and result download inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT_V = 100000


def workload_name(cfg_key, V, S, G):
    """The ONE workload string both arms print (the driver compares them)."""
    return "BASELINE config %s per GPU: synthetic V=%d S=%d G=%d, full Gibbs sweep (mu/E stats, gamma, tau, eta, ll/lp, MAP, tau sums)" % (
        cfg_key.upper(), V, S, G)


CONFIGS = {
    "c2": dict(V=10000, S=64, G=8),
    "c3": dict(V=100000, S=64, G=8),
    "c4": dict(V=100000, S=256, G=16),
    "c5": dict(V=125000, S=128, G=20),   # per-GPU shard of V=1e6 at 8 GPUs
}


# ---------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(kernel, V, S, G, slots=None):
    """DESIGN.md section 5 / SURVEY.md 8d: bytes one launch must move under the canonical packed layout."""
    if kernel in ("tau_sample", "tau_update"):      # counts once, tau read+write, gamma, eta
        return 16 * V * S + 2 * V * G + 8 * S * G + 128
    if kernel == "mu_stats" and slots:   # pattern-aggregated form: the table N[slot][s][4] (uint64) once, codes, gamma, eta, statistics
        return 32 * slots * S + 8 * slots + 8 * S * G + 128 + 8 * (S * G + 16)
    if kernel == "mu_stats":        # per-read form: counts once, tau read, gamma, eta, statistics out
        return 16 * V * S + V * G + 8 * S * G + 128 + 8 * (S * G + 16)
    raise KeyError(kernel)


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != n_gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (n_gpus, world))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(local)
        td.init_process_group(backend="nccl")
    return rank, world, local, td


def barrier(td):
    if td is not None:
        import torch
        td.barrier()
        torch.cuda.synchronize()


def max_over_ranks(td, x):
    if td is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def bcast_bytes(td, b, n):
    if td is None:
        return b
    import torch
    t = torch.zeros(n, dtype=torch.uint8, device="cuda")
    if td.get_rank() == 0:
        t = torch.tensor(list(b), dtype=torch.uint8, device="cuda")
    td.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


# ---------------------------------------------------------------------------------------------- our arm
def run_b200(args):
    from numpy.random import RandomState

    from desman_b200 import _lib, engine
    from desman_b200.HaploSNP_Sampler import HaploSNP_Sampler
    from desman_b200.synth import CHAIN_SEED, synth_counts

    rank, world, local, td = dist_setup(args.gpus)
    cfg = CONFIGS[args.config]
    V, S, G = cfg["V"], cfg["S"], cfg["G"]
    K, W = args.steps, max(args.warmup, 3)
    if args.strong:                     # strong scaling: the config's V sites shared out over the ranks
        V = V // world
    V_total = V * world
    if _lib.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; desman_b200 has no CPU fallback")
    p = synth_counts(V, S, G, shard=rank)
    from desman_b200.synth import onehot as onehot_from_index

    def make_engine():
        e = engine.Engine(local, seed=CHAIN_SEED)
        e.set_counts(p["counts"], v0=rank * V, V_total=V_total)
        if world > 1:
            uid = bcast_bytes(td, engine.Engine.comm_unique_id() if rank == 0 else b"", 128)
            e.comm_init(uid, rank, world)
        e.set_state(None, p["gamma0"], p["eta0"], G=G)
        e.set_tau_index(p["tau0"])
        return e

    # ---- device-resident timing: `value`
    e = make_engine()
    exchange_name = e.comm_kind()
    e.set_profiling(False, not args.no_flush)
    e.update(W)                                           # warm-up sweeps (untimed)
    clocks = ClockSampler(local)
    barrier(td); e.synchronize()
    clocks.start()
    t_wall0 = time.perf_counter()
    res = e.update(K)                                     # EXACTLY K timed sweeps
    e.synchronize(); barrier(td)
    t_wall = time.perf_counter() - t_wall0
    clk = clocks.stop()
    tm = e.get_timing()
    ms_total = max_over_ranks(td, tm["elapsed_ms"])
    # counted by the engine at every launch site: per sweep table_maintain, mu_binomial, mu_class, draw_gamma_eta,
    # tau_group_mma, tau_sample, ll_table, finalize_sweep, copy_tau_if; + the pre-sweep maintain/ll/finalize/copy pass and
    # flush_tau_counts (the L2 flush writes sit outside the timed events and are not counted)
    launches = int(sum(tm["kernel_launches"].values()))
    grp = e.get_group_stats()
    rank_check = None
    if td is not None:     # every rank draws gamma / eta itself from the exchanged statistics: they must agree bit for bit
        import torch
        _, g_now, e_now = e.get_state(want_tau=False)
        mine = torch.tensor(np.concatenate([g_now.ravel(), e_now.ravel()]).view(np.int64), device="cuda")
        allg = [torch.empty_like(mine) for _ in range(world)]
        td.all_gather(allg, mine)
        same = all(bool(torch.equal(allg[0], x)) for x in allg[1:])
        if not same:
            raise SystemExit("bench.py: gamma/eta differ between ranks after %d sweeps (exchange %s)" % (K, exchange_name))
        rank_check = "gamma and eta bit-identical on %d ranks after the timed sweeps" % world
    value = (V_total / UNIT_V) * K / (ms_total / 1e3)

    # ---- per-kernel pass (events around every launch) for the roofline object
    e.set_profiling(True, not args.no_flush)
    Kp = min(K, 20)
    e.update(Kp)
    tk = e.get_timing()
    kms = {k: v / Kp for k, v in tk["kernel_ms"].items()}
    # the tau update = screening pass over the pattern groups (tau_group) + the kernels that walk the sites it left undecided.
    # Timed as what it is in production -- ONE dependent chain of launches -- by a second pass with a single event pair around
    # it (an event between two launches undoes their programmatic overlap and adds its own gap; the per-kernel figures above
    # therefore add up to more than this).
    kms["tau_update_sum_of_kernels"] = kms["tau_group"] + kms["tau_sample"]
    e.set_profiling(2, not args.no_flush)
    e.update(Kp)
    kms["tau_update"] = e.get_timing()["kernel_ms"]["tau_update"] / Kp
    peak, peak_src = measured_hbm_peak()
    # the roofline object describes the tau update: the kernel pair the north star names and the only part of the sweep that
    # streams the count tensor (the statistics read the ~3 MB pattern table at G <= 8; their line is printed beside it)
    dom = "tau_update"

    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if args.config == "c3" and os.path.exists(tpath):       # dram bytes per launch from the committed ncu --set full capture
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch", {})

    def roof(kernel):
        b = algorithmic_bytes(kernel, V, S, G, slots=grp["slots"] if engine.auto_mu_mode(V, G) == 1 else None)
        ach = b / (kms[kernel] * 1e-3) / 1e9
        return dict(kernel=kernel, bound="hbm", achieved=ach, peak=peak, unit="GB/s", frac=ach / peak,
                    traffic=traffic.get(kernel),
                    algorithmic_bytes=b, ms=kms[kernel], peak_source=peak_src)
    e.close()

    # ---- NMFT initialiser on the same tensor (reported separately): device time of the iteration loop
    nmft = None
    if rank == 0 and not args.no_nmft:
        rng = np.random.default_rng(1)
        tau0 = rng.dirichlet(np.full(4, 0.01), size=V * G).reshape(V, G, 4).transpose(2, 0, 1).reshape(4 * V, G).copy()
        gam0 = rng.dirichlet(np.full(G, 0.01), size=S).T.copy()
        en = engine.Engine(local, seed=1)
        en.nmft_factorize(p["counts"], tau0, gam0, max_iter=64, min_change=0.0)          # warm-up (allocation, first launches)
        en.nmft_factorize(p["counts"], tau0, gam0, max_iter=192, min_change=0.0)
        ms, iters = engine.Engine.nmft_last_timing()
        en.close()
        per = ms / max(iters, 1)
        nb = 8 * 4 * V * S + 2 * 8 * 4 * V * G      # X once, tau read + written, per iteration (f64)
        nmft = {"iters_per_s": 1e3 / per, "ms_per_iter": per, "iters_timed": iters,
                "roofline": {"bound": "hbm", "achieved": nb / (per * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": nb / (per * 1e-3) / 1e9 / peak, "algorithmic_bytes": nb, "peak_source": peak_src}}

    # ---- end to end through the plugin call, host buffers in, host results out
    tau_host = onehot_from_index(p["tau0"])
    uid = None
    if world > 1:
        uid = bcast_bytes(td, engine.Engine.comm_unique_id() if rank == 0 else b"", 128)
    hs = HaploSNP_Sampler(p["counts"], G, RandomState(CHAIN_SEED), max_iter=K, device=local, seed=CHAIN_SEED,
                          shard=(rank * V, V_total), comm=(uid, rank, world) if world > 1 else None)
    hs.tau, hs.gamma, hs.eta = tau_host, p["gamma0"].copy(), p["eta0"].copy()
    barrier(td)
    t0 = time.perf_counter()
    hs.update()                                           # upload counts+state, K sweeps, download results
    t_e2e = time.perf_counter() - t0
    barrier(td)
    t_e2e = max_over_ranks(td, t_e2e)
    h2d = p["counts"].nbytes + tau_host.nbytes + hs.gamma.nbytes + hs.eta.nbytes
    d2h = (hs.gamma_store.nbytes + hs.eta_store.nbytes + 3 * 8 * K + 2 * tau_host.nbytes + hs._tau_sum.nbytes +
           2 * (hs.gamma.nbytes + hs.eta.nbytes))
    e2e_value = (V_total / UNIT_V) * K / t_e2e
    hs.close()

    if td is not None:
        td.barrier()
        td.destroy_process_group()
    if rank != 0:
        return
    out = {
        "metric": "Gibbs sweeps/sec (V variants x S samples x G strains)", "value": value, "unit": "sweeps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True,
        "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, V, S, G),
                   "sharding": "V_total=%d sharded over %d GPU(s), value in sweeps of the %d-site unit" % (V_total, world, UNIT_V),
                   "V_per_gpu": V, "V_total": V_total, "S": S, "G": G, "rng": "philox4x32-10 counter contract",
                   "l2": "not flushed" if args.no_flush else "flushed between sweeps (256 MiB write outside the timed events)",
                   "parallelism": "variant-position shard x%d, 1 exchange of S*G+16 int64 + 2 words per sweep" % world,
                   "collective": ("none" if world == 1 else exchange_name),
                   "e2e_call": "HaploSNP_Sampler.update() with max_iter=%d from host numpy arrays" % K},
        "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                "seconds": t_e2e},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof(dom),
        "roofline_tau_sample": roof("tau_update"),
        "roofline_mu_stats": roof("mu_stats"),
        "tau_groups": grp,
        "nmft": nmft,
        "rank_consistency": rank_check,
        "kernel_ms_per_sweep": kms,
        "wall_s_timed_region": t_wall,
        "chain": {"lp_first": float(res["lp_store"][0]), "lp_last": float(res["lp_store"][-1]),
                  "nchange_last": int(res["nchange"][-1])},
    }
    if world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(p, V, S, G, budget_s=args.cpu_seconds)
    emit(out)


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_sweep_sample(p, Vs, S, G, n_sweeps, use_ref_tau):
    """n_sweeps full sweeps on the first Vs sites: oracle port on all cores; when use_ref_tau the tau
    step is the reference's own compiled c_sample_tau (single-threaded, as the reference is)."""
    from oracle import oracle
    from desman_b200.synth import onehot as onehot_from_index
    tau = onehot_from_index(p["tau0"][:Vs])
    gamma, eta = p["gamma0"].copy(), p["eta0"].copy()
    counts = np.ascontiguousarray(p["counts"][:Vs])
    ref = oracle.RefSampleTau(23724839) if use_ref_tau else None
    t0 = time.perf_counter()
    for k in range(n_sweeps):
        sm, es = oracle.mu_stats(tau, gamma, eta, counts, 23724839, k)
        gamma = oracle.draw_gamma(sm, 0.1, 1e-6, 23724839, k)
        if ref is not None:
            ref.sample_tau(tau, gamma, eta, counts)
        else:
            oracle.sample_tau_philox(tau, gamma, eta, counts, 23724839, k)
        eta = oracle.draw_eta(es, 0.1, 23724839, k)
        oracle.logpost(tau, gamma, eta, counts)
    dt = time.perf_counter() - t0
    if ref is not None:
        ref.close()
    return dt / n_sweeps


def cpu_baseline(p, V, S, G, budget_s=15.0):
    from oracle import oracle
    cores = oracle.num_threads()
    Vs = min(V, 2000)
    t = cpu_sweep_sample(p, Vs, S, G, 1, False)
    n = max(1, min(10, int(budget_s / max(t, 1e-3))))
    t = cpu_sweep_sample(p, Vs, S, G, n, False)
    per_sweep_full = t * V / Vs
    return {"value": (V / UNIT_V) / per_sweep_full, "unit": "sweeps/s", "cores": cores, "kind": "port",
            "sample": "oracle C port (OpenMP, %d threads), %d sweeps on the first %d of %d sites, scaled linearly in V "
                      "(every step is a loop over independent sites); the reference's own Python sweep is ~600 s/sweep "
                      "on this config (BASELINE.md)" % (cores, n, Vs, V)}


def run_reference(args):
    """CPU arm: the reference's path on the host cores.  tau step = the reference's own c_sample_tau.c
    (oracle/_ref, compiled here from /root/reference; single-threaded like the reference); the numpy/Python
    steps of the reference (sampleMu, sampleGamma, sampleEta, logLikelihood) cannot travel to the GPU box and
    are timed through their C port (oracle/, OpenMP on all cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every host core whatever launched it
    ncpu = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncpu)
    from oracle import oracle
    oracle.set_num_threads(ncpu)
    from desman_b200.synth import synth_counts
    cfg = CONFIGS[args.config]
    V, S, G = cfg["V"], cfg["S"], cfg["G"]
    K, W = args.steps, args.warmup
    world = args.gpus
    use_ref = oracle.have_ref()
    # bounded sample: size Vs so that one step takes ~0.5 s and the whole --steps K --warmup W run about two minutes at most
    p = synth_counts(min(V, 20000), S, G, shard=0)
    t_probe = cpu_sweep_sample(p, 1000, S, G, 1, use_ref)
    t_step = min(0.5, 120.0 / max(K + W, 1))
    Vs = int(max(200, min(p["counts"].shape[0], 1000 * t_step / max(t_probe, 1e-4))))
    for _ in range(W):
        cpu_sweep_sample(p, Vs, S, G, 1, use_ref)
    t = cpu_sweep_sample(p, Vs, S, G, K, use_ref)
    per_sweep_unit = t * UNIT_V / Vs                       # seconds per sweep of the 100000-site unit
    value = 1.0 / per_sweep_unit                           # CPU arm: one host, independent of --gpus
    cores = oracle.num_threads()
    # honest label: only the tau step is the reference's own code (its numpy steps cannot travel to the GPU box)
    kind = "reference-tau+port" if use_ref else "port"
    sample = ("%d timed sweeps on %d of %d sites, scaled linearly in V; tau step: %s; other steps: C port of the reference's "
              "numpy code on %d OpenMP threads" % (K, Vs, V, "reference c_sample_tau.c (oracle/_ref, 1 thread)" if use_ref
                                                   else "C port (all threads)", cores))
    out = {"impl": "reference", "metric": "Gibbs sweeps/sec (V variants x S samples x G strains)", "value": value,
           "unit": "sweeps/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": per_sweep_unit * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args.config, V, S, G),
                      "V_per_gpu": V, "S": S, "G": G},
           "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def emit(out):
    """The ONE JSON line of the contract, on the process's real stdout."""
    os.write(_REAL_STDOUT, (json.dumps(out) + "\n").encode())


_REAL_STDOUT = 1


def main():
    # Libraries below us may write to fd 1 (NCCL prints its version banner there when NCCL_DEBUG is set in the box's
    # environment): everything but the result line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500, help="timed sweeps (default: the 500 iterations of BASELINE config C3)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between sweeps")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-nmft", action="store_true", help="skip the NMFT iterations/s leg")
    ap.add_argument("--strong", action="store_true", help="strong scaling: V of the config divided over the ranks (default: V per rank)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
