"""Timeline of a sweep from %globaltimer stamps recorded by a -DKPROF build (diagnosis only; the product build has none of it).

    python tools/kprof.py --build                       (here: build/libdesman_b200_kprof.so)
    DESMAN_B200_LIB=build/libdesman_b200_kprof.so python tools/kprof.py > gpurun_out/kprof.txt   (on the GPU box)
"""
import ctypes
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "build", "libdesman_b200_kprof.so")

if "--build" in sys.argv:
    from desman_b200 import build as b
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    extra = [a for a in sys.argv[1:] if a.startswith("-D")]      # e.g. -DTC_NO_TCE: kernel scopes only, no per-item events
    cmd = ["nvcc"] + b.FLAGS + ["-DKPROF"] + extra + ["-o", OUT, b.SRC, "-ldl"]
    subprocess.run(cmd, check=True, capture_output=True)
    print(OUT)
    sys.exit(0)

from desman_b200 import _lib, engine
from desman_b200.synth import CHAIN_SEED, synth_counts

NAMES = ["maintain", "mu_binomial", "mu_class", "draw", "tau_group_mma", "tau_sample", "ll_table", "finalize", "copy_tau_if",
         "tau_warp", "tgm_prologue", "mub_warp", "tc_evt", "tau_open"]
REC = np.dtype([("kid", "i4"), ("cta", "i4"), ("warp", "i4"), ("x", "i4"), ("t0", "u8"), ("t1", "u8"), ("a", "u8"), ("b", "u8"),
                ("c", "u8"), ("d", "u8")])

flush = "--no-flush" not in sys.argv
p = synth_counts(100000, 64, 8)
e = engine.Engine(0, seed=CHAIN_SEED)
e.set_counts(p["counts"]); e.set_state(None, p["gamma0"], p["eta0"], G=8); e.set_tau_index(p["tau0"])
e.set_profiling(False, flush)
e.update(20)
L = _lib.lib()
L.desman_kprof_dump.restype = ctypes.c_int
L.desman_kprof_dump.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
L.desman_kprof_dump(None, 0, 1)
NS = 4
e.update(NS)
e.synchronize()
buf = np.zeros(1 << 18, dtype=REC)
n = L.desman_kprof_dump(buf.ctypes.data, len(buf), 1)
r = buf[:n]
print("records", n, "sweep ms", e.get_timing()["elapsed_ms"] / NS, e.get_group_stats())
t_base = int(r["t0"].min())
# launches: records of one kernel id whose [t0,t1] chain overlaps
rows = []
for kid in list(range(9)) + [13]:
    k = r[r["kid"] == kid]
    if not len(k):
        continue
    k = np.sort(k, order="t0")
    cur = None
    for x in k:
        if cur is None or int(x["t0"]) > cur["exit"]:
            cur = dict(kid=kid, entry=int(x["t0"]), exit=int(x["t1"]), first_exit=int(x["t1"]), last_entry=int(x["t0"]), n=0)
            rows.append(cur)
        cur["exit"] = max(cur["exit"], int(x["t1"])); cur["first_exit"] = min(cur["first_exit"], int(x["t1"]))
        cur["last_entry"] = max(cur["last_entry"], int(x["t0"])); cur["n"] += 1
rows.sort(key=lambda d: d["entry"])
prev = None
for d in rows:
    print("%-14s start %9.2f us  gap %6.2f  dur %7.2f  first-exit %7.2f  last-entry +%5.2f  ctas %d" % (
        NAMES[d["kid"]], (d["entry"] - t_base) / 1e3, ((d["entry"] - prev) / 1e3) if prev else 0.0, (d["exit"] - d["entry"]) / 1e3,
        (d["first_exit"] - d["entry"]) / 1e3, (d["last_entry"] - d["entry"]) / 1e3, d["n"]))
    prev = d["exit"]
to = [d for d in rows if d["kid"] == 13]
if to:
    g13 = r[(r["kid"] == 13) & (r["t0"] >= to[-1]["entry"]) & (r["t1"] <= to[-1]["exit"])]
    ex = (g13["t1"].astype(np.int64) - to[-1]["entry"]) / 1e3
    en = (g13["t0"].astype(np.int64) - to[-1]["entry"]) / 1e3
    print("tau_open last launch: ctas %d  cta (post-prologue) start us: med %.2f max %.2f  exit us: min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % (
        len(g13), np.median(en), en.max(), ex.min(), np.percentile(ex, 10), np.median(ex), np.percentile(ex, 90), ex.max()))
if to:
    w = r[(r["kid"] == 9) & (r["t0"] >= to[-1]["entry"] - 20000) & (r["t1"] <= to[-1]["exit"] + 1000)]
    if len(w):
        dur = (w["t1"].astype(np.int64) - w["t0"].astype(np.int64)) / 1e3
        sites = w["x"] & 0xff; n2 = (w["x"] >> 8) & 0xfff; n3 = (w["x"] >> 20) & 0xf; fl = (w["x"] >> 24) & 0xff
        print("tau_open CTAs with sites: %d; per CTA us (after prologue): med %.2f p90 %.2f max %.2f" % (int((sites > 0).sum()), np.median(dur[sites > 0]), np.percentile(dur[sites > 0], 90), dur.max()))
        for i in list(np.argsort(dur)[-6:]) + list(np.flatnonzero(sites > 0)[:6]):
            a = int(w["a"][i])
            print("  cta %d: %.2f us  sites %d rounds %d n2 %d n3 %d flips %d  stage %.2f steps %.2f | thread 0: lane_bounds %.2f own step %.2f barrier wait %.2f" % (
                w["cta"][i], dur[i], sites[i], a & 0xff, n2[i], n3[i], fl[i], w["b"][i] / 1e3, w["c"][i] / 1e3,
                ((a >> 8) & 0xffff) / 1e3, ((a >> 24) & 0xfffff) / 1e3, (a >> 44) / 1e3))
        has = sites > 0
        print("  per site: stage med %.2f p90 %.2f   rounds+writeback med %.2f p90 %.2f max %.2f" % (
            np.median(w["b"][has] / sites[has]) / 1e3, np.percentile(w["b"][has] / sites[has], 90) / 1e3,
            np.median(w["c"][has] / sites[has]) / 1e3, np.percentile(w["c"][has] / sites[has], 90) / 1e3, (w["c"][has] / sites[has]).max() / 1e3))
# the last screening launch: per-CTA prologue end and exit
tgl = [d for d in rows if d["kid"] == 4]
if tgl:
    tg = tgl[-1]
    g = r[(r["kid"] == 4) & (r["t0"] >= tg["entry"]) & (r["t1"] <= tg["exit"])]
    gp = r[(r["kid"] == 10) & (r["t0"] >= tg["entry"]) & (r["t0"] <= tg["exit"])]
    en = (g["t0"].astype(np.int64) - tg["entry"]) / 1e3
    ex = (g["t1"].astype(np.int64) - tg["entry"]) / 1e3
    pr = (gp["t0"].astype(np.int64) - tg["entry"]) / 1e3
    print("screening last launch: ctas %d  scope entry us: med %.2f max %.2f | prologue done us: med %.2f max %.2f | exit us: min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % (
        len(g), np.median(en), en.max(), np.median(pr) if len(pr) else -1, pr.max() if len(pr) else -1, ex.min(), np.percentile(ex, 10),
        np.median(ex), np.percentile(ex, 90), ex.max()))
    re_ = r[(r["kid"] == 12) & (r["warp"] >= 100) & (r["t0"] >= tg["entry"]) & (r["t0"] <= tg["exit"] + 100000)]
    rn = ["fetcher done", "copy issuer done", "mma issuer done", "builder 0 done", "epilogue items done", "epilogue flushed", "teardown sync passed", "builder 0 start"]
    for cta in sorted(set(re_["cta"].tolist())):
        ent = int(g[g["cta"] == cta]["t0"][0]); exi = int(g[g["cta"] == cta]["t1"][0])
        print("  cta %3d (us from its scope entry):  %s | exit %.2f" % (cta, "  ".join("%s %.2f" % (rn[int(w_) - 100], (int(t_) - ent) / 1e3)
              for w_, t_ in sorted(zip(re_[re_["cta"] == cta]["warp"].tolist(), re_[re_["cta"] == cta]["t0"].tolist()))), (exi - ent) / 1e3))
    for cta in sorted(set(re_["cta"].tolist())):
        a = re_[(re_["cta"] == cta) & (re_["warp"] == 107)]; b = re_[(re_["cta"] == cta) & (re_["warp"] == 103)]
        if len(a) and len(b):
            print("  cta %3d builder 0: %.2f us = %d SM cycles -> %.0f MHz" % (cta, (int(b["t0"][0]) - int(a["t0"][0])) / 1e3, int(b["t1"][0]) - int(a["t1"][0]),
                  (int(b["t1"][0]) - int(a["t1"][0])) / ((int(b["t0"][0]) - int(a["t0"][0])) / 1e3)))
    ev = r[(r["kid"] == 12) & (r["warp"] < 100) & (r["t0"] >= tg["entry"]) & (r["t0"] <= tg["exit"] + 100000)]
    if len(ev):
        cta = int(ev["cta"].min())
        ev = ev[ev["cta"] == cta]
        ent = int(g[g["cta"] == cta]["t0"][0])
        roles = ["tma issued", "mma committed", "table start (builder 0)", "table start (builder 15)", "table done (b0)", "table done (b15)",
                 "acc seen", "item done", "mma operands ready"]
        print("tc pipeline of cta %d (us from its scope entry); columns: %s" % (cta, ", ".join(roles)))
        for it in sorted(set(ev["x"].tolist())):
            row = []
            for ro in range(9):
                m = ev[(ev["x"] == it) & (ev["warp"] == ro)]
                row.append("%6.2f" % ((int(m["t0"][0]) - ent) / 1e3) if len(m) else "   -  ")
            print("  item %2d: %s" % (it, "  ".join(row)))
    tb = r[(r["kid"] == 12) & (r["warp"] >= 200)]
    if len(tb):
        nit = int(tb["x"].max()) + 1
        print("builder warps of cta 0, ns per item build (last launches mixed: the records carry durations); rows = warps 0-15, then builder 0's P loop / shuffles / entries")
        for w_ in range(19):
            m = tb[tb["warp"] == 200 + w_]
            last = {}
            for x_, t_ in zip(m["x"].tolist(), m["t0"].tolist()):
                last[x_] = t_
            print("  %2d: %s" % (w_, " ".join("%5d" % last.get(i_, 0) for i_ in range(nit))))
    dr = [d for d in rows if d["kid"] == 3 and d["exit"] <= tg["entry"] + 2000]
    if dr:
        print("  (draw kernel before it: exit at %.2f us relative to the first screening scope entry)" % ((dr[-1]["exit"] - tg["entry"]) / 1e3))
# the last tau_sample launch: per-warp phases
if not [d for d in rows if d["kid"] == 5]:
    print("(no tau_sample records: the work list was walked by tau_open_kernel)")
    sys.exit(0)
ts = [d for d in rows if d["kid"] == 5][-1]
w = r[(r["kid"] == 9) & (r["t0"] >= ts["entry"]) & (r["t1"] <= ts["exit"])]
print("tau_sample last launch: warps", len(w), "launch entry->exit us", (ts["exit"] - ts["entry"]) / 1e3)
if not len(w):
    w = r[:0]
    print("(tau_sample: no per-warp records: the work list was walked by tau_open_kernel)")
    raise SystemExit(0) if "--full" not in sys.argv else None
pro = (w["t0"].astype(np.int64) - ts["entry"]) / 1e3
end = (w["t1"].astype(np.int64) - ts["entry"]) / 1e3
sites = w["x"] & 0xff; n2 = (w["x"] >> 8) & 0xfff; n3 = (w["x"] >> 20) & 0xf; fl = (w["x"] >> 24) & 0xff
print("prologue done  us: min %.2f med %.2f max %.2f" % (pro.min(), np.median(pro), pro.max()))
print("warp end       us: min %.2f med %.2f p90 %.2f max %.2f" % (end.min(), np.median(end), np.percentile(end, 90), end.max()))
print("(per-warp records: every 8th CTA)  sites per warp: ", np.bincount(sites)[:6], " tier3 warps", int((n3 > 0).sum()), " flips", int(fl.sum()))
has = sites > 0
for nm in ("b", "c", "d"):
    v = w[nm][has] / np.maximum(sites[has], 1) / 1e3
    print("per-site %s us: med %.2f p90 %.2f max %.2f" % ({"b": "stage", "c": "steps", "d": "move"}[nm], np.median(v), np.percentile(v, 90), v.max()))
order = np.argsort(end)[-8:]
for i in order:
    print("  slow warp cta %d w %d: pro %.2f end %.2f sites %d n2 %d n3 %d flips %d stage %.2f steps %.2f move %.2f" % (
        w["cta"][i], w["warp"][i], pro[i], end[i], sites[i], n2[i], n3[i], fl[i], w["b"][i] / 1e3, w["c"][i] / 1e3, w["d"][i] / 1e3))
# the last tau_group launch: prologue and exits
tg = [d for d in rows if d["kid"] == 4][-1]
g = r[(r["kid"] == 4) & (r["t0"] >= tg["entry"]) & (r["t1"] <= tg["exit"])]
gp = r[(r["kid"] == 10) & (r["t0"] >= tg["entry"]) & (r["t0"] <= tg["exit"])]
ex = (g["t1"].astype(np.int64) - tg["entry"]) / 1e3
print("tau_group last launch: ctas", len(g), "prologue done us: med %.2f max %.2f" % (np.median((gp["t0"].astype(np.int64) - tg["entry"]) / 1e3), ((gp["t0"].astype(np.int64) - tg["entry"]) / 1e3).max()))
print("cta exit us: min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % (ex.min(), np.percentile(ex, 10), np.median(ex), np.percentile(ex, 90), ex.max()))
for kid in (1, 2):
    m = [d for d in rows if d["kid"] == kid][-1]
    g = r[(r["kid"] == kid) & (r["t0"] >= m["entry"]) & (r["t1"] <= m["exit"])]
    ex = (g["t1"].astype(np.int64) - m["entry"]) / 1e3
    print(NAMES[kid], "cta exit us: min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % (ex.min(), np.percentile(ex, 10), np.median(ex), np.percentile(ex, 90), ex.max()))
mb = [d for d in rows if d["kid"] == 1][-1]
w = r[(r["kid"] == 11) & (r["t0"] >= mb["entry"]) & (r["t1"] <= mb["exit"] + 2000)]
if len(w):
    it = np.maximum(w["x"], 1)
    print("mu_binomial warps %d items/warp med %d max %d | per warp us: ticket med %.2f  load med %.2f  draw med %.2f  total med %.2f max %.2f" % (
        len(w), np.median(w["x"]), w["x"].max(), np.median(w["a"]) / 1e3, np.median(w["b"]) / 1e3, np.median(w["c"]) / 1e3,
        np.median(w["t1"].astype(np.int64) - w["t0"].astype(np.int64)) / 1e3, (w["t1"].astype(np.int64) - w["t0"].astype(np.int64)).max() / 1e3))
    print("  per item us (lane 0): ticket %.2f load %.2f draw %.2f" % (np.median(w["a"] / it) / 1e3, np.median(w["b"] / it) / 1e3, np.median(w["c"] / it) / 1e3))

# tensor-memory screening pass: the pipeline of one CTA of the last launch (events of every role per item, us from CTA entry)
ev = r[(r["kid"] == 12) & (r["t0"] >= tg["entry"]) & (r["t0"] <= tg["exit"] + 100000)]
if len(ev):
    cta = int(ev["cta"].min())
    ev = ev[ev["cta"] == cta]
    ent = int(g[g["cta"] == cta]["t0"][0]) if (g["cta"] == cta).any() else tg["entry"]
    roles = ["tma issued", "mma committed", "table start (builder 0)", "table start (builder 15)", "table done (b0)", "table done (b15)",
             "acc seen", "item done", "mma operands ready"]
    print("tc pipeline of cta %d (us from its entry); columns: %s" % (cta, ", ".join(roles)))
    for it in sorted(set(ev["x"].tolist())):
        row = []
        for ro in range(9):
            m = ev[(ev["x"] == it) & (ev["warp"] == ro)]
            row.append("%6.2f" % ((int(m["t0"][0]) - ent) / 1e3) if len(m) else "   -  ")
        print("  item %2d: %s" % (it, "  ".join(row)))
