#!/usr/bin/env python3
"""Turns the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (tag r1, r1b, r1c, r2, ...)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)


def launches(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[hi + 1:]:
        name = r[kn].split("(")[0].replace("void ", "")
        v = float(r[mv].replace(",", ""))
        if r[mu] == "ns":
            v /= 1e3
        elif r[mu] == "ms":
            v *= 1e3
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in tot.values())
    with open(os.path.join(OUT, "%s_launch_list_summary.csv" % tag), "w") as f:
        f.write("kernel,launches,total_us,avg_us,share_of_captured_device_time\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%s,%d,%.1f,%.2f,%.4f\n" % (k, n, t, t / n, t / total))
    return tot, total


def details(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio"]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    with open(os.path.join(OUT, "%s_ncu_full_summary.csv" % tag), "w") as f:
        f.write("kernel,metric,unit,value\n")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
            d = {}
            for w in want[1:]:
                if w in ix:
                    f.write("%s,%s,%s,%s\n" % (name, w, units[ix[w]], r[ix[w]]))
                    d[w] = (r[ix[w]], units[ix[w]])
            stalls = sorted(((h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(r[i] or 0))
                             for h, i in ix.items() if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")),
                            key=lambda t: -t[1])[:6]
            for s, v in stalls:
                f.write("%s,stall_%s,per_issue,%.3f\n" % (name, s, v))

            def to_bytes(key):
                v, u = d[key]
                v = float(v)
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            out[name] = dict(dram_bytes=to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                             duration=d["gpu__time_duration.sum"])
    return out


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    g = os.path.join(ROOT, "gpurun_out")
    tot, total = launches(os.path.join(g, "launches_%s.csv" % tag), tag)
    d = details(os.path.join(g, "prof_%s_final.ncu-rep" % tag), tag)
    traffic = {k.replace("_kernel", "").replace("<8>", "").replace("<2>", ""): v["dram_bytes"] for k, v in d.items()}
    # the steps bench.py reports: tau update = screening pass + per-site kernel; mu/E statistics = class split + within-class split
    if "tau_group_tc" in traffic:        # round 2: tensor-memory screening pass + the kernels that walk its work list
        traffic["tau_update"] = traffic["tau_group_tc"] + traffic.get("tau_open", 0.0) + traffic.get("tau_sample", 0.0)
    elif "tau_group_mma" in traffic:
        traffic["tau_update"] = traffic["tau_group_mma"] + traffic.get("tau_sample", 0.0)
    if "mu_binomial" in traffic:
        traffic["mu_stats"] = traffic["mu_binomial"] + traffic.get("mu_class", 0.0)
    json.dump({"source": "ncu --set full --clock-control none, bench.py config c3 (V=100000 S=64 G=8), one launch each",
               "dram_bytes_per_launch": traffic}, open(os.path.join(OUT, "traffic.json"), "w"), indent=1)
    print(open(os.path.join(OUT, "%s_launch_list_summary.csv" % tag)).read())
    print(json.dumps(traffic))
