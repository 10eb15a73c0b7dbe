#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import collections
import csv
import sys


def main(path, units, top=40):
    rows = list(csv.reader(open(path)))
    cur_file = ""
    agg = collections.OrderedDict()
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            # duplicated 'Source' header: first is the CUDA line, second the SASS
            continue
        if hdr is None or len(r) < 10:
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        ie = r[hdr["Instructions Executed"]]
        sm = r[hdr["# Samples"]]
        if not ie.isdigit():
            continue
        key = (cur_file, line)
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += int(ie)
        a[1] += int(sm) if sm.isdigit() else 0
    tot = sum(a[0] for a in agg.values())
    tots = sum(a[1] for a in agg.values())
    print("total %d warp-instr = %.1f per unit; samples %d" % (tot, tot / units, tots))
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-16s %4d  inst/unit %8.1f (%4.1f%%)  samples %5.1f%%  | %s" % (f, ln, a[0] / units, 100.0 * a[0] / tot,
                                                                              100.0 * a[1] / max(tots, 1), a[2]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 40)
