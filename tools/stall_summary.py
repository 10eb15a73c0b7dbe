#!/usr/bin/env python3
"""Summarise an `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` dump: stall reasons, opcode mix, hottest SASS lines."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Address")
    hdr = rows[hi]
    ix = {}
    for i, h in enumerate(hdr):
        ix.setdefault(h, i)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    data = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[ix["# Samples"]].isdigit():
            continue
        n, ie = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        st = {s: int(r[ix[s]]) for s in stalls}
        data.append((r[ix["Source"]].strip(), n, ie, st))
        tot.update(st)
    S = sum(tot.values())
    print("samples", S)
    for s, v in tot.most_common(8):
        print("  %-24s %6.1f%%" % (s, 100.0 * v / S))
    op, ops = collections.Counter(), collections.Counter()
    for src, n, ie, st in data:
        t = src.split()
        o = t[1] if t[0].startswith("@") else t[0]
        o = o.split(".")[0]
        op[o] += ie
        ops[o] += n
    T = sum(op.values())
    print("warp instructions", T)
    for o, v in op.most_common(14):
        print("  %-10s %5.1f%% of instructions, %5.1f%% of samples" % (o, 100.0 * v / T, 100.0 * ops[o] / S))
    print("hottest SASS lines")
    for i, (src, n, ie, st) in sorted(enumerate(data), key=lambda t: -t[1][1])[:top]:
        t2 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("  %5d %-58s samples %5d  exec %8d  %s" % (i, src[:58], n, ie, t2))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
