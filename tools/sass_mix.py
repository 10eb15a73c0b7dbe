#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` dump: executed warp-instructions per opcode and per site."""
import collections
import csv
import sys


def main(path, units):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, samp, tot = collections.Counter(), collections.Counter(), 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit():
            continue
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
        op = op.split(".")[0]
        n = int(r[ix["Instructions Executed"]])
        ops[op] += n
        samp[op] += int(r[ix["# Samples"]] or 0)
        tot += n
    print("total warp-instructions %d = %.1f per unit (%d units)" % (tot, tot / units, units))
    for op, n in ops.most_common(30):
        print("%-10s %10.1f /unit %5.1f%%  stall samples %d" % (op, n / units, 100.0 * n / tot, samp[op]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0)
