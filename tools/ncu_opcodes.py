#!/usr/bin/env python
"""Executed-instruction histogram by opcode for the kernels of an .ncu-rep (source page, SASS view).

usage: tools/ncu_opcodes.py report.ncu-rep <launch index in the report, 0-based> [pixels]
"""
import csv
import subprocess
import sys
from collections import Counter


def sections(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    out, cur = [], None
    for r in csv.reader(raw.splitlines()):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            out.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    return out[::2]  # ncu prints every launch twice on this page


def histogram(sec, px=None, top=28):
    hdr = sec["rows"][0]
    i_src, i_ex, i_st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, stalls = Counter(), Counter()
    n_lines = 0
    for r in sec["rows"][1:]:
        if len(r) <= max(i_src, i_ex, i_st):
            continue
        toks = r[i_src].split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        op = toks[0].rstrip(";") if toks else "?"
        head = op.split(".")[0]
        base = ".".join(op.split(".")[:2]) if head in ("LDG", "STG", "LDS", "STS", "SHFL", "BAR") else head
        ops[base] += int(r[i_ex])
        stalls[base] += int(r[i_st])
        n_lines += 1
    tot, stot = sum(ops.values()), sum(stalls.values())
    print("== %s" % sec["name"])
    print("total warp instructions %d%s, %d SASS lines" %
          (tot, ", %.1f thread-instr/px" % (tot * 32 / px) if px else "", n_lines))
    for op, n in ops.most_common(top):
        print("%-14s %12d %5.1f%%   stall samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * stalls[op] / max(stot, 1)))


if __name__ == "__main__":
    secs = sections(sys.argv[1])
    px = float(sys.argv[3]) if len(sys.argv) > 3 else None
    for i in ([int(sys.argv[2])] if len(sys.argv) > 2 and sys.argv[2] != "all" else range(len(secs))):
        histogram(secs[i], px)
