#!/usr/bin/env python
"""Basic-block view of one kernel in an .ncu-rep: contiguous SASS runs with the same execution count, with
their opcode mix.  usage: tools/ncu_blocks.py report.ncu-rep <launch index> [min share %]"""
import sys
from collections import Counter
sys.path.insert(0, __file__.rsplit("/", 1)[0])
from ncu_opcodes import sections  # noqa: E402

sec = sections(sys.argv[1])[int(sys.argv[2])]
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
hdr = sec["rows"][0]
i_src, i_ex, i_st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
rows = [r for r in sec["rows"][1:] if len(r) > max(i_src, i_ex, i_st)]
tot = sum(int(r[i_ex]) for r in rows)
stot = sum(int(r[i_st]) for r in rows)
print("==", sec["name"], "total", tot)
blocks = []
cur = None
for k, r in enumerate(rows):
    n = int(r[i_ex])
    if cur is None or cur["n"] != n:
        cur = {"n": n, "first": k, "ops": Counter(), "len": 0, "st": 0}
        blocks.append(cur)
    toks = r[i_src].split()
    if toks and toks[0].startswith("@"):
        toks = toks[1:]
    cur["ops"][toks[0].rstrip(";").split(".")[0] if toks else "?"] += 1
    cur["len"] += 1
    cur["st"] += int(r[i_st])
for b in blocks:
    share = 100.0 * b["n"] * b["len"] / tot
    if share >= min_share:
        print("lines %5d..%5d  exec/line %9d  share %5.1f%%  stall %5.1f%%  %s" %
              (b["first"], b["first"] + b["len"] - 1, b["n"], share, 100.0 * b["st"] / max(stot, 1),
               " ".join("%s:%d" % kv for kv in b["ops"].most_common(12))))
