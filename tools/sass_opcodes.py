#!/usr/bin/env python
"""Static opcode histogram of every kernel in csrc/*.o (cuobjdump -sass) and the tcgen05 / TMA lines of the PTX.

usage: tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt      (after csrc/build.sh)
"""
import glob
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "reflectance-filtering_b200", "csrc")
MARK = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "MUFU", "FFMA2", "FADD2", "FMUL2",
        "ELECT", "R2UR")


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    print("# cuobjdump -sass of reflectance-filtering_b200/csrc/*.o (sm_100a): static instruction counts per kernel.")
    print("# Tensor-core / tensor-memory opcodes: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit);")
    print("# TMA: UTMALDG (cp.async.bulk.tensor global -> shared), SYNCS (mbarrier).")
    for obj in sorted(glob.glob(os.path.join(CSRC, "*.o"))):
        sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        cur, ops = None, None
        kernels = []
        for ln in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", ln)
            if m:
                cur, ops = m.group(1), Counter()
                kernels.append((cur, ops))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", ln)
            if m and ops is not None:
                ops[m.group(1).split(".")[0]] += 1
        for name, ops in kernels:
            tot = sum(ops.values())
            print("\n== %s :: %s" % (os.path.basename(obj), demangle(name)))
            print("   %d instructions; " % tot + ", ".join("%s %d" % (k, ops[k]) for k in MARK if ops.get(k)))
            print("   top: " + ", ".join("%s %d" % kv for kv in ops.most_common(12)))
    print("\n# PTX (cuobjdump -ptx is empty for -gencode code=sm_100a objects; the inline asm of the sources is listed instead)")
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        hits = Counter()
        for ln in open(src):
            for m in re.finditer(r"(tcgen05\.[a-z0-9_.:]+|cp\.async\.bulk\.tensor[a-z0-9_.:]*|mbarrier\.[a-z0-9_.:]+|elect\.sync|"
                                 r"fma\.rn\.f32x2|add\.rn\.f32x2|mul\.rn\.f32x2|ex2\.approx\.ftz\.f32)", ln):
                hits[m.group(1)] += 1
        if hits:
            print("%s: " % os.path.basename(src) + ", ".join("%s x%d" % kv for kv in sorted(hits.items())))


if __name__ == "__main__":
    main()
