#!/usr/bin/env python
"""SASS opcode histogram per kernel of wild-gaussians_b200/lib/libgsrast.so (cuobjdump -sass), written to
profiles/<name>.  The mnemonics that prove the Blackwell paths: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR
(tcgen05.commit), UBLKCP (cp.async.bulk, the TMA engine), SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (paired fp32),
REDG (vector reductions), LDGSTS (cp.async)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "wild-gaussians_b200", "lib", "libgsrast.so")
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "UTCATOM", "FFMA2", "FMUL2", "FADD2", "REDG", "LDGSTS", "MUFU", "SHFL", "VOTE",
       "ATOMS", "LDG", "STG", "LDS", "STS", "BAR", "HMMA", "F2FP")


def main(out):
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1).split(".")[0]] += 1
    with open(out, "w") as f:
        f.write("SASS opcode histogram per kernel (cuobjdump -sass wild-gaussians_b200/lib/libgsrast.so, sm_100a)\n")
        f.write("columns: total instructions, then the counts of the mnemonics listed in tools/sass_histogram.py\n\n")
        for name, c in kernels.items():
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            f.write(f"{dem}\n    total {sum(c.values())}")
            for k in KEY:
                n = sum(v for op, v in c.items() if op.startswith(k))
                if n:
                    f.write(f"  {k} {n}")
            f.write("\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_histogram.txt"))
