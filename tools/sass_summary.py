"""Per-kernel SASS opcode summary of libsdr_b200.so (evidence of what the compiler emitted: TMA bulk copies, mbarrier
waits, cp.async, packed FP32, FP64, conversions ...).

    python tools/sass_summary.py [--out profiles/sass_summary_r2.csv]
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "FFMA2", "FADD2", "FMUL2", "FFMA", "FMUL", "FADD", "DFMA", "DMUL", "DADD",
         "F2F", "MUFU", "LOP3", "SHF", "IMAD", "ISETP", "BRA", "BAR", "SHFL", "VOTE", "POPC", "ATOMS", "RED"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default=os.path.join(ROOT, "real-time-sdr_b200", "libsdr_b200.so"))
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    txt = subprocess.run(["cuobjdump", "-sass", args.lib], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = counts.setdefault(re.sub(r"\(.*", "", name), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["_total"] += 1
    rows = ["kernel,instructions," + ",".join(WATCH)]
    for k, c in counts.items():
        rows.append(k + "," + str(c["_total"]) + "," + ",".join(str(c.get(op, 0)) for op in WATCH))
    out = "\n".join(rows) + "\n"
    if args.out:
        open(args.out, "w").write(out)
    sys.stdout.write(out)


if __name__ == "__main__":
    main()
