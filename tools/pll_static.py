"""Static schedule of k_pll<32>'s hot loop, without a GPU: compiles the kernel alone (5 s), finds the loop in the SASS,
walks its common path (the branches around the careful path are taken) and adds up ptxas' stall counts plus modelled
scoreboard waits (tools/sass_sched.py).  One warp per SM scheduler executes exactly this static schedule, so the figure
ranks source variants before any of them is timed on the box (measured cycles = static + taken branches + fetch).

    python tools/pll_static.py [-DSDRB_... ...] [--list]
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from sass_sched import VAR_LAT, ctrl, parse  # noqa: E402


def build(defs, threads=32):
    d = tempfile.mkdtemp(prefix="pllstatic")
    src = os.path.join(d, "pll_only.cu")
    open(src, "w").write('#include "sdr_kernels.cuh"\ntemplate __global__ void sdrb::k_pll<%d>(const sdrb::PllArgs);\n' % threads)
    out = os.path.join(d, "pll_only.cubin")
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
           "-I" + os.path.join(ROOT, "real-time-sdr_b200", "csrc"), "-I" + os.path.join(ROOT, "include"), "-cubin", "-o", out, src] + defs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    return out


def opname(text):
    t = text.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]


def main():
    defs = [a for a in sys.argv[1:] if a.startswith("-D") or a.startswith("-X") or a.startswith("--maxrreg")]
    ins = parse(build(defs), "k_pllILi32E")
    by_addr = {a: i for i, (a, _, _) in enumerate(ins)}
    # the hot loop: the backward branch whose body holds the most DFMA
    cands = []
    for i, (a, text, _) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < a:
            j = by_addr[int(m.group(1), 16)]
            cands.append((sum(1 for k in range(j, i) if opname(ins[k][1]) == "DFMA"), j, i))
    # the innermost loop that holds the recurrence: dense in DFMA (the careful path around it is integer code), smallest span
    dense = [c for c in cands if c[0] >= 60 and c[0] >= 0.12 * (c[2] - c[1])]
    voting = [c for c in dense if any(opname(ins[k][1]) == "VOTE" for k in range(c[1], c[2]))]  # the shipped loop votes on its flag
    _, lo, hi = min(voting or dense, key=lambda c: c[2] - c[1])
    t, bar, i, taken, n_issued, rows = 0, [0] * 6, lo, 0, 0, []
    per_op = {}
    while i <= hi:
        a, text, w = ins[i]
        c = ctrl(w)
        wait = 0
        for b in range(6):
            if (c["wait"] >> b) & 1 and bar[b] > t:
                wait = max(wait, bar[b] - t)
        t += wait
        op = opname(text)
        if c["wr"] != 7:
            bar[c["wr"]] = t + VAR_LAT.get(op, 20)
        if c["rd"] != 7:
            bar[c["rd"]] = max(bar[c["rd"]], t + 4)
        rows.append((a, t, c["stall"], wait, text))
        n_issued += 1
        st = max(1, c["stall"])
        e = per_op.setdefault(op, [0, 0])
        e[0] += 1
        e[1] += st + wait
        t += st
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m and i != hi:
            tgt = int(m.group(1), 16)
            if tgt > a and text.startswith("@"):
                j = by_addr[tgt]
                if any(opname(ins[k][1]) in ("CALL", "ATOM", "ATOMG", "RED", "REDG") for k in range(i + 1, j)):
                    i = j
                    taken += 1
                    continue
        i += 1
    n_f2f = sum(1 for r in rows if r[4].startswith("F2F.F64.F32") or " F2F.F64.F32" in r[4])
    print(f"# loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {n_issued} instructions on the common path, {t} static cycles, "
          f"{taken} taken forward branches + 1 back edge, body {(ins[hi][0] - ins[lo][0] + 16) / 1024:.1f} KB")
    n_div = sum(1 for r in rows if "BRA.DIV" in r[4])
    if n_div:
        print(f"# WARNING: {n_div} BRA.DIV in the loop: ptxas does not take the warp for converged at the vote (measured: +27 cycles per sample)")
    ops = sorted(per_op.items(), key=lambda kv: -kv[1][1])[:12]
    print("# " + ", ".join(f"{k} {v[0]}/{v[1]}" for k, v in ops))
    if "--list" in sys.argv:
        for a, tt, st, wv, text in rows:
            print(f"{a:05x} t={tt:5d} st={st:2d} w={wv:3d}  {text}")


if __name__ == "__main__":
    main()
