"""Where k_pll's measured cycles exceed ptxas' static schedule: joins the per-instruction warp-state samples of an
`ncu --set full --import-source on` capture with the stall counts decoded from the same library's SASS.

    python tools/pll_dyn_vs_static.py gpurun_out/pll.ncu-rep real-time-sdr_b200/libsdr_b200.so k_pllILi64E <cycles per loop iteration> [--list]
"""
from __future__ import annotations

import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_sched import ctrl, parse  # noqa: E402


def main():
    rep, lib, kern, cyc = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_pll"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    ex = [int(r[ix["Instructions Executed"]]) for r in data]
    hot = [i for i, e in enumerate(ex) if e >= 0.9 * max(ex)]
    lo, hi = hot[0], hot[-1]
    stat = {a: ctrl(w)["stall"] for a, _, w in parse(lib, kern)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(data[i][ix["# Samples"]]) for i in range(lo, hi + 1) if ex[i] >= 0.9 * max(ex))
    k = cyc / tot
    out = []
    agg = {}
    for i in range(lo, hi + 1):
        if ex[i] < 0.9 * max(ex):
            continue  # the careful path inside the loop
        r = data[i]
        addr = int(r[ix["Address"]], 16) & 0xFFFFF
        smp = int(r[ix["# Samples"]]) * k
        st = stat.get(addr, stat.get(addr & 0xFFFF, 0))
        top = max(stalls, key=lambda h: int(r[ix[h]]))
        out.append((addr, st, smp, top, r[ix["Source"]].strip()))
        for h in stalls:
            agg[h] = agg.get(h, 0) + int(r[ix[h]]) * k
    print(f"# {len(out)} instructions, {cyc:.0f} cycles per iteration; static stall sum {sum(o[1] for o in out)}")
    print("# " + ", ".join(f"{h[6:]} {v:.0f}" for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    # measured cycles are attributed to the instruction the warp is waiting to issue: compare with the stall count of the PREVIOUS one
    exc = []
    for j in range(1, len(out)):
        exc.append((out[j][2] - max(1, out[j - 1][1]), j))
    if "--list" in sys.argv:
        for j, o in enumerate(out):
            prev = max(1, out[j - 1][1]) if j else 0
            print(f"{o[0]:05x} static_prev={prev:2d} measured={o[2]:6.1f} {o[3][6:]:14s} {o[4]}")
    else:
        print("# largest excess of measured cycles over the preceding instruction's stall count")
        for e, j in sorted(exc, reverse=True)[:40]:
            o = out[j]
            print(f"{o[0]:05x} +{e:5.1f} (measured {o[2]:5.1f}) {o[3][6:]:14s} {o[4]}   <- {out[j-1][4]}")


if __name__ == "__main__":
    main()
