"""Per-instruction stall profile of k_pll's hot loop from an `ncu --set full --import-source on` capture.

    python tools/pll_source_profile.py gpurun_out/prof_r1g.ncu-rep profiles/pll_hot_loop_r1g.txt [cycles_per_iteration]

One warp runs per SM scheduler, so the warp-state samples of an instruction are the cycles the whole recurrence spent
at it.  The summary lists the loop's stall reasons and, per region of the loop body (two 4-sample chunks since the loop is unrolled by two) (found from the F2F that
widens the loop filter's phase once per sample), instructions and cycles (samples scaled to cycles_per_chunk, the
measured 4 x cycles per sample of bench.py).
"""
from __future__ import annotations

import csv
import io
import subprocess
import sys


def main(rep, out, cycles_per_chunk):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_pll"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    rows = rows[starts[0]:starts[1]] if len(starts) > 1 else rows  # first captured launch of the kernel
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
    ex = [int(r[idx["Instructions Executed"]]) for r in data]
    hot = [i for i, e in enumerate(ex) if e >= 0.9 * max(ex)]
    lo, hi = hot[0], hot[-1]
    smp = [int(data[i][idx["# Samples"]]) for i in range(lo, hi + 1)]
    src = [data[i][idx["Source"]].strip() for i in range(lo, hi + 1)]
    total, allk = sum(smp), sum(int(r[idx["# Samples"]]) for r in data)
    k = cycles_per_chunk / total
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(data[i][idx[h]]) for i in range(lo, hi + 1)) for h in stalls}
    lines = [f"# k_pll hot loop: {len(hot)} instructions per loop iteration, {total} of {allk} warp-state samples ({total / allk:.1%} of the kernel)",
             f"# samples scaled so that the loop is {cycles_per_chunk:.0f} cycles per iteration (bench.py: cycles per sample x samples per iteration)",
             "# stall reason, share of the loop's samples"]
    for h, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
        lines.append(f"{h}, {v / total:.3f}")
    # regions: each F2F.F64.F32 in the loop body that feeds a DADD within 2 instructions marks 'phase -> double' of one sample
    marks = [i for i, s in enumerate(src) if s.startswith("F2F.F64.F32") and any(x.startswith("DADD") for x in src[i + 1:i + 4])]
    lines.append("# region (instruction rows), instructions, cycles")
    edges = [0] + marks + [len(src)]
    names = ["loop top .. sample 1 phase->double"] + [f"sample {j + 1} NCO phase .. sample {j + 2} phase->double" for j in range(len(marks) - 1)] + \
            [f"sample {len(marks)} NCO phase .. end of chunk (remaining samples, acceptance tests, next chunk's inputs, loop branch)"]
    for a, b, n in zip(edges[:-1], edges[1:], names):
        lines.append(f"{n} ({a}-{b - 1}), {b - a}, {sum(smp[a:b]) * k:.0f}")
    by = {}
    for s, v in zip(src, smp):
        op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
        c = by.setdefault(op, [0, 0])
        c[0] += 1
        c[1] += v
    lines.append("# opcode, instructions per chunk, cycles per chunk")
    for op, (n, v) in sorted(by.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append(f"{op}, {n}, {v * k:.0f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1260.0)
