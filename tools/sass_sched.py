"""Static issue timeline of a SASS basic-block range: decodes the control words (stall count, barriers) of
`cuobjdump -sass` output and lists, per instruction, the cycle at which the in-order warp can issue it
(sum of stall counts; scoreboard waits modelled with the latencies below).

    python tools/sass_sched.py <lib.so> <kernel-substring> <start-addr-hex> <end-addr-hex> [--list]

One warp per SM scheduler (k_pll) executes exactly this static schedule, so the sum is the loop's period up to the
variable-latency waits.
"""
from __future__ import annotations

import re
import subprocess
import sys

VAR_LAT = {"F2F": 19, "MUFU": 18, "LDS": 26, "LDG": 400, "LD": 400, "I2F": 14, "F2I": 14, "LDC": 30, "DFMA": 8, "DADD": 8, "DMUL": 8,
           "STS": 10, "STG": 20, "ST": 20, "LDGSTS": 30, "DSETP": 10, "S2R": 20, "POPC": 12, "FLO": 12, "BREV": 12, "SHFL": 24}


def parse(lib, kern):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    out, on, pend = [], False, None
    for ln in txt:
        if "Function :" in ln:
            on = kern in ln
            continue
        if not on:
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", ln)
        if m:
            pend = [int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16)]
            continue
        m = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", ln)
        if m and pend:
            hi = int(m.group(1), 16)
            out.append((pend[0], pend[1], hi))
            pend = None
    return out


def ctrl(hi):
    return dict(stall=(hi >> 41) & 0xF, yld=(hi >> 45) & 1, wr=(hi >> 46) & 7, rd=(hi >> 49) & 7, wait=(hi >> 52) & 0x3F)


def main():
    lib, kern, a0, a1 = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
    ins = [i for i in parse(lib, kern) if a0 <= i[0] <= a1]
    t = 0
    bar_ready = [0] * 6
    rows = []
    for addr, text, hi in ins:
        c = ctrl(hi)
        waited = 0
        for b in range(6):
            if (c["wait"] >> b) & 1 and bar_ready[b] > t:
                waited = max(waited, bar_ready[b] - t)
        t += waited
        op = (text.split()[1] if text.startswith("@") else text.split()[0]).split(".")[0]
        if c["wr"] != 7:
            bar_ready[c["wr"]] = t + VAR_LAT.get(op, 20)
        if c["rd"] != 7:
            bar_ready[c["rd"]] = max(bar_ready[c["rd"]], t + 4)
        rows.append((addr, t, c["stall"], waited, c["wr"], c["wait"], text))
        t += max(1, c["stall"])
    print(f"# {len(ins)} instructions, {t} cycles (static), stall sum {sum(r[2] for r in rows)}, barrier waits {sum(r[3] for r in rows)}")
    if "--list" in sys.argv:
        for addr, tt, st, w, wr, wm, text in rows:
            print(f"{addr:05x} t={tt:5d} st={st:2d} w={w:3d} wr={wr if wr != 7 else '-'} wm={wm:02x}  {text}")


if __name__ == "__main__":
    main()
