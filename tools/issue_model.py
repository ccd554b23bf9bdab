"""Issue-cycle budget of the FIR kernels from an `ncu --set full --import-source on` capture (no GPU needed).

    python tools/issue_model.py gpurun_out/prof_r2q.ncu-rep profiles/fir_issue_model_r2q.txt

Per kernel: executed warp instructions by opcode class (source page), the issue cycles they take if a packed FFMA2 / FADD2 and
an FP64 operation occupy the issue port for 2 and 2.28 cycles (tools/ubench.cu; `sm__pipe_fma_cycles_active` = 2 x
`sm__inst_executed_pipe_fma` in the same capture), the share of those cycles that is the algorithm's multiply-accumulates, and
what ncu measured next to it (issue-active, FMA-pipe-busy).  MAC share x pipe-busy fraction is what `fir_frac_no_fma` reports.
"""
from __future__ import annotations

import collections
import csv
import io
import subprocess
import sys

KERNELS = [("k_rf_frontend", 2 * 7350 * 101 * 1024), ("^k_fir_bank$", 3 * 7350 * 101 * 1024), ("k_fir_bank_scalar", 7350 * 101 * 1024),
           ("k_audio_decim", 2 * 1470 * 101 * 1024), ("k_rds_backend", 2 * 2836 * 101 * 1024)]  # algorithmic MACs per launch (1024 stations)


def page(rep, which, kernel):
    return subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout


def main(rep, out):
    lines = ["# kernel, warp instructions, packed FP32 (FFMA2+FADD2), scalar FP32 (FMUL+FADD+FFMA), FP64, loads/stores, other, issue cycles (model), "
             "MAC share of issue cycles, issue-active % (ncu), FMA-pipe-busy % (ncu), duration us (ncu)"]
    for kern, macs in KERNELS:
        rows = list(csv.reader(io.StringIO(page(rep, "source", kern))))
        hdr = next((r for r in rows if "Source" in r and "# Samples" in r), None)
        if hdr is None:
            continue
        idx = {h: i for i, h in enumerate(hdr)}
        c = collections.Counter()
        for r in rows[rows.index(hdr) + 1:]:
            if len(r) != len(hdr) or r[0] == "Address":
                break
            t = r[idx["Source"]].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            n = int(r[idx["Instructions Executed"]])
            if op in ("FFMA2", "FADD2", "FMUL2"):
                c["fp2"] += n
            elif op in ("FMUL", "FADD", "FFMA"):
                c["fp1"] += n
            elif op in ("DFMA", "DMUL", "DADD"):
                c["fp64"] += n
            elif op in ("LDS", "STS", "LDG", "STG", "LDSM", "UBLKCP", "LDL", "STL"):
                c["mem"] += n
            else:
                c["other"] += n
        total = sum(c.values())
        cycles = 2 * c["fp2"] + c["fp1"] + 2.28 * c["fp64"] + c["mem"] + c["other"]
        mac_cycles = 2 * macs / 32  # one MAC per lane = two issue cycles, packed or not
        raw = list(csv.reader(io.StringIO(page(rep, "raw", kern))))
        rh, rv = raw[0], raw[2]
        get = lambda k: next((rv[i] for i, h in enumerate(rh) if h == k), "")
        lines.append(f"{kern.strip("^$")}, {total}, {c['fp2']}, {c['fp1']}, {c['fp64']}, {c['mem']}, {c['other']}, {cycles:.0f}, {mac_cycles / cycles:.3f}, "
                     f"{float(get('smsp__issue_active.avg.pct_of_peak_sustained_active')):.1f}, "
                     f"{float(get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active')):.1f}, {float(get('gpu__time_duration.sum')):.1f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
