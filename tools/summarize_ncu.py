"""Turns the ncu outputs of a gpurun call into the small committed summaries under profiles/.

    python tools/summarize_ncu.py gpurun_out/launches_r1b.csv gpurun_out/prof_r1b.ncu-rep r1b
writes profiles/launches_<tag>_summary.csv, profiles/ncu_full_<tag>_summary.csv and profiles/ncu_traffic.json
(per-kernel DRAM bytes per launch, read by bench.py for roofline.traffic).
"""
from __future__ import annotations

import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = {"k_rf_frontend": "rf_frontend", "k_fir_bank<3": "if_bands", "k_fir_bank<2": "if_bands", "k_fir_bank_scalar<1": "rds_carrier_bpf",
       "k_fir_bank<1": "rds_carrier_bpf", "k_pll": "pll", "k_mix": "mix", "k_audio": "audio", "k_rds_backend": "rds_backend"}


def main(launches, rep, tag):
    rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        if r[idx["Metric Name"]] == "gpu__time_duration.sum":
            v = float(r[idx["Metric Value"]].replace(",", ""))
            u = r[idx["Metric Unit"]]
            v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v
            agg[r[idx["Kernel Name"]].split("(")[0]].append(v)
    tot = sum(sum(v) for k, v in agg.items() if "sdrb::" in k)
    out = ["# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e",
           "# serialised and cold-cache: compare SHARES.  share = of the chain's own kernels (sdrb::); at:: kernels build the inputs",
           "# kernel, launches, mean_us, total_us, share"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        share = f"{sum(v) / tot:.3f}" if "sdrb::" in k else "-"
        out.append(f"{k}, {len(v)}, {sum(v) / len(v):.1f}, {sum(v):.1f}, {share}")
    open(os.path.join(ROOT, "profiles", f"launches_{tag}_summary.csv"), "w").write("\n".join(out) + "\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum"]
    stall = [h for h in hdr if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    lines = ["# ncu --set full --clock-control none --import-source on, one step (7 kernels), 1024 streams, mode 0 r",
             ", ".join(want + ["top stall reasons (share of samples)"])]
    traffic = {}
    for r in rows[2:]:
        vals = sorted(((h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(r[idx[h]].replace(",", "") or 0)) for h in stall), key=lambda kv: -kv[1])
        st = sum(v for _, v in vals) or 1.0
        lines.append(", ".join(r[idx[w]] if w in idx else "n/a" for w in want) + ", " + " ".join(f"{k}={v / st:.2f}" for k, v in vals[:4]))
        name = r[idx["Kernel Name"]]
        for pat, key in KEY.items():
            if pat in name:
                def to_bytes(col):
                    v = float(r[idx[col]].replace(",", ""))
                    u = rows[1][idx[col]]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                traffic[key] = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
                break
    lines.append("# units: " + ", ".join(rows[1][idx[w]] if w in idx else "" for w in want))
    open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}_summary.csv"), "w").write("\n".join(lines) + "\n")
    json.dump({"source": f"profiles/ncu_full_{tag}_summary.csv", "streams": 1024, "dram_bytes_per_launch": traffic},
              open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print("\n".join(out[-12:]))
    print("\n".join(lines[1:]))


if __name__ == "__main__":
    main(*sys.argv[1:4])
