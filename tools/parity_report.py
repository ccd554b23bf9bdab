"""Stage-by-stage parity report of the CUDA chain against the oracle (never stops at the first difference).

    python tools/parity_report.py [--blocks 24] [--streams 3] [--configs 0r,0s,0m,2m,1s,3m]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from chain_compare import FLOAT_STAGES, RDS_KEYS, diff_report, run_cuda_chain  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=24)
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--configs", default="0r,0s,0m,2m,1s,3m")
    args = ap.parse_args()
    g.build()
    capi = g._load("sdrb_capi", os.path.join(ROOT, "real-time-sdr_b200", "capi.py"))
    gen = g._load("sdrgen", os.path.join(ROOT, "real-time-sdr_b200", "sdrgen.py"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    orc = oracle_py.Oracle()
    all_ok = True
    for cfg in args.configs.split(","):
        mode, kind = int(cfg[0]), cfg[1]
        t0 = time.time()
        iqs = [gen.generate_iq(gen.Station.for_stream(k, fs=gen.mode_fs(mode)), gen.block_pairs(mode) * args.blocks)
               for k in range(args.streams)]
        stages = FLOAT_STAGES[kind]
        try:
            got = run_cuda_chain(capi, mode, kind, iqs, args.blocks, stages=stages)
        except Exception as e:  # report and go on with the next configuration
            print(json.dumps({"config": cfg, "error": repr(e)}))
            all_ok = False
            continue
        for s in range(args.streams):
            want = orc.chain(mode, kind, iqs[s], stages=stages)
            rep = diff_report(got[s], want, ["pcm"] + stages + (RDS_KEYS if kind == "r" else []))
            bad = {k: v for k, v in rep.items() if v is not None}
            all_ok &= not bad
            print(json.dumps({"config": cfg, "stream": s, "blocks": args.blocks, "identical": sorted(k for k, v in rep.items() if v is None),
                              "different": bad, "groups": int(len(want.get("groups", [])))}, default=str))
        print(json.dumps({"config": cfg, "seconds": round(time.time() - t0, 1)}))
    print("PARITY", "OK" if all_ok else "FAILED")
    return 0 if all_ok else 1


if __name__ == "__main__":
    sys.exit(main())
