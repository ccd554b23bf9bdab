"""Finds the first block at which the CUDA chain and the oracle differ in any RDS-path stage (one station, long run)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=4000)
    ap.add_argument("--station", type=int, default=0)
    ap.add_argument("--start_check", type=int, default=0)
    args = ap.parse_args()
    g.build()
    capi = g._load("sdrb_capi", os.path.join(ROOT, "real-time-sdr_b200", "capi.py"))
    gen = g._load("sdrgen", os.path.join(ROOT, "real-time-sdr_b200", "sdrgen.py"))
    import oracle_py

    orc = oracle_py.Oracle()
    gg = gen.StationGenerator(gen.Station.for_stream(args.station))
    stages = ["fm_demod", "rds_band", "gen_pilot", "IPLL", "rds_dc", "rds_filt", "rds_clean", "carrier"]
    bb = 147000
    oc = orc.lib.orc_chain_create(0, ord("r"), 1)
    pcm_ref = np.zeros(2940, np.int16)
    first = {}
    with capi.Chain(0, "r", n_streams=1, keep_stages=True) as ch:
        for b in range(args.blocks):
            iq = gg.read(73500)
            ch.process_host(iq.reshape(1, bb))
            orc.lib.orc_chain_block(oc, iq, pcm_ref)
            if b < args.start_check:
                continue
            rec = ch.read_rds()[0]
            off, ns, nbt = C.c_int(0), C.c_int(0), C.c_int(0)
            sp, bp = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
            orc.lib.orc_chain_rds_block(oc, C.byref(off), C.byref(sp), C.byref(ns), C.byref(bp), C.byref(nbt))
            if "cdr" not in first and off.value != int(rec["cdr_offset"]):
                first["cdr"] = (b, off.value, int(rec["cdr_offset"]))
            if "nbits" not in first and nbt.value != int(rec["n_bits"]):
                first["nbits"] = (b, nbt.value, int(rec["n_bits"]))
            gp = C.POINTER(C.c_uint64)()
            ng = orc.lib.orc_chain_groups(oc, C.byref(gp))
            if "groups" not in first and ng != int(rec["n_groups"]):
                first["groups"] = (b, ng, int(rec["n_groups"]))
            for st in stages:
                if st in first:
                    continue
                cnt = C.c_int(0)
                p = orc.lib.orc_chain_stage(oc, st.encode(), C.byref(cnt))
                want = np.ctypeslib.as_array(p, shape=(cnt.value,))
                got = ch.stage(st)[0]
                ne = got.view(np.uint32) != want.view(np.uint32)
                if ne.any():
                    i = int(np.argmax(ne))
                    first[st] = (b, i, float(got[i]), float(want[i]), int(ne.sum()))
            if len(first) >= 6 or ("cdr" in first and b > first["cdr"][0] + 3):
                break
            if b % 500 == 0:
                print(f"# block {b} first={first}", file=sys.stderr, flush=True)
    print(json.dumps({"blocks_run": b + 1, "first_difference": {k: list(v) for k, v in first.items()}}))


if __name__ == "__main__":
    main()
