"""Long-run parity soak: a few stations for many blocks (NCO phase far into the coarse-float regime), CUDA chain vs oracle.

    python tools/soak.py [--blocks 1200] [--stations 2]
Prints one JSON line per station: PCM / bit / group equality and how many PCM samples differ (expected 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=1200)
    ap.add_argument("--stations", type=int, default=2)
    args = ap.parse_args()
    g.build()
    capi = g._load("sdrb_capi", os.path.join(ROOT, "real-time-sdr_b200", "capi.py"))
    gen = g._load("sdrgen", os.path.join(ROOT, "real-time-sdr_b200", "sdrgen.py"))
    import oracle_py

    orc = oracle_py.Oracle()
    S, nb, bb = args.stations, args.blocks, 147000
    t0 = time.time()
    gens = [gen.StationGenerator(gen.Station.for_stream(k)) for k in range(S)]
    chunk = 100
    with capi.Chain(0, "r", n_streams=S) as ch:
        ch.set_overlap(True)
        chains = [orc.lib.orc_chain_create(0, ord("r"), 1) for _ in range(S)]
        bad_pcm = [0] * S
        bad_bits = [0] * S
        ngroups = [0] * S
        ngroups_ref = [0] * S
        import ctypes as C
        pcm_ref = np.zeros(2940, np.int16)
        for c0 in range(0, nb, chunk):
            n = min(chunk, nb - c0)
            iq = [gg.read(73500 * n) for gg in gens]
            for b in range(n):
                buf = np.stack([iq[s][b * bb:(b + 1) * bb] for s in range(S)])
                ch.process_host(buf)
                pcm = ch.read_pcm()
                rec = ch.read_rds()
                for s in range(S):
                    orc.lib.orc_chain_block(chains[s], np.ascontiguousarray(buf[s]), pcm_ref)
                    bad_pcm[s] += int((pcm[s] != pcm_ref).sum())
                    off, ns, nbt = C.c_int(0), C.c_int(0), C.c_int(0)
                    sp, bp = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
                    orc.lib.orc_chain_rds_block(chains[s], C.byref(off), C.byref(sp), C.byref(ns), C.byref(bp), C.byref(nbt))
                    ref_bits = np.ctypeslib.as_array(bp, shape=(nbt.value,)) if nbt.value else np.zeros(0, np.int32)
                    got_bits = rec[s]["bits"][: rec[s]["n_bits"]]
                    if nbt.value != int(rec[s]["n_bits"]) or not np.array_equal(ref_bits, got_bits) or off.value != int(rec[s]["cdr_offset"]):
                        bad_bits[s] += 1
                    gp = C.POINTER(C.c_uint64)()
                    ng = orc.lib.orc_chain_groups(chains[s], C.byref(gp))
                    ngroups_ref[s] += ng
                    ngroups[s] += int(rec[s]["n_groups"])
                    if ng != int(rec[s]["n_groups"]) or any(gp[i] != int(rec[s]["groups"][i]) for i in range(ng)):
                        bad_bits[s] += 1
            print(f"# {c0 + n} blocks, {time.time() - t0:.0f} s, differing pcm samples so far {bad_pcm}", file=sys.stderr, flush=True)
    for s in range(S):
        print(json.dumps({"station": s, "blocks": nb, "seconds_of_signal": round(nb * 0.030625, 1), "pcm_samples_differing": bad_pcm[s],
                          "blocks_with_rds_differences": bad_bits[s], "groups": ngroups[s], "groups_oracle": ngroups_ref[s]}))
    return 0 if not any(bad_pcm) and not any(bad_bits) else 1


if __name__ == "__main__":
    sys.exit(main())
