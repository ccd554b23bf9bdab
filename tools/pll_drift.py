"""k_pll's time and careful-path rate as a function of how long the chain has been running.

The NCO phase of fmpll is a float that grows without bound (0.497 rad per sample for the 19 kHz loop, 2.98 for the
114 kHz loop), so the kernel's fast path must hold hours into a run, not only for the first seconds a benchmark sees.
Both loops are placed at sample count n0 through the checkpoint interface (as tests/test_chain_gpu.py does), then a few
blocks are run: per n0 the mean k_pll time, cycles per sample and the number of 4-sample chunks that took the careful path.

    python tools/pll_drift.py [--streams 1024] [--blocks 6]
"""
from __future__ import annotations

import argparse
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def patch(ch, blob: bytes, S: int, n0: float) -> bytes:
    offs = (ch.state_item_offset("pll19"), ch.state_item_offset("pll114"))
    out = bytearray(blob)
    for which, freq in ((0, 19e3), (1, 114e3)):
        for s in range(S):
            o = offs[which] + s * 24
            fbI, fbQ, integ, phase, trig = struct.unpack_from("<4fd", out, o)
            th = np.float32(2 * np.pi * float(np.float32(freq) / np.float32(240000.0)) * (n0 + 13 * s) + phase)
            struct.pack_into("<4fd", out, o, float(np.float32(np.cos(np.float64(th)))), float(np.float32(np.sin(np.float64(th)))),
                             integ, phase, float(n0 + 13 * s))
    return bytes(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--blocks", type=int, default=6)
    args = ap.parse_args()
    import torch
    capi = g._load("sdrb_capi", os.path.join(ROOT, "real-time-sdr_b200", "capi.py"))
    gen = g._load("sdrgen", os.path.join(ROOT, "real-time-sdr_b200", "sdrgen.py"))
    S = args.streams
    bp = gen.block_pairs(0)
    iq1 = gen.generate_iq(gen.Station(), bp * args.blocks).reshape(args.blocks, 2 * bp)
    pitch = (2 * bp + 255) // 256 * 256
    dev_in = []
    for b in range(args.blocks):
        t = torch.full((S, pitch), 128, dtype=torch.uint8, device="cuda")
        t[:, : 2 * bp] = torch.from_numpy(iq1[b]).cuda()[None, :]
        dev_in.append(t)
    torch.cuda.synchronize()
    rows = []
    for n0 in (0.0, 1e5, 1e6, 4e6, 1.7e7, 3.4e7, 6.8e7, 1.4e8, 2.8e8, 5.6e8, 9e8, 1.2e9, 2.5e9):
        with capi.Chain(0, "r", n_streams=S) as ch:
            ch.state_load(patch(ch, ch.state_save(), S, n0))
            ch.process_device(dev_in[0].data_ptr(), pitch)  # the first block after a load takes the careful path once per lane
            ch.sync()
            r0 = ch.pll_redos()
            ch.set_profiling(True)
            for b in range(1, args.blocks):
                ch.process_device(dev_in[b].data_ptr(), pitch)
            ch.sync()
            ms = ch.kernel_times().get("pll", 0.0)
            r1 = ch.pll_redos()
            chunks = S * (args.blocks - 1) * (7350 // 4)
            det = ch.pll_redo_detail()
            names = ["in_range", "wrap", "ambig_e", "base_near_2", "binade", "r_tiny", "ambig_sa", "ambig_cr", "generic_next"]
            detail = {nm: (det[2 * (1 + i)], det[2 * (1 + i) + 1]) for i, nm in enumerate(names) if det[2 * (1 + i)] or det[2 * (1 + i) + 1]}
            rows.append({"n0": n0, "tests_failed_19_114": detail, "phase19": 0.4974 * n0, "phase114": 2.9845 * n0, "pll_ms": round(ms, 4),
                         "cycles_per_sample": round(ms * 1e-3 / 7350 * 1.965e9, 1),
                         "redo_rate19": (r1[0] - r0[0]) / chunks, "redo_rate114": (r1[1] - r0[1]) / chunks})
            print(json.dumps(rows[-1]), flush=True)


if __name__ == "__main__":
    main()
