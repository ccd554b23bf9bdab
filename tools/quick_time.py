"""Per-kernel device times of one block for a batch of streams (random IQ resident on the device).

    python tools/quick_time.py [--streams 1024] [--blocks 12] [--kind r]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=1024)
    ap.add_argument("--blocks", type=int, default=12)
    ap.add_argument("--kind", default="r")
    ap.add_argument("--mode", type=int, default=0)
    args = ap.parse_args()
    import torch
    capi = g._load("sdrb_capi", os.path.join(ROOT, "real-time-sdr_b200", "capi.py"))
    with capi.Chain(args.mode, args.kind, n_streams=args.streams) as ch:
        bb = ch.info.block_bytes
        pitch = (bb + 255) // 256 * 256
        iq = torch.randint(0, 256, (args.streams, pitch), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ch.set_profiling(True)
        rows = []
        for b in range(args.blocks):
            ch.process_device(iq.data_ptr(), pitch)
            ch.sync()
            rows.append(ch.kernel_times())
        ch.set_profiling(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        for b in range(args.blocks):
            ch.process_device(iq.data_ptr(), pitch)
        ch.sync()
        wall = (time.perf_counter() - t0) / args.blocks * 1e3
    last = rows[-1]
    tot = sum(last.values())
    print(json.dumps({"streams": args.streams, "kind": args.kind, "mode": args.mode, "kernel_ms": {k: round(v, 4) for k, v in last.items()},
                      "sum_ms": round(tot, 4), "wall_ms_per_block_unprofiled": round(wall, 4),
                      "MS_per_s": round(args.streams * ch.info.block_pairs / (wall * 1e-3) / 1e6, 1)}))
    print(json.dumps({"first_block_kernel_ms": {k: round(v, 4) for k, v in rows[0].items()}}))


if __name__ == "__main__":
    main()
