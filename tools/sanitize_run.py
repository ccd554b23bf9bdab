"""Workload for tools/sanitize.sh: a few blocks of one configuration through the C ABI, checked against the oracle.

    python tools/sanitize_run.py <mode> <m|s|r> <streams> <blocks> <overlap 0|1>
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g  # noqa: E402


def main():
    mode, kind, S, nblocks, overlap = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    capi = g._load("sdrb_capi", os.path.join(ROOT, "real-time-sdr_b200", "capi.py"))
    gen = g._load("sdrgen", os.path.join(ROOT, "real-time-sdr_b200", "sdrgen.py"))
    import oracle_py
    iq = gen.generate_iq(gen.Station.for_stream(0, fs=gen.mode_fs(mode)), gen.block_pairs(mode) * nblocks)
    want = oracle_py.Oracle().chain(mode, kind, iq)
    with capi.Chain(mode, kind, n_streams=S) as ch:
        ch.set_overlap(bool(overlap))
        bb = ch.info.block_bytes
        pcm = []
        for b in range(nblocks):
            ch.process_host(np.ascontiguousarray(np.stack([iq[b * bb:(b + 1) * bb]] * S)))
            pcm.append(ch.read_pcm()[S - 1].copy())
        ch.sync()
    ok = np.array_equal(np.concatenate(pcm), want["pcm"])
    print(f"sanitize_run mode {mode} {kind} S={S} blocks={nblocks} overlap={overlap}: parity {'ok' if ok else 'FAILED'}")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
