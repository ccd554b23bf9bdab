"""Decoded-bit streams for the RDS decoders' tests: clean 0A groups of the synthetic station, the same with bit errors
injected, bursts that make a locked decoder lose sync, and plain noise; cut into the ragged chunks (36/37 bits) the chain's
blocks deliver."""
from __future__ import annotations

import numpy as np


def chunks_of(bits: np.ndarray, sizes=(37, 36, 37, 37, 36)) -> list:
    out, pos, k = [], 0, 0
    while pos < bits.size:
        n = sizes[k % len(sizes)]
        out.append(bits[pos:pos + n].astype(np.int32))
        pos += n
        k += 1
    return out


def cases(sdrgen) -> dict:
    rng = np.random.default_rng(2024)
    clean = sdrgen.rds_bitstream(0x1234, 5, "B200-SDR", 40).astype(np.int32)          # 40 groups = 160 blocks = 4160 bits
    lead = rng.integers(0, 2, 61).astype(np.int32)                                    # the decoder starts mid-block
    c = {}
    c["clean"] = np.concatenate([lead, clean])
    e = c["clean"].copy()
    flips = rng.random(e.size) < 0.01                                                 # 1 % bit errors: some blocks fail, sync holds
    e[flips] ^= 1
    c["ber_1pct"] = e
    b = c["clean"].copy()
    b[1500:3400] = rng.integers(0, 2, 1900)                                           # a long burst: > 40 of 50 blocks bad -> "Lost Sync", then re-acquisition
    c["burst_lose_sync"] = b
    c["noise"] = rng.integers(0, 2, 3000).astype(np.int32)                            # false syndrome hits, no sync
    other = sdrgen.rds_bitstream(0xC27A, 9, "RADIO 42", 24).astype(np.int32)
    c["two_stations"] = np.concatenate([clean[:1700], other])                         # a phase slip: block boundaries move
    return c
