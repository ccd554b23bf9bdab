"""The synthetic station generator is part of the measurement contract: it must be deterministic, chunking
independent, and carry what it claims (pilot, stereo difference, RDS groups with valid check words)."""
from __future__ import annotations

import zlib

import numpy as np


def test_deterministic_and_chunking_independent(sdrgen):
    st = sdrgen.Station()
    a = sdrgen.generate_iq(st, 200000)
    g = sdrgen.StationGenerator(st)
    b = np.concatenate([g.read(n) for n in (1, 73500, 60000, 66499)])
    assert np.array_equal(a, b)
    assert a.dtype == np.uint8 and a.size == 400000
    assert 100 < a.mean() < 156 and a.min() >= 0 and a.max() <= 255


def test_streams_differ(sdrgen):
    a = sdrgen.generate_iq(sdrgen.Station.for_stream(1), 20000)
    b = sdrgen.generate_iq(sdrgen.Station.for_stream(2), 20000)
    assert zlib.crc32(a.tobytes()) != zlib.crc32(b.tobytes())


def test_rds_checkwords_have_the_standard_syndromes(sdrgen, oracle):
    bits = sdrgen.rds_bitstream(0x1234, 5, "B200-SDR", 4).astype(np.int32)
    kinds = [oracle.lib.orc_block_offset(np.ascontiguousarray(bits[i:i + 26])) for i in range(0, bits.size, 26)]
    assert kinds == [0, 1, 2, 4] * 4  # A, B, C, D


def test_block_sizes(sdrgen):
    assert [sdrgen.block_pairs(m) for m in range(4)] == [73500, 52920, 80000, 38400]
