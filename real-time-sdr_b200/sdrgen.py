"""Deterministic synthetic FM broadcast stations (stereo multiplex + RDS) as 8-bit interleaved IQ.

This is the workload generator for the parity tests and the benchmark.  The
reference ships no recordings (its model/*.py read ../data/samples*.raw, which is git-ignored there),
so every test input is synthesised:

    mpx(t) = 0.4 (L+R) + 0.1 cos(2 pi 19k t) + 0.4 (L-R) cos(2 pi 38k t) + 0.10 b(t) sin(2 pi 57k t)
    phi    = 2 pi 75e3 * cumsum(mpx) / fs
    I, Q   = clip(round(127.5 + 100 (cos phi, sin phi) + N(0,1)), 0, 255)   -> u8, interleaved I,Q,I,Q...

with L, R two audio tones and b(t) a biphase RDS baseband carrying 0A groups (PI, PTY, 8-char PS).
Pilot and 38 kHz subcarrier are both cosines: the reference regenerates the subcarrier as cos(2 theta)
of a PLL locked to the pilot (/root/reference/src/pll.cpp:52, src/stereo.cpp:77), so a cosine pair is
what yields L on even and R on odd PCM indices.

Everything is computed from exact integer phase arithmetic and +,-,*,/ on float64 only (own sin/cos
polynomials), so the byte stream does not depend on which libm/SIMD variant numpy was built with; the
only library randomness is numpy's PCG64 normal generator, seeded per (stream seed, chunk index).
"""
from __future__ import annotations

import dataclasses

import numpy as np

FS_DEFAULT = 2_400_000
_CHUNK = 1 << 16  # generator granularity (samples); noise is seeded per chunk

# RDS block check: g(x) = x^10+x^8+x^7+x^5+x^4+x^3+1, offset words A, B, C, D (IEC 62106).
_RDS_POLY = 0x5B9
_RDS_OFFSETS = (0x0FC, 0x198, 0x168, 0x1B4)


def _sincos_turns(frac: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """sin, cos of 2*pi*frac for frac in [0, 1), using only exactly-rounded float64 arithmetic."""
    x = np.asarray(frac, dtype=np.float64) * 8.0
    octant = np.floor(x + 0.5)
    r = (x - octant) * (np.pi / 4.0)  # |r| <= pi/8
    r2 = r * r
    # Taylor to r^15 / r^14: truncation < 1e-17 at pi/8
    s = r * (1.0 + r2 * (-1.0 / 6 + r2 * (1.0 / 120 + r2 * (-1.0 / 5040 + r2 * (1.0 / 362880 + r2 * (
        -1.0 / 39916800 + r2 * (1.0 / 6227020800 + r2 * (-1.0 / 1307674368000))))))))
    c = 1.0 + r2 * (-0.5 + r2 * (1.0 / 24 + r2 * (-1.0 / 720 + r2 * (1.0 / 40320 + r2 * (
        -1.0 / 3628800 + r2 * (1.0 / 479001600 + r2 * (-1.0 / 87178291200)))))))
    k = octant.astype(np.int64) & 7
    h = np.sqrt(0.5)
    # rotate by k * 45 degrees
    ck = np.array([1.0, h, 0.0, -h, -1.0, -h, 0.0, h])[k]
    sk = np.array([0.0, h, 1.0, h, 0.0, -h, -1.0, -h])[k]
    return s * ck + c * sk, c * ck - s * sk


def _tone_turns(freq_hz: int, n: np.ndarray, fs: int) -> np.ndarray:
    """Phase of a freq_hz tone at sample indices n, in turns, exact to one float64 division."""
    return ((n * freq_hz) % fs).astype(np.float64) / float(fs)


def rds_checkword(data16: int, offset: int) -> int:
    reg = 0
    for i in range(15, -1, -1):
        reg = (reg << 1) | ((data16 >> i) & 1)
        if reg & (1 << 10):
            reg ^= _RDS_POLY
    for _ in range(10):
        reg <<= 1
        if reg & (1 << 10):
            reg ^= _RDS_POLY
    return (reg & 0x3FF) ^ offset


def rds_group_0a(pi: int, pty: int, ps: str, segment: int, block_c: int = 0xE0CD) -> list[int]:
    """The four 16-bit data words of a type 0A group carrying PS characters 2*segment, 2*segment+1."""
    ps8 = (ps + " " * 8)[:8].encode("latin-1")
    b = (0 << 12) | (0 << 11) | (0 << 10) | ((pty & 0x1F) << 5) | (segment & 3)
    d = (ps8[2 * segment] << 8) | ps8[2 * segment + 1]
    return [pi & 0xFFFF, b, block_c & 0xFFFF, d]


def rds_bitstream(pi: int, pty: int, ps: str, n_groups: int) -> np.ndarray:
    """Source bits (before differential encoding) of n_groups consecutive 0A groups, 104 bits each."""
    bits = np.empty(n_groups * 104, dtype=np.int8)
    pos = 0
    for g in range(n_groups):
        words = rds_group_0a(pi, pty, ps, g & 3)
        for w, off in zip(words, _RDS_OFFSETS):
            blk = (w << 10) | rds_checkword(w, off)
            for i in range(25, -1, -1):
                bits[pos] = (blk >> i) & 1
                pos += 1
    return bits


@dataclasses.dataclass(frozen=True)
class Station:
    seed: int = 1234
    pi: int = 0x1234
    pty: int = 5
    ps: str = "B200-SDR"
    f_left: int = 1000
    f_right: int = 2500
    rds_level: float = 0.10
    noise: float = 1.0
    fs: int = FS_DEFAULT

    @staticmethod
    def for_stream(k: int, fs: int = FS_DEFAULT) -> "Station":
        """Station k of a batch: distinguishable tones, PI and PS."""
        if k == 0:
            return Station(fs=fs)
        return Station(seed=1234 + k, pi=(0x1234 + k) & 0xFFFF, ps="STN%05d" % (k % 100000),
                       f_left=300 + (700 + 7 * k) % 2700, f_right=300 + (2200 + 11 * k) % 2700, fs=fs)


class StationGenerator:
    """Sequential u8 IQ generator for one station.  read(n_pairs) returns 2*n_pairs bytes."""

    _HALF_RATE = 2375  # biphase half-symbols per second (1187.5 bit/s)
    _SPAN = 4          # pulses summed on each side of a sample

    def __init__(self, station: Station):
        self.st = station
        self.fs = int(station.fs)
        self._n = 0                # absolute index of the next sample to emit
        self._phase_turns = 0.0    # FM phase carried between chunks, wrapped to [0,1)
        self._buf = np.empty(0, dtype=np.uint8)
        self._chunk_index = 0
        # differentially encoded, biphase half-symbol sequence d_j = +-1, periodic over 4 groups
        src = rds_bitstream(station.pi, station.pty, station.ps, 4)
        self._src_bits = src
        self._half_cache: dict[int, np.ndarray] = {}

    # --- RDS baseband -------------------------------------------------------------------------
    def _half_symbols(self, j_lo: int, j_hi: int) -> np.ndarray:
        """d_j for j in [j_lo, j_hi): differential encoding restarts from 0 at bit 0 (j >= 0)."""
        out = np.zeros(j_hi - j_lo, dtype=np.float64)
        period = self._src_bits.size  # 416 source bits; XOR-sum over a period is fixed
        per_par = int(self._src_bits.sum() & 1)
        csum = np.cumsum(self._src_bits) & 1  # parity of bits 0..i
        for idx, j in enumerate(range(j_lo, j_hi)):
            if j < 0:
                continue
            bit_i = j >> 1
            full, rem = divmod(bit_i, period)
            enc = ((full * per_par) + int(csum[rem])) & 1  # differentially encoded bit value
            first = 1.0 if enc else -1.0                   # 1 -> (+1,-1), 0 -> (-1,+1)
            out[idx] = first if (j & 1) == 0 else -first
        return out

    def _rds_baseband(self, n: np.ndarray) -> np.ndarray:
        """b(t): half-symbol impulses d_j shaped by the inverse transform of cos(pi f / 4800), |f| < 2400 Hz:
        p(tau) = cos(2 pi 2400 tau) * 2a / (a^2 - (2 pi tau)^2), a = pi/4800, scaled to about unit peak."""
        fs, hr = self.fs, self._HALF_RATE
        j0 = (n * hr) // fs  # index of the half-symbol at or before each sample
        j_lo, j_hi = int(j0[0]) - self._SPAN + 1, int(j0[-1]) + self._SPAN + 1
        d = self._half_symbols(j_lo, j_hi)
        # cos(2 pi 2400 (t - j/hr)) = cos(A_n - B_j): one sincos per sample, one per half-symbol
        sa, ca = _sincos_turns(_tone_turns(2400, n, fs))
        jj = np.arange(j_lo, j_hi, dtype=np.int64)
        sb, cb = _sincos_turns(((jj * 2400) % hr).astype(np.float64) / float(hr))
        dcb, dsb = d * cb, d * sb
        acc = np.zeros(n.size, dtype=np.float64)
        a = np.pi / 4800.0
        sing = fs * hr // 9600  # |numerator| at which the pulse formula is 0/0 (limit: 2400)
        for w in range(-self._SPAN + 1, self._SPAN + 1):
            j = j0 + w
            num = n * hr - j * fs  # tau = num / (fs * hr), exact integers
            u = num.astype(np.float64) * (2.0 * np.pi / float(fs * hr))
            safe = np.abs(num) != sing
            k = j - j_lo
            cosd = ca * dcb[k] + sa * dsb[k]  # d_j * cos(A - B_j)
            pulse = np.where(safe, cosd * (2.0 * a) / np.where(safe, a * a - u * u, 1.0), d[k] * 2400.0)
            acc += pulse
        return acc * (a / 2.0) * (1.0 / 1.13)  # isolated pulse peak 2/a -> 1; summed peak 1.13 -> 1

    # --- one chunk ----------------------------------------------------------------------------
    def _make_chunk(self) -> np.ndarray:
        st, fs = self.st, self.fs
        n = np.arange(self._n, self._n + _CHUNK, dtype=np.int64)
        sl, _ = _sincos_turns(_tone_turns(st.f_left, n, fs))
        sr, _ = _sincos_turns(_tone_turns(st.f_right, n, fs))
        left, right = 0.5 * sl, 0.5 * sr
        s19, c19 = _sincos_turns(_tone_turns(19000, n, fs))
        c38 = 2.0 * c19 * c19 - 1.0               # cos 2x
        s57 = s19 * (3.0 - 4.0 * s19 * s19)       # sin 3x
        mpx = 0.4 * (left + right) + 0.1 * c19 + 0.4 * (left - right) * c38
        if st.rds_level:
            mpx = mpx + st.rds_level * self._rds_baseband(n) * s57
        turns = self._phase_turns + np.cumsum(mpx) * (75e3 / fs)
        self._phase_turns = float(turns[-1] - np.floor(turns[-1]))
        s, c = _sincos_turns(turns - np.floor(turns))
        rng = np.random.Generator(np.random.PCG64([st.seed, self._chunk_index]))
        noise = rng.standard_normal((2, _CHUNK)) * st.noise
        iq = np.empty(2 * _CHUNK, dtype=np.uint8)
        iq[0::2] = np.clip(np.rint(127.5 + 100.0 * c + noise[0]), 0, 255).astype(np.uint8)
        iq[1::2] = np.clip(np.rint(127.5 + 100.0 * s + noise[1]), 0, 255).astype(np.uint8)
        self._n += _CHUNK
        self._chunk_index += 1
        return iq

    def read(self, n_pairs: int) -> np.ndarray:
        need = 2 * n_pairs
        parts = [self._buf]
        have = self._buf.size
        while have < need:
            c = self._make_chunk()
            parts.append(c)
            have += c.size
        buf = np.concatenate(parts) if len(parts) > 1 else parts[0]
        self._buf = buf[need:]
        return np.ascontiguousarray(buf[:need])


def generate_iq(station: Station, n_pairs: int) -> np.ndarray:
    """2*n_pairs bytes of interleaved u8 IQ from the start of the station's signal."""
    return StationGenerator(station).read(n_pairs)


# Reference mode table (/root/reference/src/project.cpp:31-44,67-108): IQ pairs per block.
def block_pairs(mode: int) -> int:
    rf_decim, down, up = {0: (10, 5, 1), 1: (4, 9, 1), 2: (10, 800, 147), 3: (3, 1280, 147)}[mode]
    return (1470 * rf_decim * down) // up


def mode_fs(mode: int) -> int:
    return {0: 2_400_000, 1: 1_440_000, 2: 2_400_000, 3: 1_152_000}[mode]
