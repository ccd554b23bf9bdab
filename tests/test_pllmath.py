"""real-time-sdr_b200/csrc/pllmath.cuh (host build): the correctly-rounded float sin/cos/atan2 used by the
PLL kernels must equal (float)glibc_double_fn((double)x), which is what the reference computes
(/root/reference/src/pll.cpp:39,49,50,52).  The device build runs the same IEEE operations."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def host(_built):
    L = C.CDLL(os.path.join(ROOT, "build", "libpllmath_host.so"))
    L.crh_sincos.argtypes = [f32p, C.c_int, f32p, f32p]
    L.crh_cos.argtypes = [f32p, C.c_int, f32p]
    L.crh_atan2.argtypes = [f32p, f32p, C.c_int, f32p]
    L.crh_sincos_tier.argtypes = [f32p, C.c_int, C.c_int, f32p, f32p]
    L.crh_atan2_tier.argtypes = [f32p, f32p, C.c_int, C.c_int, f32p]
    L.crh_scan_sincos.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_uint64)]
    L.crh_scan_atan2.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.crh_pll.argtypes = [f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, f32p, C.POINTER(C.c_double)]
    return L


def test_sincos_scan_against_glibc(host):
    """Every 257th float bit pattern from 2^-40 up to 3e9 (plus the negative mirror via symmetry of the code path)."""
    out = (C.c_uint64 * 6)()
    host.crh_scan_sincos(0x2B800000, 0x4F32D05E, 257, 8, out)
    assert out[5] > 2_000_000
    assert out[0] == 0 and out[1] == 0, f"sin/cos mismatches vs glibc: {out[0]}/{out[1]}, first pattern {out[3]:#x}"
    assert out[2] < out[5] * 1e-4, "the slow tier should be rare"


def test_sincos_scan_above_3e9_against_glibc(host):
    """2^31.5 .. 2^45 rad: what the NCO phase reaches between 70 minutes and 1.8 years of a stream.  Every 97th float; the
    fast tier's quadrant count leaves 53-bit exactness there, the tiers' error bounds scale with it (reduce_rmin)."""
    out = (C.c_uint64 * 6)()
    host.crh_scan_sincos(0x4F32D05E, 0x56000000, 97, 8, out)  # 3e9 .. 2^45
    assert out[5] > 1_000_000
    assert out[0] == 0 and out[1] == 0, f"sin/cos mismatches vs glibc: {out[0]}/{out[1]}, first pattern {out[3]:#x}"
    assert out[2] < out[5] * 2e-3, "the slow tier should stay rare"


def test_sincos_dense_binades_above_3e9(host):
    """Every float of two whole binades far out (2^36 = one day of the 114 kHz loop, and the last one before the limit)."""
    for lo in (0x51800000, 0x55800000):  # [2^36, 2^37), [2^44, 2^45)
        out = (C.c_uint64 * 6)()
        host.crh_scan_sincos(lo, lo + 0x00800000, 1, 8, out)
        assert out[5] == 1 << 23
        assert out[0] == 0 and out[1] == 0, f"binade {lo:#x}: {out[0]}/{out[1]} mismatches, first {out[3]:#x}"


def test_sincos_dense_in_nco_range(host):
    """Dense scan of the range the 19 kHz NCO phase lives in during the first seconds."""
    out = (C.c_uint64 * 6)()
    host.crh_scan_sincos(0x47000000, 0x47400000, 1, 8, out)  # 32768 .. 49152, every float
    assert out[0] == 0 and out[1] == 0


def test_sincos_special_values(host):
    t = np.array([0.0, -0.0, 1e-30, -1e-30, 1.5707964, 3.1415927, 6.2831855, 1e-45, 2.5e9, -2.5e9, 4e9, -7.4e9, 3.3e13, 3.6e13, 1e20, -3e38,
                  np.inf, np.nan], np.float32)
    s, c = np.zeros_like(t), np.zeros_like(t)
    host.crh_sincos(t, t.size, s, c)
    with np.errstate(invalid="ignore"):
        ws, wc = np.sin(t.astype(np.float64)).astype(np.float32), np.cos(t.astype(np.float64)).astype(np.float32)
    assert np.array_equal(s.view(np.uint32)[:-2], ws.view(np.uint32)[:-2])
    assert np.array_equal(c.view(np.uint32)[:-2], wc.view(np.uint32)[:-2])
    assert np.isnan(s[-1]) and np.isnan(s[-2]) and np.isnan(c[-1])


@pytest.mark.parametrize("mode", [0, 1])
def test_atan2_random_against_glibc(host, mode):
    out = (C.c_uint64 * 6)()
    host.crh_scan_atan2(12345 + mode, 1_500_000, mode, 8, out)
    assert out[2] == 12_000_000
    assert out[0] == 0, f"{out[0]} atan2 mismatches vs glibc, first y={out[3]:#x} x={out[4]:#x}"


def test_atan2_special_values(host):
    y = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 1.0, 1e-38, 1e-45, 3e38, 1.0, 0.5], np.float32)
    x = np.array([1.0, 1.0, -1.0, -1.0, 0.0, 0.0, -0.0, 1e38, 1.0, 1e-38, 1.0, -0.5], np.float32)
    o = np.zeros_like(y)
    host.crh_atan2(y, x, y.size, o)
    w = np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32)
    assert np.array_equal(o.view(np.uint32), w.view(np.uint32)), (o, w)


def test_slow_tier_alone_is_also_exact(host):
    rng = np.random.default_rng(9)
    t = (rng.random(200000) * 3e6).astype(np.float32)
    s, c = np.zeros_like(t), np.zeros_like(t)
    host.crh_sincos_tier(t, t.size, 1, s, c)
    assert np.array_equal(s, np.sin(t.astype(np.float64)).astype(np.float32))
    assert np.array_equal(c, np.cos(t.astype(np.float64)).astype(np.float32))
    y = (rng.standard_normal(200000) * 0.01).astype(np.float32)
    x = (rng.standard_normal(200000) * 0.01).astype(np.float32)
    o = np.zeros_like(y)
    host.crh_atan2_tier(y, x, y.size, 1, o)
    assert np.array_equal(o, np.arctan2(y.astype(np.float64), x.astype(np.float64)).astype(np.float32))


@pytest.mark.parametrize("name,outn,freq,scale,bw", [("pilot", "carrier", 19e3, 2.0, 0.01), ("gen_pilot", "IPLL", 114e3, 0.5, 0.001)])
def test_pll_recurrence_equals_oracle(host, oracle, station_iq, name, outn, freq, scale, bw):
    """The PLL built from pll_step (what the CUDA kernels run) against the oracle's fmpll on real chain signals."""
    nblocks, n = 60, 7350
    a = oracle.chain(0, "r", station_iq(0, 0, nblocks), stages=(name, outn))
    x = a[name]
    out = np.zeros(n + 1, np.float32)
    out[n] = 1
    st = np.array([1, 0, 0, 0], np.float32)
    trig = C.c_double(0)
    ys = []
    for b in range(nblocks):
        host.crh_pll(x[b * n:(b + 1) * n], n, freq, 240000.0, scale, 0.0, bw, out, st, C.byref(trig))
        ys.append(out.copy())
    y = np.concatenate(ys)
    assert int((y.view(np.uint32) != a[outn].view(np.uint32)).sum()) == 0


def _run_fast(host, x, freq, scale, bw, n):
    host.crh_pll_fast.argtypes = [f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, f32p,
                                  C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    out = np.zeros(n + 1, np.float32)
    out[n] = 1
    st = np.array([1, 0, 0, 0], np.float32)
    trig = C.c_double(0)
    stats = (C.c_uint64 * 2)()
    ys = []
    for b in range(x.size // n):
        host.crh_pll_fast(np.ascontiguousarray(x[b * n:(b + 1) * n]), n, freq, 240000.0, scale, 0.0, bw, out, st, C.byref(trig), stats)
        ys.append(out.copy())
    return np.concatenate(ys), list(stats)


@pytest.mark.parametrize("name,outn,freq,scale,bw", [("pilot", "carrier", 19e3, 2.0, 0.01), ("gen_pilot", "IPLL", 114e3, 0.5, 0.001)])
def test_fast_recurrence_equals_oracle_on_chain_signals(host, oracle, station_iq, name, outn, freq, scale, bw):
    """pll_step_fast (rotated phase detector, what k_pll runs) on the chain's own PLL inputs, incl. state reload per block."""
    nblocks, n = 60, 7350
    a = oracle.chain(0, "r", station_iq(0, 0, nblocks), stages=(name, outn))
    y, stats = _run_fast(host, a[name], freq, scale, bw, n)
    assert int((y.view(np.uint32) != a[outn].view(np.uint32)).sum()) == 0
    assert stats[0] < 100 and stats[1] < 100, "the general path must stay rare on real signals"


def test_fast_recurrence_edge_inputs(host, oracle):
    rng = np.random.default_rng(0)
    n, nb = 2000, 8
    cases = {
        "noise": (rng.standard_normal(n * nb) * 0.02).astype(np.float32),
        "zeros": np.zeros(n * nb, np.float32),
        "subnormal": (rng.standard_normal(n * nb) * 1e-38).astype(np.float32),
        "huge": (rng.standard_normal(n * nb) * 1e30).astype(np.float32),
        "sparse": np.where(rng.random(n * nb) < 0.3, 0, rng.standard_normal(n * nb)).astype(np.float32),
        "negated_tone": (-0.05 * np.cos(2 * np.pi * 19000 / 240000 * np.arange(n * nb))).astype(np.float32),
    }
    for nm, x in cases.items():
        for freq, scale, bw in ((19e3, 2.0, 0.01), (114e3, 0.5, 0.001)):
            want, _ = oracle.pll(x, freq, 240000.0, scale, 0.0, bw, nblocks=nb)
            y, _ = _run_fast(host, x, freq, scale, bw, n)
            assert int((y.view(np.uint32) != want.view(np.uint32)).sum()) == 0, (nm, freq)


def test_lean_cosine_equals_glibc(host):
    """cos_lean_f (the NCO-output kernel's cosine) on 4M arguments over the range the scaled NCO phase visits."""
    host.crh_cos_lean.argtypes = [f32p, C.c_int, f32p]
    rng = np.random.default_rng(21)
    t = np.concatenate([(rng.random(2_000_000) * 6e6).astype(np.float32), (rng.random(1_000_000) * 100).astype(np.float32),
                        -(rng.random(500_000) * 1e4).astype(np.float32), (rng.random(500_000) * 2.9e9).astype(np.float32),
                        (rng.random(500_000) * 3.4e13).astype(np.float32), -(rng.random(200_000) * 1e11).astype(np.float32),
                        np.array([0.0, -0.0, 1e-30, 3.1415927, 1.5707964, 4e9, 1e20], np.float32)])
    c = np.zeros_like(t)
    host.crh_cos_lean(t, t.size, c)
    assert np.array_equal(c.view(np.uint32), np.cos(t.astype(np.float64)).astype(np.float32).view(np.uint32))


@pytest.mark.parametrize("freq,scale,bw", [(19e3, 2.0, 0.01), (114e3, 0.5, 0.001)])
@pytest.mark.parametrize("n0", [2.0e6, 9.0e6, 4.0e7, 3.0e8, 1.3e9, 2.0e10, 3.0e12])
def test_fast_recurrence_at_large_phase(host, oracle, oracle_mod, freq, scale, bw, n0):
    """Hours into a run the NCO phase is a float with an ulp of radians (2^22 rad is reached after 35 s / 6 s).  Start both
    recurrences from the same state far down the road: the float grid is then coarser than a quadrant, which is exactly
    where a reduction that is only approximately consistent goes wrong without the loop visibly losing lock."""
    host.crh_pll_fast.argtypes = [f32p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, f32p, f32p,
                                  C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    n, nb = 7350, 3
    t = np.arange(n * nb, dtype=np.float64) + n0
    rng = np.random.default_rng(int(n0) % 1000)
    x = (0.05 * np.cos(2 * np.pi * (freq + 2.0) / 240000.0 * t + 0.4) + 0.004 * rng.standard_normal(t.size)).astype(np.float32)
    # oracle from a hand-made state at sample count n0 (feedback consistent with the phase, as fmpll leaves it)
    st = oracle_mod.PllState()
    oracle.lib.orc_pll_init(C.byref(st))
    st.trigOffset = float(n0)
    st.phaseEst = np.float32(0.37)
    st.integrator = np.float32(1e-4)
    th = np.float32(2 * np.pi * float(np.float32(freq) / np.float32(240000.0)) * n0 + float(st.phaseEst))
    st.feedbackI = np.float32(np.cos(np.float64(th)))
    st.feedbackQ = np.float32(np.sin(np.float64(th)))
    want = []
    out_o = np.zeros(n + 1, np.float32)
    out_o[n] = 1
    out_f = out_o.copy()
    st4 = np.array([st.feedbackI, st.feedbackQ, st.integrator, st.phaseEst], np.float32)
    trig = C.c_double(float(n0))
    stats = (C.c_uint64 * 2)()
    got = []
    for b in range(nb):
        xb = np.ascontiguousarray(x[b * n:(b + 1) * n])
        oracle.lib.orc_pll(xb, n, freq, 240000.0, out_o, C.byref(st), scale, 0.0, bw)
        host.crh_pll_fast(xb, n, freq, 240000.0, scale, 0.0, bw, out_f, st4, C.byref(trig), stats)
        want.append(out_o.copy())
        got.append(out_f.copy())
    want, got = np.concatenate(want), np.concatenate(got)
    assert int((want.view(np.uint32) != got.view(np.uint32)).sum()) == 0
    assert [float(st.feedbackI), float(st.feedbackQ), float(st.integrator), float(st.phaseEst)] == [float(v) for v in st4]
    assert st.trigOffset == trig.value
    # ... and it must stay on the fast path: out here the loop cannot follow its input any more (the phase grid is coarser
    # than pi), every quadrant / input-sign combination occurs, and each one has to be handled by the speculative step.
    # (Round 1 sent the "input sign opposite to the NCO's, quadrant 0 or 2" case to the careful path: right values, 6x the time.)
    chunks = nb * (n // 4)
    assert stats[0] < 0.005 * chunks, f"{stats[0]} of {chunks} chunks left the fast path"


def test_lean_kernels_error_bound(host):
    """The speculative PLL step evaluates sin / cos of the reduced phase with kernels one coefficient shorter than fdlibm's
    (pllmath.cuh: sincos_poly2_lean).  Its acceptance tests assume a relative error below 2^-45 (512-ulp tie window for the
    float roundings of sa / cr); the kernels must stay inside that with margin, over the whole reduced range."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 200
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    host.crh_poly_lean.argtypes = [f64p, C.c_int, f64p, f64p]
    rng = np.random.default_rng(5)
    r = np.concatenate([np.linspace(1e-9, np.pi / 4, 6001), rng.random(6000) * (np.pi / 4), np.array([np.pi / 4 * (1 + 2.0 ** -50), 2.0 ** -30, 2.0 ** -12])])
    r = np.concatenate([r, -r]).astype(np.float64)
    s = np.zeros_like(r)
    c = np.zeros_like(r)
    host.crh_poly_lean(r, r.size, s, c)
    worst_s = worst_c = mp.mpf(0)
    for ri, si, ci in zip(r, s, c):
        x = mp.mpf(float(ri))
        worst_s = max(worst_s, abs((mp.mpf(float(si)) - mp.sin(abs(x))) / mp.sin(abs(x))))  # the kernel returns sin |r|
        worst_c = max(worst_c, abs((mp.mpf(float(ci)) - mp.cos(x)) / mp.cos(x)))
    assert worst_s < mp.mpf(2) ** -46.3, float(mp.log(worst_s, 2))
    assert worst_c < mp.mpf(2) ** -50, float(mp.log(worst_c, 2))  # approximation 2^-51.5 plus the rounding of the evaluation


def test_phase_detector_error_stays_inside_its_tolerance(host):
    """The speculative step accepts errorD = RN_f(e) when no float rounding tie lies within 2^-42 sa + 2^-48 of e.  That is
    only sound if |e - atan2(errorQ, errorI)| stays below that tolerance: measured here against long-double atan2 on 4M
    random (NCO phase, input) pairs, small reduced arguments (where the tolerance shrinks with sa) included."""
    f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
    host.crh_head_error.argtypes = [f32p, f32p, C.c_int, f64p]
    rng = np.random.default_rng(11)
    n = 1_000_000
    k = rng.integers(0, 4000, n)
    thetas = [
        (rng.random(n) * 6e6).astype(np.float32),                                                   # anywhere
        (k * (np.pi / 2) + rng.standard_normal(n) * 1e-3).astype(np.float32),                        # near multiples of pi/2: small sa
        (k * (np.pi / 2) + 10.0 ** rng.uniform(-7, -1, n) * rng.choice([-1, 1], n)).astype(np.float32),
        (rng.random(n) * 3e9).astype(np.float32),                                                   # coarse float grid
    ]
    for th in thetas:
        x = (rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 1, n)).astype(np.float32)
        out = np.zeros(3)
        host.crh_head_error(np.ascontiguousarray(th), x, n, out)
        assert out[1] > 0.5 * n
        assert out[0] < 0.4, f"error reaches {out[0]:.2f} of the tolerance (max abs {out[2]:.3e})"


def test_discriminator_quotient_fast_path_equals_the_exact_division(host):
    """cr::fm_quotient_fast (the divide of k_rf_frontend's FM discriminator): whatever it accepts must be
    (float)((double)num / den) bit for bit (/root/reference/src/demod.cpp:11,17), for a reciprocal seed up to four float-ulps
    off either way (the device's rcp.approx is within one); it must accept nearly everything in the signal's range and
    leave the rest (ties, sub-/supernormal operands) to the exact division."""
    u64p = C.POINTER(C.c_uint64)
    host.crh_fm_quotient_scan.argtypes = [f32p, f32p, f32p, C.c_int, C.c_int, u64p]
    rng = np.random.default_rng(11)
    n = 2_000_000
    sets = []
    # the chain's own range: |I|, |Q| <= ~1.3, numerators of differences of neighbouring samples
    I = (rng.standard_normal(n) * 0.5).astype(np.float32)
    Q = (rng.standard_normal(n) * 0.5).astype(np.float32)
    sets.append((I * (rng.standard_normal(n) * 0.3).astype(np.float32), I, Q))
    # every binade the fast path may see, and beyond (it has to decline those)
    ex = rng.integers(-70, 70, n)
    I = np.ldexp(rng.uniform(1, 2, n), ex).astype(np.float32) * rng.choice([-1, 1], n).astype(np.float32)
    Q = np.ldexp(rng.uniform(1, 2, n), ex + rng.integers(-3, 3, n)).astype(np.float32)
    num = np.ldexp(rng.uniform(1, 2, n), rng.integers(-126, 127, n)).astype(np.float32) * rng.choice([-1, 1], n).astype(np.float32)
    sets.append((num, I, Q))
    # quotients that are exactly representable or exact ties in float: small integers over powers of two
    I = np.ldexp(1.0, rng.integers(-20, 20, n)).astype(np.float32)
    num = (rng.integers(-(1 << 24), 1 << 24, n).astype(np.float64) * 0.5).astype(np.float32)
    sets.append((num, I, np.zeros(n, np.float32)))
    for k, (num, I, Q) in enumerate(sets):
        for ulps in (0, 1, -1, 4, -4):
            out = (C.c_uint64 * 3)()
            host.crh_fm_quotient_scan(np.ascontiguousarray(num), np.ascontiguousarray(I), np.ascontiguousarray(Q), n, ulps, out)
            accepted, bad, outside = out[0], out[1], out[2]
            assert bad == 0, (k, ulps, accepted, bad)
            if k == 0:
                assert outside < n * 1e-3 and accepted > (n - outside) * 0.9995, (k, ulps, accepted, outside)
    # specials: zero numerator, infinities and NaN never come out of the fast path wrong (they are declined or exact)
    num = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, 1.0, 1e-45, 3e38], np.float32)
    I = np.array([1.0, 1.0, 1.0, 1.0, 1.0, np.inf, np.nan, 1.0, 1e-19], np.float32)
    Q = np.zeros_like(I)
    out = (C.c_uint64 * 3)()
    host.crh_fm_quotient_scan(num, I, Q, num.size, 0, out)
    assert out[1] == 0
