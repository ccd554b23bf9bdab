"""real-time-sdr_b200/csrc/res_lanes.h: the lane tables of the 247/640 RDS resampler in k_rds_backend (two adjacent output
residues per lane sharing one input sample per loop step, lags chosen for conflict-free shared-memory banks).  The host
build replays the kernel's per-thread loop on the tables the chain uploads; the result must equal the reference's
convolveFIR(y, x, h, state, 247, 640) (/root/reference/src/filter.cpp:123-147, src/rds.cpp:61,130) bit for bit, every
output must be written exactly once, and no warp may have two lanes in one bank."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
UP, DOWN, TAPS = 247, 640, 101


@pytest.fixture(scope="module")
def host(_built):
    L = C.CDLL(os.path.join(ROOT, "build", "libpllmath_host.so"))
    L.crh_resampler_lanes.argtypes = [f32p, f32p, C.c_int, C.c_int, f32p, i32p]
    L.crh_resampler_lanes.restype = C.c_int
    return L


@pytest.fixture(scope="module")
def lpf(oracle):
    return oracle.design("lpf_gain", TAPS * UP, Fs=240000.0 * UP, a=3e3, u=UP)  # src/rds.cpp:61


def replay(host, lpf, x_with_state, n_in):
    n_out = n_in * UP // DOWN
    y = np.full(n_out, np.nan, np.float32)
    info = np.zeros(4, np.int32)
    assert host.crh_resampler_lanes(lpf, np.ascontiguousarray(x_with_state, np.float32), n_in, n_out, y, info) == 0
    return y, info


@pytest.mark.parametrize("n_in", [7350, 7349, 6400, 641, 100])
def test_lane_replay_equals_reference_resampler(host, oracle, lpf, n_in):
    rng = np.random.default_rng(n_in)
    nblocks = 3
    x = (rng.standard_normal(n_in * nblocks) * rng.choice([1e-3, 1.0, 50.0], n_in * nblocks)).astype(np.float32)
    want = oracle.fir_updown(x, lpf, UP, DOWN, nblocks=nblocks, nstate=100)
    n_out = n_in * UP // DOWN
    assert want.size == nblocks * n_out
    state = np.zeros(100, np.float32)
    for b in range(nblocks):
        blk = x[b * n_in:(b + 1) * n_in]
        y, info = replay(host, lpf, np.concatenate([state, blk]), n_in)
        assert info.tolist() == [0, 1, 0, 0], info  # every pair placed, one lane per bank and warp, each output written once
        np.testing.assert_array_equal(y.view(np.uint32), want[b * n_out:(b + 1) * n_out].view(np.uint32))
        state = np.concatenate([state, blk])[-100:]


def test_nothing_outside_a_residues_own_window_reaches_its_output(host, oracle, lpf):
    """inf at the samples just outside an output's 101-sample window, and NaN everywhere the kernel may load beyond the
    block (the replay fills that with NaN itself): a zero tap instead of a skipped MAC would turn these into NaN."""
    n_in = 7350
    rng = np.random.default_rng(5)
    x = rng.standard_normal(n_in).astype(np.float32)
    x[rng.integers(0, n_in, 40)] = np.inf
    x[rng.integers(0, n_in, 40)] = -np.inf
    want = oracle.fir_updown(x, lpf, UP, DOWN, nblocks=1, nstate=100)
    y, info = replay(host, lpf, np.concatenate([np.zeros(100, np.float32), x]), n_in)
    assert info.tolist() == [0, 1, 0, 0]
    np.testing.assert_array_equal(y.view(np.uint32), want.view(np.uint32))
    assert np.isfinite(want).any() and (~np.isfinite(want)).any()
