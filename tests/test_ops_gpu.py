"""The stand-alone batched primitives of the C ABI (sdrb_fir_decim, _fir_updown, _fm_demod, _pll, _cdr)
against the oracle's functions, bit-exact, with torch only providing device memory."""
from __future__ import annotations

import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    t = pytest.importorskip("torch")
    assert t.cuda.is_available(), "GPU tests need a CUDA device"
    return t


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_fir_decim_blocks(capi, oracle, torch):
    rng = np.random.default_rng(1)
    L = capi.lib()
    for nh, decim, nx, nblocks in ((101, 10, 3000, 3), (101, 1, 700, 4), (33, 4, 512, 2), (101, 5, 1000, 2)):
        h = (rng.standard_normal(nh) / nh).astype(np.float32)
        S = 3
        x = rng.standard_normal((S, nx * nblocks)).astype(np.float32)
        want = np.stack([oracle.fir_decim(x[s], h, decim, nblocks=nblocks, state=np.zeros(nh - 1, np.float32)) for s in range(S)])
        state = torch.zeros((S, nh - 1), dtype=torch.float32, device="cuda")
        ys = []
        for b in range(nblocks):
            xb = dev(torch, x[:, b * nx:(b + 1) * nx])
            y = torch.zeros((S, nx // decim), dtype=torch.float32, device="cuda")
            capi.check(L.sdrb_fir_decim(xb.data_ptr(), nx, nx, h, nh, state.data_ptr(), y.data_ptr(), nx // decim, decim, S, None))
            torch.cuda.synchronize()
            ys.append(y.cpu().numpy())
        got = np.concatenate(ys, axis=1)
        assert got.tobytes() == want.astype(np.float32).tobytes(), (nh, decim)


def test_fir_updown_blocks(capi, oracle, torch):
    rng = np.random.default_rng(2)
    L = capi.lib()
    for up, down, nx, nblocks in ((147, 800, 8000, 2), (247, 640, 7350, 2), (1, 5, 7350, 2), (3, 7, 500, 3)):
        nh = 101 * up
        h = (rng.standard_normal(nh) / 50).astype(np.float32)
        S = 2
        x = rng.standard_normal((S, nx * nblocks)).astype(np.float32)
        want = np.stack([oracle.fir_updown(x[s], h, up, down, nblocks=nblocks, nstate=100) for s in range(S)])
        state = torch.zeros((S, 100), dtype=torch.float32, device="cuda")
        ny = nx * up // down
        ys = []
        for b in range(nblocks):
            xb = dev(torch, x[:, b * nx:(b + 1) * nx])
            y = torch.zeros((S, ny), dtype=torch.float32, device="cuda")
            capi.check(L.sdrb_fir_updown(xb.data_ptr(), nx, nx, h, nh, state.data_ptr(), 100, y.data_ptr(), ny, up, down, S, None))
            torch.cuda.synchronize()
            ys.append(y.cpu().numpy())
        got = np.concatenate(ys, axis=1)
        assert got.tobytes() == want.tobytes(), (up, down)


def test_fm_demod(capi, oracle, torch):
    rng = np.random.default_rng(3)
    L = capi.lib()
    n, nblocks, S = 1000, 3, 2
    I = rng.standard_normal((S, n * nblocks)).astype(np.float32)
    Q = rng.standard_normal((S, n * nblocks)).astype(np.float32)
    I[0, 5] = Q[0, 5] = 0.0  # guarded branch
    prev = torch.zeros((S, 2), dtype=torch.float32, device="cuda")
    outs = []
    for b in range(nblocks):
        di, dq = dev(torch, I[:, b * n:(b + 1) * n]), dev(torch, Q[:, b * n:(b + 1) * n])
        o = torch.zeros((S, n), dtype=torch.float32, device="cuda")
        capi.check(L.sdrb_fm_demod(di.data_ptr(), dq.data_ptr(), n, n, prev.data_ptr(), o.data_ptr(), n, S, None))
        torch.cuda.synchronize()
        outs.append(o.cpu().numpy())
    got = np.concatenate(outs, axis=1)
    for s in range(S):
        want, _ = oracle.fmdemod(I[s], Q[s], nblocks=nblocks)
        assert got[s].tobytes() == want.tobytes()


@pytest.mark.parametrize("freq,scale,bw", [(19e3, 2.0, 0.01), (114e3, 0.5, 0.001)])
def test_pll_matches_glibc_path(capi, oracle, torch, freq, scale, bw):
    """40 lanes (ragged), 3 blocks of 7350: the NCO output is bit-identical to the oracle's fmpll."""
    L = capi.lib()
    n, nblocks, S = 7350, 3, 40
    t = np.arange(n * nblocks, dtype=np.float64)
    rng = np.random.default_rng(4)
    x = np.stack([(0.02 + 0.01 * s) * np.cos(2 * np.pi * (freq + 3.0 * s) / 240000.0 * t + 0.1 * s)
                  + 0.002 * rng.standard_normal(t.size) for s in range(S)]).astype(np.float32)
    st = np.zeros(S, capi.PLL_STATE_DTYPE)
    st["feedbackI"] = 1.0
    st["last_out"] = 1.0
    st["lastCarrier"] = 1.0
    dst = torch.from_numpy(st.view(np.uint8).reshape(S, -1).copy()).cuda()
    outs = []
    for b in range(nblocks):
        xb = dev(torch, x[:, b * n:(b + 1) * n])
        o = torch.zeros((S, n + 1), dtype=torch.float32, device="cuda")
        capi.check(L.sdrb_pll(xb.data_ptr(), n, n, freq, 240000.0, scale, 0.0, bw, dst.data_ptr(), o.data_ptr(), n + 1, S, None))
        torch.cuda.synchronize()
        outs.append(o.cpu().numpy())
    got = np.concatenate(outs, axis=1)
    for s in (0, 1, 17, 31, 32, 39):
        want, _ = oracle.pll(x[s], freq, 240000.0, scale, 0.0, bw, nblocks=nblocks)
        bad = int((got[s].view(np.uint32) != want.view(np.uint32)).sum())
        assert bad == 0, f"lane {s}: {bad} NCO samples differ"


def test_cdr(capi, oracle, torch):
    rng = np.random.default_rng(5)
    L = capi.lib()
    S, n = 5, 2836
    x = (rng.standard_normal((S, n)) * 2.5).astype(np.float32)
    x[1] *= 0.1  # every |x| < 1: all sums are 0 -> offset 0
    off = torch.zeros(S, dtype=torch.int32, device="cuda")
    capi.check(L.sdrb_cdr(dev(torch, x).data_ptr(), n, n, 39, off.data_ptr(), S, None))
    torch.cuda.synchronize()
    got = off.cpu().numpy()
    want = np.array([oracle.cdr(x[s], 39)[0] for s in range(S)])
    assert np.array_equal(got, want)
