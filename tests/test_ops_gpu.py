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


def _run_bits_gpu(capi, torch, sym_blocks, block0):
    """sdrb_manchester_decode + sdrb_differential_decode over consecutive blocks for a batch of streams.
    sym_blocks[b][s] = symbols of stream s in block b.  Returns per stream (manchester, decoded, lens, state)."""
    L = capi.lib()
    S = len(sym_blocks[0])
    pitch = max(max(len(x) for x in blk) for blk in sym_blocks) + 1
    mst = torch.zeros(S * 2, dtype=torch.int32, device="cuda")
    last = torch.zeros(S, dtype=torch.int32, device="cuda")
    man = [[] for _ in range(S)]
    dec = [[] for _ in range(S)]
    lens = [[] for _ in range(S)]
    bc = block0
    for blk in sym_blocks:
        sym = np.full((S, pitch), 7, np.int32)  # 7: a value a correct kernel never reads as a symbol
        for s, x in enumerate(blk):
            sym[s, :len(x)] = x
        nsym = dev(torch, np.array([len(x) for x in blk], np.int32))
        d_sym = dev(torch, sym)
        bits = torch.full((S, pitch), -1, dtype=torch.int32, device="cuda")
        out = torch.full((S, pitch), -1, dtype=torch.int32, device="cuda")
        nb = torch.zeros(S, dtype=torch.int32, device="cuda")
        capi.check(L.sdrb_manchester_decode(d_sym.data_ptr(), pitch, nsym.data_ptr(), bc, mst.data_ptr(), bits.data_ptr(), pitch,
                                            nb.data_ptr(), S, None))
        capi.check(L.sdrb_differential_decode(bits.data_ptr(), pitch, nb.data_ptr(), bc, last.data_ptr(), out.data_ptr(), pitch, S, None))
        torch.cuda.synchronize()
        nbh, bh, oh = nb.cpu().numpy(), bits.cpu().numpy(), out.cpu().numpy()
        for s in range(S):
            man[s].append(bh[s, :nbh[s]])
            dec[s].append(oh[s, :nbh[s]])
            lens[s].append(int(nbh[s]))
            assert (bh[s, nbh[s]:] == -1).all() and (oh[s, nbh[s]:] == -1).all(), "wrote past the reported length"
        bc += 1
    msth, lasth = mst.cpu().numpy().reshape(S, 2), last.cpu().numpy()
    return [(np.concatenate(man[s]), np.concatenate(dec[s]), np.array(lens[s], np.int32), (int(msth[s, 0]), int(msth[s, 1]), int(lasth[s])))
            for s in range(S)]


def test_manchester_differential_golden_and_batch(capi, oracle, torch):
    """The reference's own vectors (tests/golden/ops.npz, made from src/rds_utilities.cpp:34-88), then a ragged random
    batch against the oracle, including block_count == 0 (the pairing-phase estimate) and a one-symbol block."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ops.npz"))
    cuts = np.cumsum(z["bits_lens"])[:-1]
    blocks = np.split(z["bits_symbols"], cuts)
    got = _run_bits_gpu(capi, torch, [[b, b] for b in blocks], block0=6)
    for s in range(2):
        man, dec, lens, state = got[s]
        assert np.array_equal(man, z["bits_manchester"]) and np.array_equal(dec, z["bits_decoded"])
        assert np.array_equal(lens, z["bits_out_lens"])
        assert list(state) == [int(v) for v in z["bits_state"]]  # (half_symbol, start, last_bit)
    rng = np.random.default_rng(5)
    for block0 in (0, 3):
        S, nblocks = 37, 6
        sym_blocks = [[rng.integers(0, 2, int(rng.integers(4, 150))).astype(np.int32) for _ in range(S)] for _ in range(nblocks)]
        sym_blocks[2][5] = np.array([1, 0, 1], np.int32)
        got = _run_bits_gpu(capi, torch, sym_blocks, block0)
        for s in range(S):
            man, dec, lens, state = oracle.bits([sym_blocks[b][s] for b in range(nblocks)], block0=block0)
            gm, gd, gl, gs = got[s]
            assert np.array_equal(gm, man) and np.array_equal(gd, dec) and np.array_equal(gl, lens), (block0, s)
            assert gs == tuple(state), (block0, s, gs, state)


def test_frame_sync_golden_and_batch(capi, oracle, torch):
    """sdrb_frame_sync on the reference's vectors (two calls of 545 bits: groups, carry, register), then many streams
    with different bit streams and chunkings against the oracle."""
    import os
    L = capi.lib()
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ops.npz"))

    def run(chunks_per_stream, max_groups=32):
        S = len(chunks_per_stream)
        ncalls = len(chunks_per_stream[0])
        st = torch.zeros(S * capi.FRAMESYNC_STATE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        groups = [[] for _ in range(S)]
        per_call = [[] for _ in range(S)]
        for c in range(ncalls):
            pitch = max(len(chunks_per_stream[s][c]) for s in range(S)) + 3
            bits = np.zeros((S, pitch), np.int32)
            for s in range(S):
                bits[s, :len(chunks_per_stream[s][c])] = chunks_per_stream[s][c]
            nb = dev(torch, np.array([len(chunks_per_stream[s][c]) for s in range(S)], np.int32))
            d_bits = dev(torch, bits)
            g = torch.zeros((S, max_groups), dtype=torch.int64, device="cuda")
            ng = torch.zeros(S, dtype=torch.int32, device="cuda")
            capi.check(L.sdrb_frame_sync(d_bits.data_ptr(), pitch, nb.data_ptr(), pitch, st.data_ptr(), g.data_ptr(), max_groups,
                                         ng.data_ptr(), max_groups, S, None))
            torch.cuda.synchronize()
            gh, ngh = g.cpu().numpy().view(np.uint64), ng.cpu().numpy()
            for s in range(S):
                groups[s].extend(gh[s, :ngh[s]].tolist())
                per_call[s].append(int(ngh[s]))
        sth = st.cpu().numpy().view(capi.FRAMESYNC_STATE_DTYPE)
        return groups, per_call, sth

    cuts = np.cumsum(z["fs_lens"])[:-1]
    chunks = np.split(z["fs_bits"], cuts)
    groups, per_call, st = run([chunks, chunks])
    for s in range(2):
        assert np.array_equal(np.array(groups[s], np.uint64), z["fs_groups"])
        assert np.array_equal(np.array(per_call[s], np.int32), z["fs_groups_per_call"])
        assert int(st[s]["reg"]) == int(z["fs_state"][0])
        assert np.array_equal(st[s]["carry"][: st[s]["ncarry"]].astype(np.int32), z["fs_carry"])

    # many streams: the golden bit stream rotated by s bits (every alignment of the 26-bit grid), cut at random places,
    # with bit errors in some streams, plus calls shorter than one window
    rng = np.random.default_rng(9)
    base = z["fs_bits"]
    S = 41
    per_stream = []
    for s in range(S):
        b = np.roll(base, s).copy()
        if s % 3 == 2:
            b[rng.integers(0, b.size, 12)] ^= 1
        cuts = np.sort(rng.integers(0, b.size, 3))
        if s == 7:
            cuts = np.array([5, 20, 30])
        per_stream.append([c.astype(np.int32) for c in np.split(b, cuts)])
    groups, per_call, st = run(per_stream)
    for s in range(S):
        want_groups, want_calls, _, (reg, _, _), carry = oracle.frame_sync(per_stream[s])
        assert np.array_equal(np.array(groups[s], np.uint64), want_groups), s
        assert np.array_equal(np.array(per_call[s], np.int32), want_calls), s
        assert int(st[s]["reg"]) == int(reg), s
        assert np.array_equal(st[s]["carry"][: st[s]["ncarry"]].astype(np.int32), carry), s


def _run_rds_sync(capi, torch, chunks_per_stream, want_syndromes=False):
    """sdrb_rds_sync over per-stream chunk lists (same number of calls per stream).  Returns per stream (events, state, syndromes)."""
    L = capi.lib()
    S = len(chunks_per_stream)
    ncalls = max(len(c) for c in chunks_per_stream)
    st = torch.zeros(S * capi.RDS_SYNC_STATE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    events = [[] for _ in range(S)]
    syn = [[] for _ in range(S)]
    for c in range(ncalls):
        lens = [len(chunks_per_stream[s][c]) if c < len(chunks_per_stream[s]) else 0 for s in range(S)]
        pitch = max(lens) + 5
        bits = np.zeros((S, pitch), np.int32)
        for s in range(S):
            if lens[s]:
                bits[s, :lens[s]] = chunks_per_stream[s][c]
        cap = max(lens) // 26 + 8
        nb = dev(torch, np.array(lens, np.int32))
        d_bits = dev(torch, bits)
        ev = torch.zeros(S * cap * capi.RDS_SYNC_EVENT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        nev = torch.zeros(S, dtype=torch.int32, device="cuda")
        d_syn = torch.zeros((S, pitch), dtype=torch.int16, device="cuda") if want_syndromes else None
        capi.check(L.sdrb_rds_sync(d_bits.data_ptr(), pitch, nb.data_ptr(), max(lens), st.data_ptr(), ev.data_ptr(), cap, nev.data_ptr(), cap,
                                   d_syn.data_ptr() if want_syndromes else None, S, None))
        torch.cuda.synchronize()
        evh = ev.cpu().numpy().view(capi.RDS_SYNC_EVENT_DTYPE).reshape(S, cap)
        nevh = nev.cpu().numpy()
        for s in range(S):
            assert nevh[s] <= cap
            events[s].extend((int(e["type"]), int(e["bit"]), int(e["a"]), int(e["b"]), int(e["value"])) for e in evh[s, :nevh[s]])
            if want_syndromes:
                syn[s].append(d_syn.cpu().numpy().view(np.uint16)[s, :lens[s]].copy())
    sth = st.cpu().numpy().view(capi.RDS_SYNC_STATE_DTYPE)
    return events, sth, syn


def test_rds_sync_state_machine(capi, oracle, oracle_mod, sdrgen, torch):
    """sdrb_rds_sync = the reference's error_detection (src/rds_utilities.cpp:202-311): on every stream of tests/rds_streams.py,
    in the chain's ragged chunks and in one piece, as one ragged batch: events 1-4 and all carried state equal the oracle's
    (which equals the unmodified reference function, tests/test_oracle_vs_ref.py and tests/golden/errdet.npz); the per-bit
    syndromes equal calc_syndrome on the running register; the extension's events (type 5) are exactly the groups whose
    four blocks are intact."""
    import rds_streams
    cases = rds_streams.cases(sdrgen)
    names = sorted(cases)
    per_stream = [rds_streams.chunks_of(cases[n]) for n in names]
    per_stream += [rds_streams.chunks_of(cases[n], sizes=(701, 13, 1, 999, 2600)) for n in names]   # other chunkings, one-bit calls
    events, st, syn = _run_rds_sync(capi, torch, per_stream, want_syndromes=True)
    for s, chunks in enumerate(per_stream):
        name = names[s % len(names)]
        want_ev, _, st64, wst, nun = oracle.error_detection(chunks, debug_lines=False)
        ref_events = [e for e in events[s] if e[0] != 5]
        assert ref_events == [tuple(int(v) for v in e) for e in want_ev], (name, s)
        assert int(st[s]["reg"]) == int(st64[0]), name
        for k in oracle_mod.ERRDET_STATE_NAMES:
            assert int(st[s][k]) == wst[k], (name, k)
        # calc_syndrome(reg, 26) after every bit
        bits = cases[name]
        got = np.concatenate(syn[s])
        reg = 0
        for i in range(0, bits.size, 97):  # spot checks along the stream (the oracle function is scalar Python-speed)
            reg = 0
            for b in bits[max(0, i - 25):i + 1]:
                reg = (reg << 1) | int(b)
            assert int(got[i]) == int(oracle.lib.orc_calc_syndrome(reg, 26)), (name, i)
    # the extension: complete groups.  Clean stream: every group after the sync point, in order, with the right content.
    ext = [e[4] for e in events[names.index("clean")] if e[0] == 5]
    words = [sdrgen.rds_group_0a(0x1234, 5, "B200-SDR", g & 3) for g in range(40)]
    all_groups = [(w[0] << 48) | (w[1] << 32) | (w[2] << 16) | w[3] for w in words]
    assert len(ext) >= 38 and ext == all_groups[-len(ext):]
    dec = capi.RdsTextDecoder()
    for g in ext:
        dec.feed(g)
    assert b"PI: 1234" in dec.text and b"PTY: Rock" in dec.text and b"Program Service: B200-SDR" in dec.text
    # 1 % bit errors: fewer groups, but never a wrong one
    ext_ber = [e[4] for e in events[names.index("ber_1pct")] if e[0] == 5]
    assert 5 < len(ext_ber) < len(ext) and set(ext_ber) <= set(all_groups)
    assert [e for e in events[names.index("noise")] if e[0] == 5] == []


def test_rds_sync_argument_checks(capi, torch):
    L = capi.lib()
    z = torch.zeros(64, dtype=torch.int32, device="cuda")
    assert L.sdrb_rds_sync(None, 8, z.data_ptr(), 8, z.data_ptr(), z.data_ptr(), 1, z.data_ptr(), 1, None, 1, None) == capi.SDRB_ERR_INVALID
    assert L.sdrb_rds_sync(z.data_ptr(), 8, z.data_ptr(), 9000, z.data_ptr(), z.data_ptr(), 1, z.data_ptr(), 1, None, 1, None) == capi.SDRB_ERR_INVALID
