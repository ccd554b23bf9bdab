"""Parity of the CUDA receive chain (through the C ABI) against the oracle, bit for bit.

Every comparison is exact: u8 unpack, decimation indexing, every float32 intermediate, the int16 PCM,
CDR offsets, symbols, RDS bits, group registers and the stderr text (the float
tolerance of 1e-4 the north star allows is not needed because the chain reproduces the reference's
operation order and its libm roundings).
"""
from __future__ import annotations

import numpy as np
import pytest

from chain_compare import FLOAT_STAGES, RDS_KEYS, diff_report, run_cuda_chain

pytestmark = pytest.mark.gpu


def _assert_same(got, want, keys, what):
    rep = diff_report(got, want, keys)
    bad = {k: v for k, v in rep.items() if v is not None}
    assert not bad, f"{what}: stages that differ from the oracle: {bad}"


@pytest.mark.parametrize("mode,kind,nblocks", [(0, "r", 40), (0, "s", 12), (0, "m", 12), (2, "m", 10), (1, "s", 8),
                                               (1, "m", 6), (3, "m", 6), (2, "s", 6), (3, "s", 5)])
def test_single_stream_all_stages(capi, oracle, station_iq, mode, kind, nblocks):
    iq = station_iq(0, mode, nblocks)
    stages = FLOAT_STAGES[kind]
    got = run_cuda_chain(capi, mode, kind, [iq], nblocks, stages=stages)[0]
    want = oracle.chain(mode, kind, iq, stages=stages)
    _assert_same(got, want, ["pcm"] + stages + (RDS_KEYS if kind == "r" else []), f"mode {mode} {kind}")
    if kind == "r":
        assert len(want["groups"]) >= 5, "the synthetic station must decode RDS groups for this test to mean anything"
        assert b"PI: 1234" in bytes(got["text"]) and b"PTY: Rock" in bytes(got["text"])


def test_long_run_rds_text(capi, oracle, station_iq):
    """130 blocks (4 s): the PS name appears and every bit and group matches."""
    nblocks = 130
    iq = station_iq(0, 0, nblocks)
    got = run_cuda_chain(capi, 0, "r", [iq], nblocks)[0]
    want = oracle.chain(0, "r", iq)
    _assert_same(got, want, ["pcm"] + RDS_KEYS, "mode 0 r, 130 blocks")
    assert b"Program Service: B200-SDR" in bytes(got["text"])


def test_batch_equals_single_streams(capi, oracle, station_iq):
    """Stream k of a ragged batch (33 streams: one full warp of PLL lanes plus one) == its own oracle run."""
    nblocks, S = 10, 33
    iqs = [station_iq(k % 5, 0, nblocks) for k in range(S)]
    got = run_cuda_chain(capi, 0, "r", iqs, nblocks, pitch_pad=64)
    wants = {k: oracle.chain(0, "r", station_iq(k, 0, nblocks)) for k in range(5)}
    for s in range(S):
        _assert_same(got[s], wants[s % 5], ["pcm"] + RDS_KEYS, f"stream {s}")


def test_device_input_and_unpadded_pitch(capi, oracle, station_iq):
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    nblocks = 8
    iq = station_iq(1, 0, nblocks)
    got = run_cuda_chain(capi, 0, "s", [iq, iq], nblocks, device_input=True)
    want = oracle.chain(0, "s", iq)
    for s in range(2):
        _assert_same(got[s], want, ["pcm"], f"device input stream {s}")


def test_overlap_mode_matches(capi, oracle, station_iq):
    nblocks = 24
    iq = station_iq(2, 0, nblocks)
    got = run_cuda_chain(capi, 0, "r", [iq] * 3, nblocks, overlap=True)
    want = oracle.chain(0, "r", iq)
    for s in range(3):
        _assert_same(got[s], want, ["pcm"] + RDS_KEYS, f"overlap stream {s}")


def test_state_save_load_resumes_bit_exact(capi, oracle, station_iq):
    nblocks, cut = 24, 11
    iq = station_iq(0, 0, nblocks)
    want = oracle.chain(0, "r", iq)
    bb = None
    pcm, bits = [], []
    with capi.Chain(0, "r", n_streams=2) as a:
        bb = a.info.block_bytes
        for b in range(cut):
            blk = np.stack([iq[b * bb:(b + 1) * bb]] * 2)
            a.process_host(blk)
            pcm.append(a.read_pcm()[1].copy())
            r = a.read_rds()[1]
            bits.append(r["bits"][: r["n_bits"]].astype(np.int32))
        blob = a.state_save()
    with capi.Chain(0, "r", n_streams=2) as c2:
        c2.state_load(blob)
        for b in range(cut, nblocks):
            blk = np.stack([iq[b * bb:(b + 1) * bb]] * 2)
            c2.process_host(blk)
            pcm.append(c2.read_pcm()[1].copy())
            r = c2.read_rds()[1]
            bits.append(r["bits"][: r["n_bits"]].astype(np.int32))
    assert np.array_equal(np.concatenate(pcm), want["pcm"])
    assert np.array_equal(np.concatenate(bits), want["rds_bits"])


def test_edge_inputs(capi, oracle):
    """All-128 input (I = Q = 0: the discriminator's guarded branch), all-0, all-255 and white noise."""
    rng = np.random.default_rng(7)
    bp = 73500
    for name, iq in (("zeros128", np.full(2 * bp * 3, 128, np.uint8)), ("min", np.zeros(2 * bp * 3, np.uint8)),
                     ("max", np.full(2 * bp * 3, 255, np.uint8)), ("noise", rng.integers(0, 256, 2 * bp * 8, dtype=np.uint8))):
        nblocks = iq.size // (2 * bp)
        stages = FLOAT_STAGES["r"]
        got = run_cuda_chain(capi, 0, "r", [iq], nblocks, stages=stages)[0]
        want = oracle.chain(0, "r", iq, stages=stages)
        _assert_same(got, want, ["pcm"] + stages + RDS_KEYS, name)


def test_golden_fixture(capi):
    """The committed fixture (made from the unmodified reference sources, tests/golden/make_golden.py)."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "chain_m0_r.npz")
    if not os.path.exists(path):
        pytest.skip("golden fixture missing")
    z = np.load(path)
    from conftest import load_module
    gen = load_module("sdrgen", "real-time-sdr_b200/sdrgen.py")
    nblocks = int(z["nblocks"])
    iq = gen.generate_iq(gen.Station(), gen.block_pairs(0) * nblocks)
    got = run_cuda_chain(capi, 0, "r", [iq], nblocks, stages=["fm_demod", "carrier", "rds_clean"])[0]
    assert np.array_equal(got["pcm"], z["pcm"])
    assert np.array_equal(got["rds_bits"], z["rds_bits"])
    assert np.array_equal(got["groups"], z["groups"])
    assert np.array_equal(got["cdr_offset"], z["cdr_offset"])
    for st in ("fm_demod", "carrier", "rds_clean"):
        sl = z[st + "_slice"]
        off = int(z[st + "_offset"])
        assert got[st][off:off + sl.size].tobytes() == sl.tobytes(), st


def test_errors_are_reported_not_fatal(capi):
    with capi.Chain(0, "m", 1) as ch:
        with pytest.raises(capi.SdrError):
            ch.read_pcm()              # nothing processed yet
        with pytest.raises(capi.SdrError):
            ch.read_rds()              # no RDS in a mono chain
        with pytest.raises(capi.SdrError):
            ch.process_host(np.zeros((1, 100), np.uint8), 100)  # pitch smaller than a block


def test_overlap_pipelined_lagged_reads(capi, oracle, station_iq):
    """Overlap mode driven the way bench.py's e2e loop drives it: block i is issued before the results of block i-1
    are read (lag 1); three streams of different stations, host input, 30 blocks."""
    nblocks, S = 30, 3
    iqs = [station_iq(k, 0, nblocks) for k in range(S)]
    wants = [oracle.chain(0, "r", iqs[k]) for k in range(S)]
    with capi.Chain(0, "r", n_streams=S) as ch:
        ch.set_overlap(True)
        bb = ch.info.block_bytes
        bufs = [np.stack([iqs[s][b * bb:(b + 1) * bb] for s in range(S)]) for b in range(nblocks)]
        pcm = np.zeros((S, ch.info.pcm_per_block), np.int16)
        rec = np.zeros(S, capi.RDS_RECORD_DTYPE)
        got_pcm = [[] for _ in range(S)]
        got_bits = [[] for _ in range(S)]
        got_groups = [[] for _ in range(S)]

        def take():
            for s in range(S):
                got_pcm[s].append(pcm[s].copy())
                got_bits[s].append(rec[s]["bits"][: rec[s]["n_bits"]].astype(np.int32))
                got_groups[s].extend(int(g) for g in rec[s]["groups"][: rec[s]["n_groups"]])

        ch.process_host(bufs[0])
        for b in range(1, nblocks):
            ch.process_host(bufs[b])
            ch.read_results(1, pcm, rec)
            take()
        ch.read_results(0, pcm, rec)
        take()
    for s in range(S):
        assert np.array_equal(np.concatenate(got_pcm[s]), wants[s]["pcm"]), s
        assert np.array_equal(np.concatenate(got_bits[s]), wants[s]["rds_bits"]), s
        assert np.array_equal(np.array(got_groups[s], np.uint64), wants[s]["groups"]), s


def test_full_size_batch_1024_streams(capi, oracle, station_iq):
    """BASELINE.json configs[4] at full width: 1024 stations (8 distinct ones, each delayed by s // 8 blocks like bench.py
    builds them) for 24 blocks in overlap mode with lagged reads.  Every one of the 1024 streams must equal the oracle run
    of its own block sequence: PCM bit for bit, the same CDR offset in every block, the same RDS bits and the same group
    registers from the two frame-sync calls the 40 blocks contain; no capacity counter may have moved."""
    S, nblocks, M, NB = 1024, 40, 8, 8
    bb = 147000
    src = np.stack([station_iq(k, 0, NB).reshape(NB, bb) for k in range(M)])  # [station][block][bytes]
    sidx = np.arange(S)

    def step_input(g):
        return src[sidx % M, (g + sidx // M) % NB]

    # oracle: stream s sees station s % M starting at block (s // M) % NB, wrapping inside the NB generated blocks
    want = {}
    for k in range(M):
        for d in range(NB):
            seq = np.concatenate([src[k, (g + d) % NB] for g in range(nblocks)])
            want[(k, d)] = oracle.chain(0, "r", seq)
    with capi.Chain(0, "r", n_streams=S) as ch:
        ch.set_overlap(True)
        pcm = np.zeros((S, ch.info.pcm_per_block), np.int16)
        rec = np.zeros(S, capi.RDS_RECORD_DTYPE)
        all_pcm = np.zeros((nblocks, S, ch.info.pcm_per_block), np.int16)
        nbits = np.zeros((nblocks, S), np.int32)
        bits = np.zeros((nblocks, S, 48), np.uint8)
        offs = np.zeros((nblocks, S), np.int32)
        ngr = np.zeros((nblocks, S), np.int32)
        grp = np.zeros((nblocks, S, 8), np.uint64)

        def keep(b):
            all_pcm[b], nbits[b], bits[b], offs[b], ngr[b], grp[b] = pcm, rec["n_bits"], rec["bits"], rec["cdr_offset"], rec["n_groups"], rec["groups"]

        bufs = [np.ascontiguousarray(step_input(g)) for g in range(min(nblocks, NB))]
        ch.process_host(bufs[0])
        for b in range(1, nblocks):
            ch.process_host(bufs[b % NB])
            ch.read_results(1, pcm, rec)
            keep(b - 1)
        ch.read_results(0, pcm, rec)
        keep(nblocks - 1)
        assert ch.rds_overflows() == (0, 0, 0)
    bad = []
    for s in range(S):
        w = want[(s % M, (s // M) % NB)]
        got_bits = np.concatenate([bits[b, s, : nbits[b, s]] for b in range(nblocks)]).astype(np.int32)
        got_groups = np.concatenate([grp[b, s, : ngr[b, s]] for b in range(nblocks)])
        got_offs = offs[6:, s]  # the decoder is gated for the first six blocks (-1)
        if (not np.array_equal(all_pcm[:, s, :].reshape(-1), w["pcm"]) or not np.array_equal(got_bits, w["rds_bits"])
                or not np.array_equal(got_groups, w["groups"]) or not np.array_equal(got_offs, w["cdr_offset"][-got_offs.size:])
                or not (offs[:6, s] == -1).all()):
            bad.append(s)
    assert sum(len(w["groups"]) for w in want.values()) > 0, "the 40 blocks must complete RDS groups for this test to mean anything"
    assert not bad, f"{len(bad)} of {S} streams differ from their oracle run, first: {bad[:8]}"


def test_state_save_load_in_overlap_mode_single_stream(capi, oracle, station_iq):
    """Checkpoint taken while the pipeline is in flight (overlap mode, one stream): the resumed run continues bit-exactly."""
    nblocks, cut = 20, 9
    iq = station_iq(4, 0, nblocks)
    want = oracle.chain(0, "s", iq)
    pcm = []
    with capi.Chain(0, "s", n_streams=1) as a:
        a.set_overlap(True)
        bb = a.info.block_bytes
        for b in range(cut):
            a.process_host(iq[b * bb:(b + 1) * bb].reshape(1, bb))
            pcm.append(a.read_pcm()[0].copy())
        a.process_host(iq[cut * bb:(cut + 1) * bb].reshape(1, bb))  # in flight when the state is taken
        blob = a.state_save()
        pcm.append(a.read_pcm()[0].copy())
    with capi.Chain(0, "s", n_streams=1) as c2:
        c2.set_overlap(True)
        c2.state_load(blob)
        for b in range(cut + 1, nblocks):
            c2.process_host(iq[b * bb:(b + 1) * bb].reshape(1, bb))
            pcm.append(c2.read_pcm()[0].copy())
    assert np.array_equal(np.concatenate(pcm), want["pcm"])


@pytest.mark.parametrize("mode,kind,nblocks,S", [(2, "s", 4, 5), (1, "s", 5, 7), (3, "m", 5, 40), (2, "m", 4, 33)])
def test_other_modes_batched(capi, oracle, station_iq, mode, kind, nblocks, S):
    """Modes 1-3 (other front-end decimations, 147/800 and 147/1280 audio resamplers) as ragged batches, overlap mode."""
    iqs = [station_iq(k % 3, mode, nblocks) for k in range(S)]
    wants = [oracle.chain(mode, kind, station_iq(k, mode, nblocks)) for k in range(3)]
    got = run_cuda_chain(capi, mode, kind, iqs, nblocks, overlap=True)
    for s in range(S):
        assert np.array_equal(got[s]["pcm"], wants[s % 3]["pcm"]), (mode, kind, s)


def test_argument_checks(capi):
    torch = pytest.importorskip("torch")
    with capi.Chain(0, "m", 2) as ch:
        buf = torch.zeros((2, ch.info.block_bytes + 16), dtype=torch.uint8, device="cuda")
        with pytest.raises(capi.SdrError):
            ch.process_device(buf.data_ptr() + 1, ch.info.block_bytes + 16)   # odd address
        with pytest.raises(capi.SdrError):
            ch.process_device(buf.data_ptr(), ch.info.block_bytes + 1)        # odd pitch
        with pytest.raises(capi.SdrError):
            ch.read_results(2, None, None)                                     # lag out of range
        ch.process_device(buf.data_ptr() + 2, ch.info.block_bytes + 16)       # 2-byte aligned, not 16: plain-load path
        ch.sync()
        assert ch.read_pcm().shape == (2, 1470)


def _patch_pll_sample_count(ch, blob: bytes, S: int, n0: float) -> bytes:
    """Rewrites both PLL states inside a state blob (their offsets come from sdrb_chain_state_item_offset)."""
    import struct
    offs = (ch.state_item_offset("pll19"), ch.state_item_offset("pll114"))
    assert min(offs) > 0
    out = bytearray(blob)
    for which, freq in ((0, 19e3), (1, 114e3)):
        for s in range(S):
            o = offs[which] + s * 24
            fbI, fbQ, integ, phase, trig = struct.unpack_from("<4fd", out, o)
            assert (fbI, fbQ, integ, phase, trig) == (1.0, 0.0, 0.0, 0.0, 0.0), "fresh PLL state expected here"
            th = np.float32(2 * np.pi * float(np.float32(freq) / np.float32(240000.0)) * n0 + phase)
            struct.pack_into("<4fd", out, o, float(np.float32(np.cos(np.float64(th)))), float(np.float32(np.sin(np.float64(th)))),
                             integ, phase, float(n0))
    return bytes(out)


@pytest.mark.parametrize("n0", [3.0e6, 2.0e7, 4.0e8, 1.3e9, 5.0e10])
def test_chain_at_large_nco_phase(capi, oracle, station_iq, n0):
    """The batched PLL kernel hours into a run: both loops are placed at sample count n0 through the checkpoint interface
    (the oracle through its test hook), where the float NCO phase has an ulp of up to 16384 rad (n0 = 1.3e9 is 90 minutes in:
    the 114 kHz loop's phase is beyond the 3e9 at which round 1's own reduction stopped; 5e10 is 58 hours).  Everything
    downstream of the PLLs (carrier, IPLL, audio, RDS samples) must still be bit-identical, and the kernel must still be
    on its fast path (the careful path produces the same values at several times the cost)."""
    nblocks, S = 10, 2
    iq = station_iq(0, 0, nblocks)
    stages = ["carrier", "IPLL", "stereo_filt", "rds_clean"]
    want = oracle.chain(0, "r", iq, stages=stages, pll_sample_count=n0)
    acc = {k: [] for k in stages}
    pcm = []
    with capi.Chain(0, "r", n_streams=S, keep_stages=True) as ch:
        ch.state_load(_patch_pll_sample_count(ch, ch.state_save(), S, n0))
        bb = ch.info.block_bytes
        for b in range(nblocks):
            ch.process_host(np.stack([iq[b * bb:(b + 1) * bb]] * S))
            pcm.append(ch.read_pcm()[1].copy())
            for k in stages:
                acc[k].append(ch.stage(k)[1])
        redo = ch.pll_redos()
    got = {k: np.concatenate(v) for k, v in acc.items()}
    got["pcm"] = np.concatenate(pcm)
    _assert_same(got, want, ["pcm"] + stages, f"n0={n0}")
    chunks = S * nblocks * (7350 // 4)
    assert redo[0] < 0.01 * chunks and redo[1] < 0.01 * chunks, f"careful-path chunks {redo} of {chunks} per loop"


@pytest.mark.parametrize("cap,S", [("1", 70), ("4", 200), ("2", 33)])
def test_pll_cta_sizes(capi, oracle, station_iq, monkeypatch, cap, S):
    """k_pll with 256-, 128- and 64-thread CTAs (what large batches use; forced here through the SM-budget knob):
    every stream still equals its own oracle run, including the partly filled last CTA."""
    monkeypatch.setenv("SDRB_PLL_MAX_CTAS", cap)
    nblocks = 8
    iqs = [station_iq(k % 3, 0, nblocks) for k in range(S)]
    got = run_cuda_chain(capi, 0, "r", iqs, nblocks)
    wants = {k: oracle.chain(0, "r", station_iq(k, 0, nblocks)) for k in range(3)}
    for s in range(S):
        _assert_same(got[s], wants[s % 3], ["pcm"] + RDS_KEYS, f"cap {cap} stream {s}")


@pytest.mark.parametrize("setting", ["0", "1", None])
def test_sm_partition_is_optional_and_changes_nothing(capi, oracle, station_iq, monkeypatch, setting):
    """The overlap-mode streams live in two green contexts (PLL: 32 SMs, FIR: the rest) unless the driver refuses, the
    batch is FIR-bound (more than 1024 stereo+RDS stations) or SDRB_SM_PARTITION=0: the results are the oracle's either way."""
    if setting is None:
        monkeypatch.delenv("SDRB_SM_PARTITION", raising=False)
    else:
        monkeypatch.setenv("SDRB_SM_PARTITION", setting)
    with capi.Chain(0, "r", n_streams=5) as ch:
        part = ch.sm_partition()
    if setting == "0":
        assert part == (0, 0)
    else:
        assert part == (0, 0) or (part[0] >= 32 and part[1] > 0 and part[0] + part[1] <= 148), part
    with capi.Chain(0, "r", n_streams=1100) as ch:  # 70 PLL warps: the FIR kernels bound the step, no partition by default
        assert setting == "1" or ch.sm_partition() == (0, 0)
    nblocks = 8
    iqs = [station_iq(k % 3, 0, nblocks) for k in range(5)]
    got = run_cuda_chain(capi, 0, "r", iqs, nblocks, overlap=True)
    wants = {k: oracle.chain(0, "r", station_iq(k, 0, nblocks)) for k in range(3)}
    for s in range(5):
        _assert_same(got[s], wants[s % 3], ["pcm"] + RDS_KEYS, f"partition {setting} stream {s}")


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_type_r_outside_mode0_is_the_stereo_chain_with_gated_records(capi, oracle, station_iq, mode):
    """`project <1|2|3> r` in the reference plays stereo audio and its RDS thread prints nothing
    (tests/test_oracle_vs_ref.py::test_reference_rds_thread_is_silent_outside_mode0)."""
    nblocks = 5
    iq = station_iq(0, mode, nblocks)
    want = oracle.chain(mode, "s", iq)
    with capi.Chain(mode, "r", n_streams=2) as ch:
        bb = ch.info.block_bytes
        assert ch.info.rds_block == 0
        pcm = []
        for b in range(nblocks):
            ch.process_host(np.stack([iq[b * bb:(b + 1) * bb]] * 2))
            pcm.append(ch.read_pcm()[1].copy())
            rec = ch.read_rds()
            assert (rec["cdr_offset"] == -1).all() and (rec["n_bits"] == 0).all() and (rec["n_groups"] == 0).all()
    assert np.array_equal(np.concatenate(pcm), want["pcm"])


def test_state_blob_is_size_checked_and_results_need_a_block(capi, station_iq):
    iq = station_iq(0, 0, 3)
    with capi.Chain(0, "r", n_streams=2) as a:
        bb = a.info.block_bytes
        for b in range(2):
            a.process_host(np.stack([iq[b * bb:(b + 1) * bb]] * 2))
        blob = a.state_save()
        assert a.rds_overflows() == (0, 0, 0)
    with capi.Chain(0, "r", n_streams=2) as c2:
        with pytest.raises(capi.SdrError) as e:
            c2.state_load(blob[: len(blob) // 2])   # truncated file
        assert e.value.code == capi.SDRB_ERR_INVALID
        with pytest.raises(capi.SdrError):
            c2.state_load(blob[:16])                 # shorter than the header
        c2.state_load(blob)
        with pytest.raises(capi.SdrError) as e:
            c2.read_pcm()                            # loaded, but no block processed since: nothing to read
        assert e.value.code == capi.SDRB_ERR_STATE
        with pytest.raises(capi.SdrError):
            c2.stage("fm_demod")
        c2.process_host(np.stack([iq[2 * bb:3 * bb]] * 2))
        c2.read_pcm()
        assert c2.stage("fm_demod").shape == (2, c2.info.if_block)  # ring-backed stages need no keep_stages
        with pytest.raises(capi.SdrError):
            c2.stage("carrier")                      # flat dumps do
    with capi.Chain(0, "s", n_streams=2) as c3:
        with pytest.raises(capi.SdrError):
            c3.state_load(blob)                      # a blob of another configuration


def test_input_consumed_query(capi, station_iq):
    iq = station_iq(0, 0, 4)
    with capi.Chain(0, "s", n_streams=4) as ch:
        ch.set_overlap(True)
        bb = ch.info.block_bytes
        with pytest.raises(capi.SdrError):
            ch.input_consumed(0)
        for b in range(3):
            ch.process_host(np.stack([iq[b * bb:(b + 1) * bb]] * 4))
        ch.sync()
        assert ch.input_consumed(0) and ch.input_consumed(1)


@pytest.mark.parametrize("mode,kind,S,nblocks,overlap", [(0, "r", 33, 8, True), (2, "m", 3, 3, False), (3, "s", 3, 3, False), (1, "s", 5, 3, True),
                                                        (0, "r", 1, 22, False)])
def test_guard_zones_stay_intact(capi, oracle, station_iq, monkeypatch, mode, kind, S, nblocks, overlap):
    """compute-sanitizer is closed on the GPU pool (profiles/sanitizer_r2.txt); its memcheck is replaced by canary zones
    around every device allocation of the chain (SDRB_GUARD=1): after a run on ragged stream counts no kernel may have written
    outside its buffers, and the results must still equal the oracle's."""
    monkeypatch.setenv("SDRB_GUARD", "1")
    iq = station_iq(0, mode, nblocks)
    want = oracle.chain(mode, kind, iq)
    with capi.Chain(mode, kind, n_streams=S) as ch:
        ch.set_overlap(overlap)
        bb = ch.info.block_bytes
        pcm = []
        for b in range(nblocks):
            ch.process_host(np.ascontiguousarray(np.stack([iq[b * bb:(b + 1) * bb]] * S)))
            pcm.append(ch.read_pcm()[S - 1].copy())
        assert ch.check_guards() >= 8
    assert np.array_equal(np.concatenate(pcm), want["pcm"])
    monkeypatch.delenv("SDRB_GUARD")
    with capi.Chain(mode, kind, n_streams=1) as ch:
        with pytest.raises(capi.SdrError):
            ch.check_guards()
