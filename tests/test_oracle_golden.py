"""Pins the oracle (oracle/sdr_oracle.c) against fixtures produced by the unmodified reference sources
(tests/golden/make_golden.py) and against the known-answer values recorded in SURVEY.md section 8."""
from __future__ import annotations

import os
import zlib

import numpy as np
import pytest

G = os.path.join(os.path.dirname(__file__), "golden")


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def same(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.tobytes() == b.astype(a.dtype).tobytes()


MODES = {0: (2.4e6, 10, 240e3, 1, 39), 1: (1.44e6, 4, 360e3, 1, 39), 2: (2.4e6, 10, 240e3, 147, 20), 3: (1.152e6, 3, 384e3, 147, 20)}


def _designs(design):
    """Every tap table of every mode through `design(kind, n, Fs, a, b, u)`, keyed like taps.npz."""
    out = {}
    for m, (rf_fs, dec, if_fs, up, sps) in MODES.items():
        sfx = f"_m{m}"
        out["rf_h" + sfx] = design("lpf", 101, Fs=rf_fs, a=100e3)
        out["audio_h" + sfx] = design("lpf_gain", 101 * up, Fs=float(np.float32(if_fs) * np.float32(up)), a=16e3, u=up)
        f = float(int(rf_fs) // dec)
        out["pilot_h" + sfx] = design("bpf", 101, Fs=f, a=18.5e3, b=19.5e3)
        out["carrier_h" + sfx] = design("bpf", 101, Fs=f, a=37.5e3, b=38.5e3)
        out["stereo_h" + sfx] = design("bpf", 101, Fs=f, a=22e3, b=54e3)
        out["rds_lpf_h" + sfx] = design("lpf_gain", 101 * 247, Fs=float(int(if_fs) * 247), a=3e3, u=247)
        out["rds_h" + sfx] = design("bpf", 101, Fs=if_fs, a=54e3, b=60e3)
        out["rds_pilot_h" + sfx] = design("bpf", 101, Fs=if_fs, a=113.5e3, b=114.5e3)
        out["rrc_h" + sfx] = design("rrc", 101, Fs=float(2375 * sps))
    out["apf_h"] = design("apf", 101, a=1.0)
    return out


def test_oracle_taps_match_reference(oracle):
    z = np.load(os.path.join(G, "taps.npz"))
    got = _designs(oracle.design)
    assert set(got) == set(z.files)
    for k in z.files:
        assert same(got[k], z[k]), k


def test_library_designers_match_reference(capi):
    """sdrb_design_* run on the host: this is product code checked on the CPU box."""
    z = np.load(os.path.join(G, "taps.npz"))
    got = _designs(capi.design)
    for k in z.files:
        assert same(got[k], z[k]), k


def test_tap_known_answers(oracle):
    """values printed by the reference's designers (probe P5)."""
    rf = oracle.design("lpf", 101, Fs=2.4e6, a=1e5)
    assert rf[0] == 0 and np.float32(rf[1]) == np.float32(1.6261771e-06) and np.float32(rf[50]) == np.float32(0.083313182)
    assert abs(float(rf.astype(np.float64).sum()) - 1.00100359) < 1e-6
    rrc = oracle.design("rrc", 101, Fs=2375 * 39.0)
    assert np.float32(rrc[0]) == np.float32(-0.0115760313) and rrc[50] == rrc[51] and np.float32(rrc[50]) == np.float32(1.2452141)
    pil = oracle.design("bpf", 101, Fs=240000.0, a=18.5e3, b=19.5e3)
    assert np.float32(pil[1]) == np.float32(6.60748765e-06) and np.float32(pil[50]) == np.float32(0.00804743543)
    apf = oracle.design("apf", 101, a=1.0)
    assert apf[50] == 1 and np.count_nonzero(apf) == 1


def test_oracle_ops_match_reference(oracle):
    z = np.load(os.path.join(G, "ops.npz"))
    assert same(oracle.fir_decim(z["fir_x"], z["fir_h"], 10, nblocks=3), z["fir_decim10_y"])
    assert same(oracle.fir_decim(z["fir_x"], z["fir_h"], 1, nblocks=3), z["fir_decim1_y"])
    assert same(oracle.fir_updown(z["fir_x"], z["fir_h3"], 3, 7, nblocks=3), z["fir_updown_3_7_y"])
    assert same(oracle.fir_updown(z["fir_x"], z["fir_h"], 1, 5, nblocks=3), z["fir_updown_1_5_y"])
    y, prev = oracle.fmdemod(z["dem_I"], z["dem_Q"], nblocks=2)
    assert same(y, z["dem_y"]) and same(np.array(prev, np.float32), z["dem_prev"])
    for tag, freq, scale, bw in (("pll19", 19e3, 2.0, 0.01), ("pll114", 114e3, 0.5, 0.001)):
        yy, st = oracle.pll(z[tag + "_x"], freq, 240000.0, scale, 0.0, bw, nblocks=3)
        assert same(yy, z[tag + "_y"]), tag
        assert st.trigOffset == float(z[tag + "_trig"][0])
    assert same(oracle.cdr(z["cdr_x"], 39, nblocks=3), z["cdr_offset"])
    lens = z["bits_lens"]
    chunks = np.split(z["bits_symbols"], np.cumsum(lens)[:-1])
    man, dec, out_lens, st = oracle.bits(chunks, block0=6)
    assert same(man, z["bits_manchester"]) and same(dec, z["bits_decoded"]) and same(out_lens, z["bits_out_lens"])
    assert list(st) == list(z["bits_state"])
    fchunks = np.split(z["fs_bits"], np.cumsum(z["fs_lens"])[:-1])
    groups, per_call, text, state, carry = oracle.frame_sync(fchunks)
    assert same(groups, z["fs_groups"]) and same(per_call, z["fs_groups_per_call"])
    assert text.encode("latin-1") == bytes(z["fs_text"])
    assert [int(v) for v in state] == [int(v) for v in z["fs_state"][:3]]
    assert same(carry, z["fs_carry"])


def test_oracle_chain_matches_reference_mode0_rds(oracle, sdrgen):
    z = np.load(os.path.join(G, "chain_m0_r.npz"))
    nblocks = int(z["nblocks"])
    iq = sdrgen.generate_iq(sdrgen.Station(), sdrgen.block_pairs(0) * nblocks)
    assert crc(iq) == int(z["iq_crc"]), "the synthetic generator no longer produces the bytes the fixture was made from"
    stages = ("fm_demod", "carrier", "rds_clean", "stereo_filt", "IPLL")
    r = oracle.chain(0, "r", iq, stages=stages)
    for k in ("pcm", "cdr_offset", "n_symbols", "n_bits", "rds_bits", "groups", "group_block", "text"):
        assert same(r[k], z[k]), k
    for st in stages:
        assert crc(r[st]) == int(z[st + "_crc"]), st
        off = int(z[st + "_offset"])
        assert same(r[st][off:off + z[st + "_slice"].size], z[st + "_slice"]), st
    text = bytes(r["text"]).decode()
    assert "PI: 1234" in text and "PTY: Rock" in text and len(r["groups"]) == 9


@pytest.mark.parametrize("mode,kind", [(0, "m"), (0, "s"), (2, "m"), (1, "s"), (3, "m"), (1, "m"), (2, "s")])
def test_oracle_chain_matches_reference_other_modes(oracle, sdrgen, mode, kind):
    z = np.load(os.path.join(G, "chain_misc.npz"))
    tag = f"m{mode}_{kind}"
    nb = int(z[tag + "_nblocks"])
    iq = sdrgen.generate_iq(sdrgen.Station.for_stream(0, fs=sdrgen.mode_fs(mode)), sdrgen.block_pairs(mode) * nb)
    assert crc(iq) == int(z[tag + "_iq_crc"])
    r = oracle.chain(mode, kind, iq, stages=("fm_demod",))
    assert same(r["pcm"], z[tag + "_pcm"])
    assert crc(r["fm_demod"]) == int(z[tag + "_fm_crc"])


def test_parser_registers_of_the_reference_test(oracle, capi):
    """test/parser_test.cpp:79-136 of the reference: 56 captured group registers.  Through the parser of
    src/rds_utilities.cpp:172-199 they give PI c27a, PTY Rock and three PS strings (SURVEY.md probe P9)."""
    regs = np.load(os.path.join(G, "parser_registers.npz"))["regs"]
    text = oracle.parse_groups(regs)
    assert text.count("PI: c27a\n") == 56 and text.count("PTY: Rock\n") == 56
    ps = [l[len("Program Service: "):] for l in text.split("\n") if l.startswith("Program Service: ")]
    assert ps == ["", "  Love  ", "  Dies  "] or ps == ["\x00\x00\x00\x00ters"[:0], "  Love  ", "  Dies  "], ps
    dec = capi.RdsTextDecoder()
    for r in regs:
        dec.feed(int(r))
    assert dec.text.decode("latin-1") == text


def test_error_detection_fixture(oracle, sdrgen):
    """tests/golden/errdet.npz: what the UNMODIFIED reference function error_detection (src/rds_utilities.cpp:202-311) printed
    and left in its state variables on the streams of tests/rds_streams.py; the oracle restatement must reproduce it."""
    import rds_streams
    z = np.load(os.path.join(G, "errdet.npz"))
    names = sorted(k[:-5] for k in z.files if k.endswith("_bits"))
    assert len(names) == 5
    for name in names:
        chunks = np.split(z[name + "_bits"].astype(np.int32), np.cumsum(z[name + "_lens"])[:-1])
        ev, text, st64, st, nun = oracle.error_detection(chunks)
        assert text == bytes(z[name + "_text"]), name
        assert [int(v) for v in st64] == [int(v) for v in z[name + "_state64"]], name
        from conftest import load_module
        keys = load_module("oracle_py", "oracle/oracle_py.py").ERRDET_STATE_NAMES
        assert [st[k] for k in keys] == [int(v) for v in z[name + "_state"]], name
        assert np.array_equal(z[name + "_bits"].astype(np.int32), rds_streams.cases(sdrgen)[name]), "the fixture's streams are those of tests/rds_streams.py"
