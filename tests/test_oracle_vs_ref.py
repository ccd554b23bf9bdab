"""Oracle restatement vs the unmodified reference sources, live (only where oracle/_ref was built, i.e. in
the container that has /root/reference; elsewhere the committed fixtures of test_oracle_golden.py apply)."""
from __future__ import annotations

import os
import subprocess

import numpy as np
import pytest

STAGES = ("I_ds", "Q_ds", "fm_demod", "pilot", "carrier", "stereo_band", "stereo_dc", "mono_delay", "mono_filt",
          "stereo_filt", "rds_band", "gen_pilot", "IPLL", "rds_band_delay", "rds_dc", "rds_filt", "rds_clean")


def _cmp(a, b, keys):
    for k in keys:
        if k in b:
            x, y = np.ascontiguousarray(a[k]), np.ascontiguousarray(b[k])
            assert x.shape == y.shape and x.tobytes() == y.astype(x.dtype).tobytes(), k


def test_chain_all_stages_mode0_rds(oracle, ref, station_iq):
    iq = station_iq(3, 0, 36)
    a = oracle.chain(0, "r", iq, stages=STAGES)
    b = ref.chain(0, "r", iq, stages=STAGES)
    _cmp(a, b, STAGES + ("pcm", "cdr_offset", "n_symbols", "n_bits", "rds_bits", "symbols", "groups", "group_block", "text"))


@pytest.mark.parametrize("mode,kind,nb", [(0, "m", 5), (0, "s", 5), (1, "m", 3), (1, "s", 3), (2, "m", 4), (2, "s", 3), (3, "m", 3), (3, "s", 3)])
def test_chain_other_modes(oracle, ref, station_iq, mode, kind, nb):
    iq = station_iq(1, mode, nb)
    stages = ("fm_demod", "audio_filt") if kind == "m" else ("fm_demod", "carrier", "mono_filt", "stereo_filt")
    a = oracle.chain(mode, kind, iq, stages=stages)
    b = ref.chain(mode, kind, iq, stages=stages)
    _cmp(a, b, stages + ("pcm",))


def test_edge_inputs(oracle, ref):
    rng = np.random.default_rng(11)
    for iq in (np.full(147000 * 3, 128, np.uint8), np.zeros(147000 * 2, np.uint8), rng.integers(0, 256, 147000 * 8, dtype=np.uint8)):
        a = oracle.chain(0, "r", iq, stages=("fm_demod", "carrier", "IPLL", "rds_clean"))
        b = ref.chain(0, "r", iq, stages=("fm_demod", "carrier", "IPLL", "rds_clean"))
        _cmp(a, b, ("fm_demod", "carrier", "IPLL", "rds_clean", "pcm", "cdr_offset", "rds_bits", "groups"))


def _threaded(ref, mode, kind, iq):
    """The reference's own three-thread binary (oracle/_ref/project) on a raw IQ byte string: (stdout, stderr)."""
    if not os.path.exists(ref.project):
        pytest.skip("oracle/_ref/project not built (no /root/reference here)")
    r = subprocess.run([ref.project, str(mode), kind], input=iq.tobytes(), capture_output=True, timeout=300)
    assert r.returncode == 1  # exit(1) at EOF, src/rffrontend.cpp:50-52
    return r.stdout, r.stderr


@pytest.mark.parametrize("mode,kind,nb", [(0, "r", 60), (0, "s", 8), (0, "m", 8), (2, "m", 6)])
def test_threaded_reference_binary_end_to_end(oracle, ref, sdrgen, station_iq, mode, kind, nb):
    """the unmodified threaded binary against the oracle, end to end.  Three padding blocks keep the
    binary's EOF race (it exits while its consumers may still hold 1-2 blocks) away from the compared part."""
    pad = 3
    iq = station_iq(0, mode, nb + pad)
    out, err = _threaded(ref, mode, kind, iq)
    want = oracle.chain(mode, kind, iq)
    pcm = want["pcm"].tobytes()
    per_block = len(pcm) // (nb + pad)
    assert len(out) >= nb * per_block and len(out) % 2 == 0
    assert out == pcm[: len(out)]
    if kind == "r":
        text = bytes(want["text"])
        assert err == text[: len(err)]
        first = oracle.chain(mode, kind, iq[: nb * 2 * sdrgen.block_pairs(mode)])
        assert len(err) >= len(bytes(first["text"])) > 0
    else:
        assert err == b""


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_reference_rds_thread_is_silent_outside_mode0(oracle, ref, station_iq, mode):
    """`project <1|2|3> r`: stereo audio as in `s`, and an RDS thread that runs at rates it was not designed for and prints
    nothing.  That is what lets libsdr_b200 serve type 'r' in those modes with the stereo chain and gated RDS records."""
    nb, pad = 24, 3
    iq = station_iq(0, mode, nb + pad)
    out, err = _threaded(ref, mode, "r", iq)
    assert err == b""
    pcm = oracle.chain(mode, "s", iq)["pcm"].tobytes()
    assert len(out) >= nb * (len(pcm) // (nb + pad)) and out == pcm[: len(out)]


def test_error_detection_state_machine(oracle, ref, sdrgen):
    """The sync-state-machine decoder the reference declares and defines but never calls (include/rds_utilities.h:14,
    src/rds_utilities.cpp:202-311): the restatement against the unmodified function, on clean, noisy and sync-losing bit
    streams handed over in the chain's ragged chunks - every line it prints (the per-bit debug lines and the sticky std::hex
    after its one parse() call included) and every state variable it leaves behind."""
    import rds_streams
    for name, bits in rds_streams.cases(sdrgen).items():
        chunks = rds_streams.chunks_of(bits)
        ev, text, st64, st, nun = oracle.error_detection(chunks)
        r = ref.error_detection(chunks)
        assert bytes(r["text"]) == text, name
        assert tuple(int(v) for v in r["state64"]) == tuple(int(v) for v in st64), name
        assert [int(v) for v in r["state"]] == [st[k] for k in oracle_names()], name
    kinds = {e[0] for e in oracle.error_detection(rds_streams.chunks_of(rds_streams.cases(sdrgen)["burst_lose_sync"]))[0]}
    assert {1, 2, 3} <= kinds, "the burst case must lose sync and find it again for this test to mean anything"


def oracle_names():
    from conftest import load_module
    return load_module("oracle_py", "oracle/oracle_py.py").ERRDET_STATE_NAMES
