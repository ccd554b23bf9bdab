"""Oracle restatement vs the unmodified reference sources, live (only where oracle/_ref was built, i.e. in
the container that has /root/reference; elsewhere the committed fixtures of test_oracle_golden.py apply)."""
from __future__ import annotations

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
