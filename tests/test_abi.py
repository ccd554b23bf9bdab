"""The C-ABI library on a CPU-only box: it loads, exports every symbol include/sdr_b200.h declares, its
host-side entry points work, and everything that needs a GPU fails loudly instead of falling back."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sdr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdrb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(capi):
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol(capi):
    L = capi.lib()
    for name in declared_symbols():
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sdrb_\w+)", out))
    assert set(declared_symbols()) <= exported


def test_library_is_sm100a_only(capi):
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_product_does_not_link_the_oracle(capi):
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for f in os.listdir(os.path.join(ROOT, "real-time-sdr_b200", "csrc")):
        assert "oracle" not in open(os.path.join(ROOT, "real-time-sdr_b200", "csrc", f)).read().lower().replace("// the oracle", ""), f


def test_mode_table(capi):
    """src/project.cpp:31-44,67-108 of the reference."""
    L = capi.lib()
    want = {0: (2400000, 10, 5, 1, 240000, 48000), 1: (1440000, 4, 9, 1, 360000, 40000),
            2: (2400000, 10, 800, 147, 240000, 44100), 3: (1152000, 3, 1280, 147, 384000, 44100)}
    for mode, (fs, dec, ad, up, if_fs, afs) in want.items():
        cfg = capi.Config()
        capi.check(L.sdrb_config_for_mode(mode, ord("s"), 4, C.byref(cfg)))
        assert (cfg.rf_Fs, cfg.rf_decim, cfg.audio_decim, cfg.audio_upsample, cfg.if_Fs, cfg.audio_Fs) == (fs, dec, ad, up, if_fs, afs)
        assert cfg.rf_taps == 101 and cfg.rf_Fc == 100000 and cfg.audio_Fc == 16000 and cfg.rds_on == 0 and cfg.n_streams == 4
    cfg = capi.Config()
    capi.check(L.sdrb_config_for_mode(0, ord("r"), 1, C.byref(cfg)))
    assert cfg.rds_on == 1
    assert L.sdrb_config_for_mode(7, ord("m"), 1, C.byref(cfg)) == capi.SDRB_ERR_INVALID
    assert L.sdrb_config_for_mode(0, ord("x"), 1, C.byref(cfg)) == capi.SDRB_ERR_INVALID
    assert b"type" in L.sdrb_last_error()


def test_no_gpu_means_loud_failure(capi):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.SdrError) as e:
        capi.Chain(0, "r", 1)
    assert e.value.code in (capi.SDRB_ERR_NO_DEVICE, capi.SDRB_ERR_CUDA)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_invalid_configs_rejected_before_touching_the_gpu(capi):
    L = capi.lib()
    cfg = capi.Config()
    capi.check(L.sdrb_config_for_mode(0, ord("m"), 1, C.byref(cfg)))
    h = C.c_void_p()
    for field, val in (("n_streams", 0), ("rf_taps", 51), ("rf_decim", 7), ("type", ord("q"))):
        bad = capi.Config.from_buffer_copy(cfg)
        setattr(bad, field, val)
        assert L.sdrb_chain_create(C.byref(bad), C.byref(h)) == capi.SDRB_ERR_INVALID, field
