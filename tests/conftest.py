"""Shared fixtures.  `gpu` marks tests that need a B200 (run with `-m gpu`); everything else runs on CPU."""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); selected with -m gpu")


def load_module(name: str, relpath: str):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Everything native is built once per session (no-op when up to date)."""
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session")
def capi(_built):
    return load_module("sdrb_capi", "real-time-sdr_b200/capi.py")


@pytest.fixture(scope="session")
def sdrgen():
    return load_module("sdrgen", "real-time-sdr_b200/sdrgen.py")


@pytest.fixture(scope="session")
def oracle_mod(_built):
    return load_module("oracle_py", "oracle/oracle_py.py")


@pytest.fixture(scope="session")
def oracle(oracle_mod):
    return oracle_mod.Oracle()


@pytest.fixture(scope="session")
def ref(oracle_mod):
    r = oracle_mod.RefHarness()
    if not r.available():
        pytest.skip("oracle/_ref/ref_harness not built (needs /root/reference at build time)")
    return r


@pytest.fixture(scope="session")
def station_iq(sdrgen):
    """Cached synthetic IQ per (station index, mode, blocks)."""
    cache = {}

    def get(k: int, mode: int, nblocks: int) -> np.ndarray:
        key = (k, mode)
        need = sdrgen.block_pairs(mode) * nblocks
        if key not in cache or cache[key].size < 2 * need:
            cache[key] = sdrgen.generate_iq(sdrgen.Station.for_stream(k, fs=sdrgen.mode_fs(mode)), need)
        return cache[key][: 2 * need]

    return get


def bits_equal(a: np.ndarray, b: np.ndarray) -> bool:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()
