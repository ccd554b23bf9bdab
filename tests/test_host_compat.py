"""Host-side boundary on a CPU-only box: the compat/ headers that let code written against the reference's headers
compile unchanged, the hand-off queue's contract, and the reference's own main() built against them (host/project_dropin).
The GPU behaviour of project_dropin is in tests/test_shim_gpu.py."""
from __future__ import annotations

import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "real-time-sdr_b200", "host")
REF_INCLUDE = "/root/reference/include"


def test_compat_covers_every_header_the_reference_main_includes(_built):
    """src/project.cpp:9-21 includes 13 project headers; each must exist under compat/ (same file name)."""
    want = ["dy4.h", "filter.h", "fourier.h", "genfunc.h", "iofunc.h", "logfunc.h", "mono.h", "stereo.h", "rds.h", "utilities.h",
            "rffrontend.h", "threadsafequeue.h", "args.h", "demod.h", "pll.h", "rds_utilities.h"]
    have = set(os.listdir(os.path.join(HOST, "compat")))
    assert set(want) <= have
    if os.path.isdir(REF_INCLUDE):  # where the reference is present: nothing it ships is missing
        assert set(os.listdir(REF_INCLUDE)) <= have


def test_queue_contract(tmp_path, _built):
    exe = tmp_path / "queue_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", os.path.join(HOST, "compat"), "-o", str(exe),
                    os.path.join(ROOT, "tests", "cpp", "queue_check.cpp")], check=True)
    for _ in range(3):
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0 and r.stdout.strip() == "ok"


def test_thread_bodies_are_exported(_built):
    out = subprocess.run(["nm", "-DC", "--defined-only", os.path.join(HOST, "libdy4_b200.so")], capture_output=True, text=True).stdout
    for sig in ("RF_frontend(args*)", "mono(args*)", "stereo(args*)", "rds(args*)", "fmpll(", "convolveFIR(", "start_frame_sync("):
        assert sig in out, sig


def _dropin():
    exe = os.path.join(HOST, "project_dropin")
    if not os.path.exists(exe):
        pytest.skip("host/project_dropin is built from /root/reference/src/project.cpp, which is not present here")
    return exe


def test_reference_main_builds_unchanged_and_rejects_bad_arguments_like_the_reference(_built):
    exe = _dropin()
    ref = os.path.join(ROOT, "oracle", "_ref", "project")
    for argv in (["9", "m"], ["0", "x"]):
        r = subprocess.run([exe] + argv, input=b"", capture_output=True, timeout=60)
        assert r.returncode == 1 and r.stdout == b"" and b"Valid modes are" in r.stderr
        if os.path.exists(ref):
            q = subprocess.run([ref] + argv, input=b"", capture_output=True, timeout=60)
            assert (q.returncode, q.stdout, q.stderr) == (r.returncode, r.stdout, r.stderr)


def test_dropin_without_a_gpu_fails_loudly(_built):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = subprocess.run([_dropin(), "0", "r"], input=b"\x80" * 147000, capture_output=True, timeout=120)
    assert r.returncode == 1 and r.stdout == b""
    assert b"no CPU fallback" in r.stderr or b"CUDA" in r.stderr
