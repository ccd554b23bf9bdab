"""ctypes binding of libsdr_b200.so (include/sdr_b200.h) for the Python test-suite and bench.py.

The product's host side is C++ (real-time-sdr_b200/host/dy4_api.h mirrors the reference's C++ API on
top of the same C ABI); this module only lets Python drive the C ABI: it holds no DSP and has no
fallback — if the shared library is missing or no GPU is usable, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDRB_LIB") or os.path.join(_HERE, "libsdr_b200.so")  # SDRB_LIB: a variant build (tools/ only)

SDRB_OK, SDRB_ERR_INVALID, SDRB_ERR_CUDA, SDRB_ERR_NO_DEVICE, SDRB_ERR_STATE = range(5)

# every symbol include/sdr_b200.h declares (tests check the library exports all of them)
EXPORTS = [
    "sdrb_last_error", "sdrb_version", "sdrb_design_lpf", "sdrb_design_lpf_gain", "sdrb_design_bpf",
    "sdrb_design_apf", "sdrb_design_rrc", "sdrb_fir_decim", "sdrb_fir_updown", "sdrb_fm_demod", "sdrb_pll",
    "sdrb_cdr", "sdrb_config_for_mode", "sdrb_chain_create", "sdrb_chain_destroy", "sdrb_chain_get_info",
    "sdrb_chain_process_device", "sdrb_chain_process_host", "sdrb_chain_sync", "sdrb_chain_read_pcm",
    "sdrb_chain_pcm_device", "sdrb_chain_read_rds", "sdrb_rds_parse", "sdrb_chain_stage",
    "sdrb_chain_state_bytes", "sdrb_chain_state_save", "sdrb_chain_state_load", "sdrb_chain_kernel_times",
    "sdrb_chain_set_profiling", "sdrb_chain_launch_count", "sdrb_chain_set_overlap", "sdrb_pinned_alloc",
    "sdrb_pinned_free", "sdrb_chain_set_stream", "sdrb_chain_join", "sdrb_chain_read_results",
    "sdrb_manchester_decode", "sdrb_differential_decode", "sdrb_frame_sync",
    "sdrb_chain_state_load_n", "sdrb_chain_input_consumed", "sdrb_chain_rds_overflows",
    "sdrb_chain_pll_redos", "sdrb_chain_pll_redo_detail", "sdrb_chain_check_guards", "sdrb_chain_sm_partition",
    "sdrb_chain_state_item_offset", "sdrb_rds_sync",
]


class Config(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "rf_Fs", "rf_Fc", "rf_taps", "rf_decim", "audio_decim", "audio_upsample", "if_Fs", "audio_Fc", "audio_Fs",
        "symbol_Fs", "rds_on", "type", "n_streams", "device", "keep_stages")]


class ChainInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "block_pairs", "block_bytes", "if_block", "audio_block", "pcm_per_block", "rds_block", "max_bits", "max_groups")]


class RdsRecord(C.Structure):
    _fields_ = [("cdr_offset", C.c_int32), ("n_symbols", C.c_int32), ("n_bits", C.c_int32), ("n_groups", C.c_int32),
                ("bits", C.c_uint8 * 48), ("groups", C.c_uint64 * 8)]


MANCHESTER_STATE_DTYPE = np.dtype([("half_symbol", "<i4"), ("start", "<i4")])
FRAMESYNC_STATE_DTYPE = np.dtype([("reg", "<u8"), ("window", "<i4", (4,)), ("nwindow", "<i4"), ("ncarry", "<i4"), ("carry", "u1", (64,))])
RDS_SYNC_STATE_DTYPE = np.dtype([("reg", "<u8"), ("sync", "<i4"), ("prevsync", "<i4"), ("lastseen_offset", "<i4"), ("rds_bit_cont", "<i4"),
                                 ("lastseen_offset_cont", "<i4"), ("block_distance", "<i4"), ("block_number", "<i4"), ("block_bit_cont", "<i4"),
                                 ("blocks_cont", "<i4"), ("wrong_blocks_cont", "<i4"), ("group_assembly_started", "<i4"),
                                 ("group_good_blocks_cont", "<i4"), ("ext_reg", "<u8"), ("ext_good", "<i4"), ("reserved", "<i4")])
RDS_SYNC_EVENT_DTYPE = np.dtype([("type", "<i4"), ("bit", "<i4"), ("a", "<i4"), ("b", "<i4"), ("value", "<u8")])
assert RDS_SYNC_STATE_DTYPE.itemsize == 72 and RDS_SYNC_EVENT_DTYPE.itemsize == 24
RDS_RECORD_DTYPE = np.dtype([("cdr_offset", "<i4"), ("n_symbols", "<i4"), ("n_bits", "<i4"), ("n_groups", "<i4"),
                             ("bits", "u1", (48,)), ("groups", "<u8", (8,))])
assert RDS_RECORD_DTYPE.itemsize == C.sizeof(RdsRecord)


class PllState(C.Structure):
    _fields_ = [("feedbackI", C.c_float), ("feedbackQ", C.c_float), ("integrator", C.c_float), ("phaseEst", C.c_float),
                ("trigOffset", C.c_double), ("lastCarrier", C.c_float), ("last_out", C.c_float)]


PLL_STATE_DTYPE = np.dtype([("feedbackI", "<f4"), ("feedbackQ", "<f4"), ("integrator", "<f4"), ("phaseEst", "<f4"),
                            ("trigOffset", "<f8"), ("lastCarrier", "<f4"), ("last_out", "<f4")])
assert PLL_STATE_DTYPE.itemsize == C.sizeof(PllState)


class SdrError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsdr_b200 error {code}: {msg}")
        self.code = code


_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def load(path: str | None = None) -> C.CDLL:
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    L.sdrb_last_error.restype = C.c_char_p
    for name in ("sdrb_design_lpf",):
        getattr(L, name).argtypes = [C.c_float, C.c_float, C.c_int, _f32p]
    L.sdrb_design_lpf_gain.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, _f32p]
    L.sdrb_design_bpf.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, _f32p]
    L.sdrb_design_apf.argtypes = [C.c_float, C.c_int, _f32p]
    L.sdrb_design_rrc.argtypes = [C.c_float, C.c_int, _f32p]
    vp, sz, ci, cf = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    L.sdrb_fir_decim.argtypes = [vp, sz, ci, _f32p, ci, vp, vp, sz, ci, ci, vp]
    L.sdrb_fir_updown.argtypes = [vp, sz, ci, _f32p, ci, vp, ci, vp, sz, ci, ci, ci, vp]
    L.sdrb_fm_demod.argtypes = [vp, vp, sz, ci, vp, vp, sz, ci, vp]
    L.sdrb_pll.argtypes = [vp, sz, ci, cf, cf, cf, cf, cf, vp, vp, sz, ci, vp]
    L.sdrb_cdr.argtypes = [vp, sz, ci, ci, vp, ci, vp]
    L.sdrb_manchester_decode.argtypes = [vp, sz, vp, ci, vp, vp, sz, vp, ci, vp]
    L.sdrb_differential_decode.argtypes = [vp, sz, vp, ci, vp, vp, sz, ci, vp]
    L.sdrb_frame_sync.argtypes = [vp, sz, vp, ci, vp, vp, sz, vp, ci, ci, vp]
    L.sdrb_rds_sync.argtypes = [vp, sz, vp, ci, vp, vp, sz, vp, ci, vp, ci, vp]
    L.sdrb_config_for_mode.argtypes = [ci, ci, ci, C.POINTER(Config)]
    L.sdrb_chain_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.sdrb_chain_destroy.argtypes = [vp]
    L.sdrb_chain_get_info.argtypes = [vp, C.POINTER(ChainInfo)]
    L.sdrb_chain_process_device.argtypes = [vp, vp, sz]
    L.sdrb_chain_process_host.argtypes = [vp, vp, sz]
    L.sdrb_chain_sync.argtypes = [vp]
    L.sdrb_chain_read_pcm.argtypes = [vp, vp, sz]
    L.sdrb_chain_pcm_device.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.sdrb_chain_read_rds.argtypes = [vp, vp]
    L.sdrb_rds_parse.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_char_p, ci]
    L.sdrb_chain_stage.argtypes = [vp, C.c_char_p, vp, ci, C.POINTER(ci)]
    L.sdrb_chain_state_bytes.argtypes = [vp]
    L.sdrb_chain_state_bytes.restype = sz
    L.sdrb_chain_state_save.argtypes = [vp, vp]
    L.sdrb_chain_state_load.argtypes = [vp, vp]
    L.sdrb_chain_state_load_n.argtypes = [vp, vp, sz]
    L.sdrb_chain_input_consumed.argtypes = [vp, ci]
    L.sdrb_chain_state_item_offset.argtypes = [vp, C.c_char_p]
    L.sdrb_chain_state_item_offset.restype = C.c_longlong
    L.sdrb_chain_check_guards.argtypes = [vp, C.POINTER(ci)]
    L.sdrb_chain_pll_redo_detail.argtypes = [vp, C.POINTER(C.c_ulonglong * 20)]
    L.sdrb_chain_pll_redos.argtypes = [vp, C.POINTER(C.c_ulonglong * 2)]
    L.sdrb_chain_rds_overflows.argtypes = [vp, C.POINTER(C.c_uint * 3)]
    L.sdrb_chain_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(cf), ci, C.POINTER(ci)]
    L.sdrb_chain_set_profiling.argtypes = [vp, ci]
    L.sdrb_chain_sm_partition.argtypes = [vp, C.POINTER(C.c_int * 2)]
    L.sdrb_chain_launch_count.argtypes = [vp]
    L.sdrb_chain_launch_count.restype = C.c_longlong
    L.sdrb_chain_set_overlap.argtypes = [vp, ci]
    L.sdrb_chain_set_stream.argtypes = [vp, vp]
    L.sdrb_chain_join.argtypes = [vp]
    L.sdrb_chain_read_results.argtypes = [vp, ci, vp, sz, vp]
    L.sdrb_pinned_alloc.argtypes = [sz, C.POINTER(vp)]
    L.sdrb_pinned_free.argtypes = [vp]
    return L


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc: int) -> None:
    if rc != SDRB_OK:
        raise SdrError(rc, lib().sdrb_last_error().decode("utf-8", "replace"))


def design(kind: str, n: int, Fs: float = 0.0, a: float = 0.0, b: float = 0.0, u: int = 1) -> np.ndarray:
    """Tap designers (host side of the library; no GPU needed)."""
    L = lib()
    h = np.zeros(n, np.float32)
    if kind == "lpf":
        check(L.sdrb_design_lpf(Fs, a, n, h))
    elif kind == "lpf_gain":
        check(L.sdrb_design_lpf_gain(Fs, a, n, u, h))
    elif kind == "bpf":
        check(L.sdrb_design_bpf(Fs, a, b, n, h))
    elif kind == "apf":
        check(L.sdrb_design_apf(a, n, h))
    elif kind == "rrc":
        check(L.sdrb_design_rrc(Fs, n, h))
    else:
        raise ValueError(kind)
    return h


class PinnedBuffer:
    """Page-locked host memory from the library (cudaHostAlloc), viewed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        check(lib().sdrb_pinned_alloc(nbytes, C.byref(self.ptr)))
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def free(self):
        if self.ptr:
            lib().sdrb_pinned_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Chain:
    """One batched receive chain (n_streams stations on one GPU)."""

    def __init__(self, mode: int, kind: str, n_streams: int = 1, device: int = 0, keep_stages: bool = False):
        L = self.L = lib()
        self.cfg = Config()
        check(L.sdrb_config_for_mode(mode, ord(kind), n_streams, C.byref(self.cfg)))
        self.cfg.device = device
        self.cfg.keep_stages = 1 if keep_stages else 0
        self.h = C.c_void_p()
        check(L.sdrb_chain_create(C.byref(self.cfg), C.byref(self.h)))
        self.info = ChainInfo()
        check(L.sdrb_chain_get_info(self.h, C.byref(self.info)))
        self.n_streams = n_streams
        self.kind = kind
        self.mode = mode

    def close(self):
        if self.h:
            self.L.sdrb_chain_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- processing
    def process_host(self, iq: np.ndarray, pitch: int | None = None):
        """iq: uint8 [n_streams, >= block_bytes] (C-contiguous rows); pitch in bytes."""
        assert iq.dtype == np.uint8
        pitch = pitch if pitch is not None else (iq.strides[0] if iq.ndim == 2 else iq.size // self.n_streams)
        check(self.L.sdrb_chain_process_host(self.h, iq.ctypes.data, pitch))

    def process_host_ptr(self, ptr: int, pitch: int):
        check(self.L.sdrb_chain_process_host(self.h, ptr, pitch))

    def process_device(self, ptr: int, pitch: int):
        check(self.L.sdrb_chain_process_device(self.h, ptr, pitch))

    def sync(self):
        check(self.L.sdrb_chain_sync(self.h))

    def join(self):
        check(self.L.sdrb_chain_join(self.h))

    def read_results(self, lag: int, pcm: np.ndarray | None, rec: np.ndarray | None):
        check(self.L.sdrb_chain_read_results(self.h, lag, pcm.ctypes.data if pcm is not None else None,
                                             (pcm.strides[0] // 2) if pcm is not None else 0,
                                             rec.ctypes.data if rec is not None else None))

    def read_pcm(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.n_streams, self.info.pcm_per_block), np.int16)
        check(self.L.sdrb_chain_read_pcm(self.h, out.ctypes.data, out.strides[0] // 2))
        return out

    def read_rds(self) -> np.ndarray:
        rec = np.zeros(self.n_streams, RDS_RECORD_DTYPE)
        check(self.L.sdrb_chain_read_rds(self.h, rec.ctypes.data))
        return rec

    def stage(self, name: str) -> np.ndarray:
        cap = max(self.info.if_block + 1, self.info.rds_block, self.info.audio_block)
        out = np.zeros((self.n_streams, cap), np.float32)
        cnt = C.c_int(0)
        check(self.L.sdrb_chain_stage(self.h, name.encode(), out.ctypes.data, cap, C.byref(cnt)))
        return out[:, :cnt.value].copy()

    def state_save(self) -> bytes:
        n = self.L.sdrb_chain_state_bytes(self.h)
        buf = C.create_string_buffer(n)
        check(self.L.sdrb_chain_state_save(self.h, buf))
        return buf.raw

    def state_item_offset(self, name: str) -> int:
        return int(self.L.sdrb_chain_state_item_offset(self.h, name.encode()))

    def state_load(self, blob: bytes):
        check(self.L.sdrb_chain_state_load_n(self.h, blob, len(blob)))

    def input_consumed(self, lag: int = 0) -> bool:
        r = self.L.sdrb_chain_input_consumed(self.h, lag)
        if r < 0:
            raise SdrError(-r, self.L.sdrb_last_error().decode("utf-8", "replace"))
        return bool(r)

    def pll_redos(self) -> tuple:
        c = (C.c_ulonglong * 2)()
        check(self.L.sdrb_chain_pll_redos(self.h, C.byref(c)))
        return tuple(int(v) for v in c)

    def check_guards(self) -> int:
        n = C.c_int(0)
        check(self.L.sdrb_chain_check_guards(self.h, C.byref(n)))
        return n.value

    def pll_redo_detail(self) -> list:
        c = (C.c_ulonglong * 20)()
        check(self.L.sdrb_chain_pll_redo_detail(self.h, C.byref(c)))
        return [int(v) for v in c]

    def rds_overflows(self) -> tuple:
        c = (C.c_uint * 3)()
        check(self.L.sdrb_chain_rds_overflows(self.h, C.byref(c)))
        return tuple(int(v) for v in c)

    def set_profiling(self, on: bool):
        check(self.L.sdrb_chain_set_profiling(self.h, 1 if on else 0))

    def set_stream(self, cuda_stream: int):
        check(self.L.sdrb_chain_set_stream(self.h, cuda_stream))

    def set_overlap(self, on: bool):
        check(self.L.sdrb_chain_set_overlap(self.h, 1 if on else 0))

    def kernel_times(self) -> dict:
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = C.c_int(0)
        check(self.L.sdrb_chain_kernel_times(self.h, names, ms, 32, C.byref(n)))
        return {names[i].decode(): float(ms[i]) for i in range(n.value)}

    def launch_count(self) -> int:
        return int(self.L.sdrb_chain_launch_count(self.h))

    def sm_partition(self) -> tuple:
        """(SMs owned by the PLL stream, SMs owned by the FIR streams); (0, 0): no partition (plain priority streams)."""
        c = (C.c_int * 2)()
        check(self.L.sdrb_chain_sm_partition(self.h, C.byref(c)))
        return int(c[0]), int(c[1])


class RdsTextDecoder:
    """Host-side parse() state for one stream (chars/output of src/rds.cpp:68-69): group registers -> stderr text."""

    def __init__(self):
        self.chars = C.c_uint64(0)
        self.output = C.c_uint64(0)
        self.text = b""

    def feed(self, group: int) -> bytes:
        buf = C.create_string_buffer(256)
        n = lib().sdrb_rds_parse(int(group), C.byref(self.chars), C.byref(self.output), buf, 256)
        self.text += buf.raw[:n]
        return buf.raw[:n]
