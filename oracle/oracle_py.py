"""TEST INFRASTRUCTURE ONLY.  Python bindings for the checkers under oracle/.

* `Oracle`     — ctypes binding of oracle/liboracle.so (the C restatement, sdr_oracle.c)
* `RefHarness` — subprocess wrapper around oracle/_ref/ref_harness (the unmodified reference sources
                  compiled by oracle/Makefile; present only where it was built)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
from __future__ import annotations

import ctypes as C
import importlib.util
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def _load_recfile():
    spec = importlib.util.spec_from_file_location("_sdr_recfile", os.path.join(_ROOT, "real-time-sdr_b200", "recfile.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


recfile = _load_recfile()

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i16p = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class PllState(C.Structure):
    _fields_ = [("feedbackI", C.c_float), ("feedbackQ", C.c_float), ("integrator", C.c_float),
                ("phaseEst", C.c_float), ("trigOffset", C.c_double), ("lastCarrier", C.c_float)]


class SyncState(C.Structure):
    _fields_ = [("carry", C.c_int * 64), ("ncarry", C.c_int), ("reg", C.c_uint64), ("chars", C.c_uint64),
                ("output", C.c_uint64), ("first_time", C.c_int), ("window", C.c_int * 4), ("nwindow", C.c_int)]


class ErrdetState(C.Structure):
    _fields_ = [("reg", C.c_uint64), ("chars", C.c_uint64), ("output", C.c_uint64), ("first_time", C.c_int), ("hex", C.c_int)] + \
               [(n, C.c_int) for n in ("sync", "prevsync", "lastseen_offset", "rds_bit_cont", "lastseen_offset_cont", "block_distance",
                                       "block_number", "block_bit_cont", "blocks_cont", "wrong_blocks_cont", "group_assembly_started",
                                       "group_good_blocks_cont")]


class ErrdetEvent(C.Structure):
    _fields_ = [("type", C.c_int), ("bit", C.c_int), ("a", C.c_int), ("b", C.c_int), ("value", C.c_uint64)]


ERRDET_STATE_NAMES = ("sync", "prevsync", "lastseen_offset", "rds_bit_cont", "lastseen_offset_cont", "block_distance", "block_number",
                      "block_bit_cont", "blocks_cont", "wrong_blocks_cont", "group_assembly_started", "group_good_blocks_cont")


class ChainInfo(C.Structure):
    _fields_ = [("mode", C.c_int), ("type", C.c_int), ("block_pairs", C.c_int), ("if_block", C.c_int),
                ("audio_block", C.c_int), ("rds_block", C.c_int)]


class Oracle:
    """The C restatement.  Array-in/array-out wrappers; state objects are explicit."""

    def __init__(self, path: str | None = None):
        path = path or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = self.lib = C.CDLL(path)
        L.orc_lpf.argtypes = [C.c_float, C.c_float, C.c_int, _f32p]
        L.orc_lpf_gain.argtypes = [C.c_float, C.c_float, C.c_int, C.c_int, _f32p]
        L.orc_bpf.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int, _f32p]
        L.orc_apf.argtypes = [C.c_float, C.c_int, _f32p]
        L.orc_rrc.argtypes = [C.c_float, C.c_int, _f32p]
        L.orc_fir_decim.argtypes = [_f32p, _f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.c_int]
        L.orc_fir_updown.argtypes = [_f32p, _f32p, C.c_int, _f32p, C.c_int, _f32p, C.c_int, C.c_int, C.c_int]
        L.orc_fir_updown.restype = C.c_int
        L.orc_fmdemod.argtypes = [_f32p, _f32p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), _f32p]
        L.orc_pll_init.argtypes = [C.POINTER(PllState)]
        L.orc_pll.argtypes = [_f32p, C.c_int, C.c_float, C.c_float, _f32p, C.POINTER(PllState), C.c_float, C.c_float,
                              C.c_float]
        L.orc_cdr.argtypes = [C.c_int, _f32p, C.c_int]
        L.orc_cdr.restype = C.c_int
        L.orc_manchester.argtypes = [_i32p, _i32p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_manchester.restype = C.c_int
        L.orc_differential.argtypes = [_i32p, _i32p, C.c_int, C.POINTER(C.c_int), C.c_int]
        L.orc_sync_init.argtypes = [C.POINTER(SyncState)]
        L.orc_block_offset.argtypes = [_i32p]
        L.orc_block_offset.restype = C.c_int
        L.orc_frame_sync.argtypes = [C.POINTER(SyncState), _i32p, C.c_int, _u64p, C.c_int, C.c_char_p, C.c_int]
        L.orc_frame_sync.restype = C.c_int
        L.orc_parse.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_char_p, C.c_int]
        L.orc_parse.restype = C.c_int
        L.orc_chain_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.orc_chain_create.restype = C.c_void_p
        L.orc_chain_destroy.argtypes = [C.c_void_p]
        L.orc_chain_get_info.argtypes = [C.c_void_p, C.POINTER(ChainInfo)]
        L.orc_chain_block.argtypes = [C.c_void_p, _u8p, _i16p]
        L.orc_chain_block.restype = C.c_int
        L.orc_chain_stage.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int)]
        L.orc_chain_stage.restype = C.POINTER(C.c_float)
        L.orc_chain_rds_block.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int)),
                                          C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.c_int)]
        L.orc_chain_groups.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint64))]
        L.orc_chain_groups.restype = C.c_int
        L.orc_chain_text.argtypes = [C.c_void_p]
        L.orc_chain_text.restype = C.c_char_p
        L.orc_chain_set_pll_sample_count.argtypes = [C.c_void_p, C.c_double]
        L.orc_run_batch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _u8p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_run_batch.restype = C.c_int

    # ---- designers
    def design(self, kind: str, n: int, Fs: float = 0.0, a: float = 0.0, b: float = 0.0, u: int = 1) -> np.ndarray:
        h = np.zeros(n, dtype=np.float32)
        if kind == "lpf":
            self.lib.orc_lpf(Fs, a, n, h)
        elif kind == "lpf_gain":
            self.lib.orc_lpf_gain(Fs, a, n, u, h)
        elif kind == "bpf":
            self.lib.orc_bpf(Fs, a, b, n, h)
        elif kind == "apf":
            self.lib.orc_apf(a, n, h)
        elif kind == "rrc":
            self.lib.orc_rrc(Fs, n, h)
        else:
            raise ValueError(kind)
        return h

    # ---- primitives over a sequence of blocks (x is split into nblocks equal blocks)
    def fir_decim(self, x, h, decim, nblocks=1, state=None):
        x = np.ascontiguousarray(x, np.float32)
        h = np.ascontiguousarray(h, np.float32)
        blk = x.size // nblocks
        st = np.zeros(h.size - 1, np.float32) if state is None else np.ascontiguousarray(state, np.float32).copy()
        out = []
        for b in range(nblocks):
            y = np.zeros(blk // decim, np.float32)
            self.lib.orc_fir_decim(y, x[b * blk:(b + 1) * blk], blk, h, h.size, st, st.size, decim)
            out.append(y)
        return np.concatenate(out)

    def fir_updown(self, x, h, up, down, nblocks=1, nstate=100):
        x = np.ascontiguousarray(x, np.float32)
        h = np.ascontiguousarray(h, np.float32)
        blk = x.size // nblocks
        st = np.zeros(nstate, np.float32)
        out = []
        for b in range(nblocks):
            y = np.zeros(blk * up // down + 1, np.float32)
            ny = self.lib.orc_fir_updown(y, x[b * blk:(b + 1) * blk], blk, h, h.size, st, nstate, up, down)
            out.append(y[:ny])
        return np.concatenate(out)

    def fmdemod(self, I, Q, nblocks=1):
        I = np.ascontiguousarray(I, np.float32)
        Q = np.ascontiguousarray(Q, np.float32)
        blk = I.size // nblocks
        pi, pq = C.c_float(0), C.c_float(0)
        out = np.zeros(I.size, np.float32)
        for b in range(nblocks):
            self.lib.orc_fmdemod(I[b * blk:(b + 1) * blk], Q[b * blk:(b + 1) * blk], blk, C.byref(pi), C.byref(pq),
                                 out[b * blk:(b + 1) * blk])
        return out, (pi.value, pq.value)

    def pll(self, x, freq, Fs, scale, adjust, bw, nblocks=1):
        """Returns the (blk+1)-long output vector of every block, concatenated, and the final state."""
        x = np.ascontiguousarray(x, np.float32)
        blk = x.size // nblocks
        st = PllState()
        self.lib.orc_pll_init(C.byref(st))
        out = np.zeros(blk + 1, np.float32)
        out[blk] = 1.0
        ys = []
        for b in range(nblocks):
            self.lib.orc_pll(x[b * blk:(b + 1) * blk], blk, freq, Fs, out, C.byref(st), scale, adjust, bw)
            ys.append(out.copy())
        return np.concatenate(ys), st

    def cdr(self, x, sps, nblocks=1):
        x = np.ascontiguousarray(x, np.float32)
        blk = x.size // nblocks
        return np.array([self.lib.orc_cdr(sps, x[b * blk:(b + 1) * blk], blk) for b in range(nblocks)], np.int32)

    def bits(self, symbols_per_block, block0=6):
        """manchester + differential over consecutive blocks; returns (manchester, decoded, lens, state)."""
        half, start, last = C.c_int(0), C.c_int(0), C.c_int(0)
        man_all, dec_all, lens = [], [], []
        bc = block0
        for s in symbols_per_block:
            s = np.ascontiguousarray(s, np.int32)
            bits = np.zeros(s.size + 2, np.int32)
            nb = self.lib.orc_manchester(bits, s, s.size, bc, C.byref(half), C.byref(start))
            dec = np.zeros(max(nb, 1), np.int32)
            self.lib.orc_differential(dec, bits, nb, C.byref(last), bc)
            man_all.append(bits[:nb].copy())
            dec_all.append(dec[:nb].copy())
            lens.append(nb)
            bc += 1
        return (np.concatenate(man_all), np.concatenate(dec_all), np.array(lens, np.int32),
                (half.value, start.value, last.value))

    def frame_sync(self, chunks):
        st = SyncState()
        self.lib.orc_sync_init(C.byref(st))
        text = C.create_string_buffer(1 << 16)
        groups, per_call = [], []
        for ch in chunks:
            ch = np.ascontiguousarray(ch, np.int32)
            g = np.zeros(32, np.uint64)
            n = self.lib.orc_frame_sync(C.byref(st), ch, ch.size, g, 32, text, len(text))
            groups.extend(g[:n].tolist())
            per_call.append(n)
        carry = np.array(st.carry[:st.ncarry], np.int32)
        return (np.array(groups, np.uint64), np.array(per_call, np.int32), text.value.decode("latin-1"),
                (st.reg, st.chars, st.output), carry)

    def error_detection(self, chunks, debug_lines=True):
        """The reference's never-called sync-state-machine decoder (src/rds_utilities.cpp:202-311) over bit chunks.
        Returns (events [(type, bit, a, b, value)], text, state64 (reg, chars, output), state dict, n_unsynced)."""
        L = self.lib
        L.orc_errdet_init.argtypes = [C.POINTER(ErrdetState)]
        L.orc_error_detection.argtypes = [C.POINTER(ErrdetState), _i32p, C.c_int, C.POINTER(ErrdetEvent), C.c_int, C.c_char_p, C.c_int,
                                          C.c_int, C.POINTER(C.c_longlong)]
        L.orc_error_detection.restype = C.c_int
        st = ErrdetState()
        L.orc_errdet_init(C.byref(st))
        total = int(sum(len(c) for c in chunks))
        text = C.create_string_buffer(max(1 << 16, 80 * total + 4096))
        nun = C.c_longlong(0)
        events = []
        for ch in chunks:
            ch = np.ascontiguousarray(ch, np.int32)
            cap = ch.size // 26 + 8
            ev = (ErrdetEvent * cap)()
            n = L.orc_error_detection(C.byref(st), ch, ch.size, ev, cap, text, len(text), 1 if debug_lines else 0, C.byref(nun))
            assert n <= cap
            events.extend((e.type, e.bit, e.a, e.b, e.value) for e in ev[:n])
        state = {k: getattr(st, k) for k in ERRDET_STATE_NAMES}
        return events, text.value, (st.reg, st.chars, st.output), state, nun.value

    def parse_groups(self, regs):
        chars, output = C.c_uint64(0), C.c_uint64(0)
        text = C.create_string_buffer(1 << 16)
        for r in regs:
            self.lib.orc_parse(int(r), C.byref(chars), C.byref(output), text, len(text))
        return text.value.decode("latin-1")

    # ---- whole chain
    def chain(self, mode: int, kind: str, iq: np.ndarray, stages=(), with_rds_dsp=False, max_blocks=None,
              pll_sample_count: float | None = None) -> dict:
        """Run one stream; returns the same record names as RefHarness.chain()."""
        L = self.lib
        c = L.orc_chain_create(mode, ord(kind), 1 if (with_rds_dsp or "rds_clean" in stages) else 0)
        if pll_sample_count is not None:
            L.orc_chain_set_pll_sample_count(c, float(pll_sample_count))
        info = ChainInfo()
        L.orc_chain_get_info(c, C.byref(info))
        iq = np.ascontiguousarray(iq, np.uint8)
        nblocks = iq.size // (2 * info.block_pairs)
        if max_blocks is not None:
            nblocks = min(nblocks, max_blocks)
        per = info.audio_block * (1 if kind == "m" else 2)
        pcm = np.zeros(nblocks * per, np.int16)
        acc = {s: [] for s in stages}
        offs, nsyms, nbits, bits_all, groups, group_block, syms_all = [], [], [], [], [], [], []
        for b in range(nblocks):
            L.orc_chain_block(c, iq[b * 2 * info.block_pairs:(b + 1) * 2 * info.block_pairs], pcm[b * per:(b + 1) * per])
            for s in stages:
                cnt = C.c_int(0)
                p = L.orc_chain_stage(c, s.encode(), C.byref(cnt))
                if p:
                    acc[s].append(np.ctypeslib.as_array(p, shape=(cnt.value,)).copy())
            if kind == "r":
                off, ns, nb = C.c_int(0), C.c_int(0), C.c_int(0)
                sp, bp = C.POINTER(C.c_int)(), C.POINTER(C.c_int)()
                L.orc_chain_rds_block(c, C.byref(off), C.byref(sp), C.byref(ns), C.byref(bp), C.byref(nb))
                offs.append(off.value)
                nsyms.append(ns.value)
                nbits.append(nb.value)
                if nb.value:
                    bits_all.append(np.ctypeslib.as_array(bp, shape=(nb.value,)).copy())
                    syms_all.append(np.ctypeslib.as_array(sp, shape=(ns.value,)).copy())
                gp = C.POINTER(C.c_uint64)()
                ng = L.orc_chain_groups(c, C.byref(gp))
                for g in range(ng):
                    groups.append(gp[g])
                    group_block.append(b)
        out = {"meta": np.array([mode, ord(kind), nblocks, info.block_pairs, info.if_block], np.int32), "pcm": pcm}
        for s in stages:
            if acc[s]:
                out[s] = np.concatenate(acc[s])
        if kind == "r":
            out["cdr_offset"] = np.array(offs, np.int32)
            out["n_symbols"] = np.array(nsyms, np.int32)
            out["n_bits"] = np.array(nbits, np.int32)
            out["rds_bits"] = np.concatenate(bits_all).astype(np.int32) if bits_all else np.zeros(0, np.int32)
            out["symbols"] = np.concatenate(syms_all).astype(np.int32) if syms_all else np.zeros(0, np.int32)
            out["groups"] = np.array(groups, np.uint64)
            out["group_block"] = np.array(group_block, np.int32)
            out["text"] = np.frombuffer(L.orc_chain_text(c), dtype=np.uint8).copy()
        L.orc_chain_destroy(c)
        return out

    def run_batch(self, mode, kind, iq, nstreams, nblocks, nthreads, want_pcm=False):
        iq = np.ascontiguousarray(iq, np.uint8)
        groups = np.zeros(nstreams, np.int32)
        pcm = None
        if want_pcm:
            per = 1470 * (1 if kind == "m" else 2)
            pcm = np.zeros(nstreams * nblocks * per, np.int16)
        self.lib.orc_run_batch(mode, ord(kind), nstreams, nblocks, iq, pcm.ctypes.data if pcm is not None else None,
                               groups.ctypes.data, nthreads)
        return groups, pcm


class RefHarness:
    """oracle/_ref/ref_harness: the reference's own functions, driven through files."""

    def __init__(self, path: str | None = None):
        self.path = path or os.path.join(_HERE, "_ref", "ref_harness")
        self.project = os.path.join(_HERE, "_ref", "project")

    def available(self) -> bool:
        return os.path.exists(self.path) and os.access(self.path, os.X_OK)

    def _run(self, args):
        subprocess.run([self.path] + args, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)

    def taps(self) -> dict:
        with tempfile.TemporaryDirectory() as d:
            out = os.path.join(d, "taps.rec")
            self._run(["taps", out])
            return recfile.read(out)

    def op(self, name: str, **arrays) -> dict:
        with tempfile.TemporaryDirectory() as d:
            inp, out = os.path.join(d, "in.rec"), os.path.join(d, "out.rec")
            recs = {"op": name}
            for k, v in arrays.items():
                if isinstance(v, (int, np.integer)):
                    v = np.array([v], np.int32)
                recs[k] = v
            recfile.write(inp, recs)
            self._run(["op", inp, out])
            return recfile.read(out)

    def error_detection(self, chunks) -> dict:
        bits = np.concatenate([np.asarray(c, np.int32) for c in chunks]) if len(chunks) else np.zeros(0, np.int32)
        return self.op("errdet", bits=bits, lens=np.array([len(c) for c in chunks], np.int32))

    def chain(self, mode: int, kind: str, iq: np.ndarray, stages=(), max_blocks=None) -> dict:
        with tempfile.TemporaryDirectory() as d:
            inp, out = os.path.join(d, "iq.raw"), os.path.join(d, "out.rec")
            np.ascontiguousarray(iq, np.uint8).tofile(inp)
            self._run(["chain", str(mode), kind, inp, out, str(-1 if max_blocks is None else max_blocks),
                       ",".join(stages) if stages else "out"])
            return recfile.read(out)
