"""Generates the committed golden fixtures from the UNMODIFIED reference sources.

Needs oracle/_ref/ref_harness (built by `make -C oracle` where /root/reference exists) — i.e. it runs in
the build container, not on the GPU box.  Outputs (small, committed):
  taps.npz            every tap table the chain uses, for all four modes (impulseResponse* of the reference)
  chain_m0_r.npz      mode 0 stereo+RDS, station 0, 40 blocks: PCM, CDR offsets, bits, groups, text, stage slices
  chain_misc.npz      PCM of mode 0 m / 0 s / 2 m / 1 s / 3 m (few blocks) and a checksum of each float stage
  ops.npz             per-function vectors: convolveFIR (both forms), fmDemodNoArctan, fmpll, cdr, bits, framesync
  parser_registers.npz  the 56 group registers of the reference's test/parser_test.cpp:79-136

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import os
import re
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import __graft_entry__ as g  # noqa: E402
import oracle_py  # noqa: E402


def crc(a: np.ndarray) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


def main():
    g.build()
    gen = g._load("sdrgen", os.path.join(ROOT, "real-time-sdr_b200", "sdrgen.py"))
    ref = oracle_py.RefHarness()
    assert ref.available(), "build oracle/_ref first (make -C oracle)"

    np.savez_compressed(os.path.join(HERE, "taps.npz"), **ref.taps())

    nblocks = 40
    iq = gen.generate_iq(gen.Station(), gen.block_pairs(0) * nblocks)
    r = ref.chain(0, "r", iq, stages=("fm_demod", "carrier", "rds_clean", "stereo_filt", "IPLL"))
    out = {"nblocks": np.int32(nblocks), "iq_crc": np.uint32(crc(iq))}
    for k in ("pcm", "cdr_offset", "n_symbols", "n_bits", "rds_bits", "groups", "group_block", "text"):
        out[k] = r[k]
    for st, off, n in (("fm_demod", 7350 * 7 - 300, 1200), ("carrier", 7351 * 20 - 200, 1000),
                       ("rds_clean", 2836 * 11 - 150, 900), ("stereo_filt", 1470 * 30 - 100, 600), ("IPLL", 7351 * 38, 800)):
        out[st + "_offset"] = np.int64(off)
        out[st + "_slice"] = r[st][off:off + n]
        out[st + "_crc"] = np.uint32(crc(r[st]))
    np.savez_compressed(os.path.join(HERE, "chain_m0_r.npz"), **out)

    misc = {}
    for mode, kind, nb in ((0, "m", 6), (0, "s", 6), (2, "m", 5), (1, "s", 4), (3, "m", 4), (1, "m", 3), (2, "s", 3)):
        iqm = gen.generate_iq(gen.Station.for_stream(0, fs=gen.mode_fs(mode)), gen.block_pairs(mode) * nb)
        rr = ref.chain(mode, kind, iqm, stages=("fm_demod",))
        tag = f"m{mode}_{kind}"
        misc[tag + "_nblocks"] = np.int32(nb)
        misc[tag + "_pcm"] = rr["pcm"]
        misc[tag + "_fm_crc"] = np.uint32(crc(rr["fm_demod"]))
        misc[tag + "_iq_crc"] = np.uint32(crc(iqm))
    np.savez_compressed(os.path.join(HERE, "chain_misc.npz"), **misc)

    rng = np.random.default_rng(2026)
    ops = {}
    x = rng.standard_normal(3000).astype(np.float32)
    h = (rng.standard_normal(101) / 20).astype(np.float32)
    ops["fir_x"], ops["fir_h"] = x, h
    ops["fir_decim10_y"] = ref.op("fir_decim", x=x, h=h, decim=10, nblocks=3)["y"]
    ops["fir_decim1_y"] = ref.op("fir_decim", x=x, h=h, decim=1, nblocks=3)["y"]
    h3 = (rng.standard_normal(303) / 20).astype(np.float32)
    ops["fir_h3"] = h3
    ops["fir_updown_3_7_y"] = ref.op("fir_updown", x=x, h=h3, up=3, down=7, nblocks=3)["y"]
    ops["fir_updown_1_5_y"] = ref.op("fir_updown", x=x, h=h, up=1, down=5, nblocks=3)["y"]
    I, Q = rng.standard_normal(600).astype(np.float32), rng.standard_normal(600).astype(np.float32)
    I[7] = Q[7] = 0
    d = ref.op("fmdemod", I=I, Q=Q, nblocks=2)
    ops["dem_I"], ops["dem_Q"], ops["dem_y"], ops["dem_prev"] = I, Q, d["y"], d["prev"]
    t = np.arange(6000)
    for tag, freq, scale, bw in (("pll19", 19e3, 2.0, 0.01), ("pll114", 114e3, 0.5, 0.001)):
        xin = (0.05 * np.cos(2 * np.pi * (freq + 4.0) / 240000.0 * t + 0.3) + 0.003 * rng.standard_normal(t.size)).astype(np.float32)
        p = ref.op("pll", x=xin, p=np.array([freq, 240000.0, scale, 0.0, bw], np.float32), nblocks=3)
        ops[tag + "_x"], ops[tag + "_y"], ops[tag + "_state"], ops[tag + "_trig"] = xin, p["y"], p["state"], p["trig_offset"]
    xc = (rng.standard_normal(2836 * 3) * 2.2).astype(np.float32)
    ops["cdr_x"] = xc
    ops["cdr_offset"] = ref.op("cdr", x=xc, sps=39, nblocks=3)["offset"]
    lens = np.array([73, 72, 73, 73, 72, 73, 72], np.int32)
    sym = rng.integers(0, 2, int(lens.sum())).astype(np.int32)
    b = ref.op("bits", symbols=sym, lens=lens, block0=6)
    ops["bits_symbols"], ops["bits_lens"] = sym, lens
    ops["bits_manchester"], ops["bits_decoded"], ops["bits_out_lens"], ops["bits_state"] = b["manchester"], b["decoded"], b["lens"], b["state"]
    # frame sync on the bit stream of the real chain, cut in the chain's own 15-block chunks
    nb15 = r["n_bits"][6:]
    chunks = [int(nb15[i:i + 15].sum()) for i in range(0, len(nb15) - len(nb15) % 15, 15)]
    bits = r["rds_bits"][: sum(chunks)]
    f = ref.op("framesync", bits=bits.astype(np.int32), lens=np.array(chunks, np.int32))
    ops["fs_bits"], ops["fs_lens"] = bits.astype(np.int32), np.array(chunks, np.int32)
    ops["fs_groups"], ops["fs_groups_per_call"], ops["fs_text"], ops["fs_state"], ops["fs_carry"] = (
        f["groups"], f["groups_per_call"], f["text"], f["state"], f["carry"])
    np.savez_compressed(os.path.join(HERE, "ops.npz"), **ops)

    # error_detection (declared include/rds_utilities.h:14, defined src/rds_utilities.cpp:202-311, never called by the
    # reference): the unmodified function on the streams of tests/rds_streams.py, in the chain's ragged chunks
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import rds_streams
    ed = {}
    for name, bits in rds_streams.cases(gen).items():
        chunks = rds_streams.chunks_of(bits)
        e = ref.error_detection(chunks)
        ed[name + "_bits"] = bits.astype(np.int8)
        ed[name + "_lens"] = np.array([len(c) for c in chunks], np.int32)
        ed[name + "_text"], ed[name + "_state64"], ed[name + "_state"] = e["text"], e["state64"], e["state"]
    np.savez_compressed(os.path.join(HERE, "errdet.npz"), **ed)

    src = open("/root/reference/test/parser_test.cpp").read()
    body = src[src.index("uint64_t regs[] = {"):]
    body = body[: body.index("};")]
    regs = np.array([int(v) for v in re.findall(r"\b(\d{15,20})\b", body)], np.uint64)
    assert regs.size == 56
    np.savez_compressed(os.path.join(HERE, "parser_registers.npz"), regs=regs)
    for f_ in sorted(os.listdir(HERE)):
        if f_.endswith(".npz"):
            print(f_, os.path.getsize(os.path.join(HERE, f_)))


if __name__ == "__main__":
    main()
