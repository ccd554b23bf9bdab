"""Runs the CUDA chain and the oracle on the same IQ and compares every stage (shared by the GPU tests,
tools/parity_report.py and smoke)."""
from __future__ import annotations

import numpy as np

FLOAT_STAGES = {
    "m": ["I_ds", "Q_ds", "fm_demod", "audio_filt"],
    "s": ["I_ds", "Q_ds", "fm_demod", "pilot", "carrier", "stereo_band", "stereo_dc", "mono_filt", "stereo_filt"],
    "r": ["I_ds", "Q_ds", "fm_demod", "pilot", "carrier", "stereo_band", "stereo_dc", "mono_filt", "stereo_filt",
          "rds_band", "gen_pilot", "IPLL", "rds_band_delay", "rds_dc", "rds_filt", "rds_clean"],
}


def run_cuda_chain(capi, mode: int, kind: str, iq_streams: list, nblocks: int, stages=(), pitch_pad: int = 0,
                   device_input: bool = False, overlap: bool | None = None):
    """iq_streams: list of uint8 arrays (one per stream).  Returns per-stream dicts like Oracle.chain()."""
    S = len(iq_streams)
    out = [dict() for _ in range(S)]
    with capi.Chain(mode, kind, n_streams=S, keep_stages=bool(stages)) as ch:
        if overlap is not None:
            ch.set_overlap(overlap)
        bb = ch.info.block_bytes
        pitch = bb + pitch_pad
        acc = {st: [[] for _ in range(S)] for st in stages}
        pcm = [[] for _ in range(S)]
        rds = {k: [[] for _ in range(S)] for k in ("cdr_offset", "n_symbols", "n_bits", "bits", "groups", "group_block")}
        buf = np.zeros((S, pitch), np.uint8)
        dev = None
        if device_input:
            import torch
            dev = torch.zeros((S, pitch), dtype=torch.uint8, device="cuda")
        for b in range(nblocks):
            for s in range(S):
                buf[s, :bb] = iq_streams[s][b * bb:(b + 1) * bb]
            if device_input:
                import torch
                ch.sync()  # previous block may still be reading the device buffer
                dev.copy_(torch.from_numpy(buf))
                torch.cuda.synchronize()
                ch.process_device(dev.data_ptr(), pitch)
            else:
                ch.process_host(buf, pitch)
            p = ch.read_pcm()
            for st in stages:
                a = ch.stage(st)
                for s in range(S):
                    acc[st][s].append(a[s])
            rec = ch.read_rds() if kind == "r" else None
            for s in range(S):
                pcm[s].append(p[s].copy())
                if rec is not None:
                    r = rec[s]
                    rds["cdr_offset"][s].append(int(r["cdr_offset"]))
                    rds["n_symbols"][s].append(int(r["n_symbols"]))
                    rds["n_bits"][s].append(int(r["n_bits"]))
                    rds["bits"][s].append(r["bits"][: r["n_bits"]].astype(np.int32))
                    for g in range(int(r["n_groups"])):
                        rds["groups"][s].append(int(r["groups"][g]))
                        rds["group_block"][s].append(b)
        launches = ch.launch_count()
    for s in range(S):
        out[s]["pcm"] = np.concatenate(pcm[s])
        for st in stages:
            out[s][st] = np.concatenate(acc[st][s])
        if kind == "r":
            out[s]["cdr_offset"] = np.array(rds["cdr_offset"][s], np.int32)
            out[s]["n_symbols"] = np.array(rds["n_symbols"][s], np.int32)
            out[s]["n_bits"] = np.array(rds["n_bits"][s], np.int32)
            out[s]["rds_bits"] = np.concatenate(rds["bits"][s]) if rds["bits"][s] else np.zeros(0, np.int32)
            out[s]["groups"] = np.array(rds["groups"][s], np.uint64)
            out[s]["group_block"] = np.array(rds["group_block"][s], np.int32)
            dec = capi.RdsTextDecoder()
            for g in rds["groups"][s]:
                dec.feed(g)
            out[s]["text"] = np.frombuffer(dec.text, np.uint8).copy()
        out[s]["launches"] = launches
    return out


def diff_report(got: dict, want: dict, keys) -> dict:
    """Per key: None if bit-identical, else a dict describing the first mismatch."""
    rep = {}
    for k in keys:
        if k not in want:
            continue
        if k not in got:
            rep[k] = {"missing": True}
            continue
        g, w = np.asarray(got[k]), np.asarray(want[k])
        if g.shape != w.shape:
            rep[k] = {"shape_got": g.shape, "shape_want": w.shape}
            continue
        if g.dtype.kind == "f":
            ne = g.view(np.uint32) != w.astype(np.float32).view(np.uint32)
        else:
            ne = g.astype(np.int64) != w.astype(np.int64) if g.dtype != np.uint64 else g != w.astype(np.uint64)
        n = int(ne.sum())
        if n == 0:
            rep[k] = None
        else:
            i = int(np.argmax(ne))
            d = {"mismatches": n, "of": int(g.size), "first": i, "got": g[i].item(), "want": w[i].item()}
            if g.dtype.kind == "f":
                d["max_abs"] = float(np.nanmax(np.abs(g.astype(np.float64) - w.astype(np.float64))))
            rep[k] = d
    return rep


RDS_KEYS = ["cdr_offset", "n_symbols", "n_bits", "rds_bits", "groups", "group_block", "text"]
